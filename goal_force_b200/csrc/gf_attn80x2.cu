// gf_attn80x2.cu -- CTA-pair (cta_group::2) version of the decoupled 80-row-block attention kernel (gf_attn80.cu).
//
// Why: a tcgen05.mma whose A operand comes from shared memory costs max(N/2, 32 + N/4) cycles per K = 16 step at
// M = 128 (tools/microbench/mma_rate.cu, measured on B200): the tensor core re-reads the 4 KB A slice and N*32 B of B
// from shared memory at 128 B/clk.  For the QK product of gf_attn80.cu (N = 80) that is 52 cycles instead of the 40
// the arithmetic needs, so QK costs 416 cycles per tile and kv block instead of 320 and the tensor pipe is the
// limiter at 1472 of ~1670 cycles per block.  With a CTA pair the K block is split between the two SMs (each stages
// and feeds 40 of the 80 rows), so each tensor core reads 4 KB + 1.25 KB per step: 42 cycles, compute-bound again.
// PV keeps A = P in tensor memory and splits V along head_dim (64 columns per CTA).
//
// Work item = (head, 512 query rows): CTA r of the pair owns rows [256 r, 256 r + 256) as two 128-row tiles; MMA
// "tile i" spans tile i of both CTAs (M = 256).  Roles per CTA as in gf_attn80.cu (16 softmax warps, a TMA producer
// that stages this CTA's half of every K / V block); the two MMA issuers run in the leader CTA only and multicast
// their completion barriers to both CTAs; the softmax warps of the second CTA arrive on the leader's barriers through
// the cluster address space.  Same arithmetic, same rounding points, same results as gf_attn80.cu.
#include <type_traits>
#include "gf_attn_common.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int X2_D = 128;
constexpr int X2_BM = 128;
constexpr int X2_BN = 80;
constexpr int X2_HC = X2_BN / 2;
constexpr int X2_THREADS = 640;
constexpr int X2_SOFTMAX_REGS = 104;
constexpr int X2_SERVICE_REGS = 64;
constexpr int X2_SLOTS = 12;                     // ring slots of 10 KB: this CTA's half of a K or V block
constexpr int X2_Q_BYTES = X2_BM * X2_D * 2;     // 32 KB per tile
constexpr int X2_QHALF = X2_BM * 64 * 2;         // 16 KB
constexpr int X2_SLOT_BYTES = X2_BN * X2_D;      // 10 KB: K half = two [40][64] boxes, V half = one [80][64] box
constexpr int X2_KBOX = (X2_BN / 2) * 64 * 2;    // 5 KB
constexpr int X2_XCHG_BYTES = 2 * 2 * 2 * X2_BM * 4;
constexpr int X2_SMEM_BYTES = 2 * X2_Q_BYTES + X2_SLOTS * X2_SLOT_BYTES + X2_XCHG_BYTES + 1024 + 512;
constexpr float X2_RESCALE_THRESHOLD = 8.0f;

struct Attn80x2Params {
  AttnOut out;
  int Lq, Lk, heads;
  int q_blocks;          // ceil(Lq / 512): work items per head (one per CTA pair)
  int n_full;            // pairs [0, n_full) take a whole 512-row item; pairs beyond take one 256-row half of a tail item
  float scale_log2;
};

template <int kEmuPairs>
__global__ void __launch_bounds__(X2_THREADS, 1)
gf_attn80x2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const Attn80x2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t kv_smem = smem_base + 2 * X2_Q_BYTES;
  const uint32_t xchg_smem = kv_smem + X2_SLOTS * X2_SLOT_BYTES;
  const uint32_t bar_base = xchg_smem + X2_XCHG_BYTES;
  const uint32_t q_full = bar_base;                                          // leader's copy is the one waited on
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };             // leader's copy
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + X2_SLOTS + s); }; // both CTAs (multicast commit)
  auto s_full = [&](int i) { return bar_base + 8u * (1 + 2 * X2_SLOTS + i); };   // both CTAs (multicast commit)
  auto s_free = [&](int i) { return bar_base + 8u * (3 + 2 * X2_SLOTS + i); };   // leader's copy, 16 arrivals
  auto p_full = [&](int i) { return bar_base + 8u * (5 + 2 * X2_SLOTS + i); };   // leader's copy, 16 arrivals
  auto p_free = [&](int i) { return bar_base + 8u * (7 + 2 * X2_SLOTS + i); };   // both CTAs (multicast commit)
  const uint32_t tmem_ptr_smem = bar_base + 8u * (9 + 2 * X2_SLOTS);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader_cta = cta_rank == 0;
  // Work items are (head, 512 query rows) per CTA pair.  As in gf_attn80.cu, the items of a mostly empty last wave are
  // split in two (`single`): such a pair takes 256 rows, one 128-row tile per CTA; tile 1's warps and issuer idle.
  const int pair = (int)blockIdx.x >> 1;
  const bool single = pair >= p.n_full;
  const int tail_idx = single ? pair - p.n_full : 0;
  const int item = single ? p.n_full + (tail_idx >> 1) : pair;
  const int head = item / p.q_blocks;
  const int qb = item % p.q_blocks;
  const int n_tiles = single ? 1 : 2;
  const int q0 = qb * 4 * X2_BM + (single ? (tail_idx & 1) * 2 * X2_BM + (int)cta_rank * X2_BM
                                          : (int)cta_rank * 2 * X2_BM);     // first query row of this CTA
  const int n_kv = (p.Lk + X2_BN - 1) / X2_BN;
  const int col0 = head * X2_D;

  if (warp == 16 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 17) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < X2_SLOTS; ++s) {
        mbar_init(kv_full(s), 1);
        mbar_init(kv_empty(s), single ? 1 : 2);   // every active MMA issuer releases a slot (multicast to both CTAs)
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(s_full(i), 1);
        mbar_init(s_free(i), 16);    // one arrive per softmax warp of the tile, both CTAs
        mbar_init(p_full(i), 16);
        mbar_init(p_free(i), 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<2>(tmem_ptr_smem, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  auto tmem_S = [&](int i) { return tmem_base + uint32_t(i) * 128u; };
  auto tmem_P = [&](int i) { return tmem_base + uint32_t(i) * 128u + uint32_t(X2_BN); };
  auto tmem_O = [&](int i) { return tmem_base + 256u + uint32_t(i) * 128u; };

  if (warp >= 16) {
   setmaxnreg_dec<X2_SERVICE_REGS>();
   if (warp == 16) {
    // ===================================================== TMA producer: this CTA's Q tiles and its half of K / V
    // Every byte is accounted on the LEADER's barriers (q_full, kv_full): the leader's issuers start an MMA only when
    // both halves have landed.
    const bool lead = elect_one();
    const uint32_t q_full_l = mapa(q_full, 0);
    if (lead) {
      if (leader_cta) mbar_arrive_expect_tx(q_full, 2 * n_tiles * X2_Q_BYTES);
      for (int i = 0; i < n_tiles; ++i)
        for (int h = 0; h < 2; ++h)
          tma_load_2d_cg2(q_smem + i * X2_Q_BYTES + h * X2_QHALF, &tmQ, q_full_l, col0 + h * 64, q0 + i * X2_BM);
    }
    int slot = 0;
    uint32_t phase = 0;
    auto load_k = [&](int blk) {
      mbar_wait(kv_empty(slot), phase ^ 1u);
      if (lead) {
        if (leader_cta) mbar_arrive_expect_tx(kv_full(slot), 2 * X2_SLOT_BYTES);
        const uint32_t full_l = mapa(kv_full(slot), 0);
        for (int h = 0; h < 2; ++h)      // rows [blk*80 + 40 r, +40) of K, head columns in two 64-wide boxes
          tma_load_2d_cg2(kv_smem + slot * X2_SLOT_BYTES + h * X2_KBOX, &tmK, full_l, col0 + h * 64,
                          blk * X2_BN + (int)cta_rank * (X2_BN / 2));
      }
      __syncwarp();
      if (++slot == X2_SLOTS) { slot = 0; phase ^= 1u; }
    };
    auto load_v = [&](int blk) {
      mbar_wait(kv_empty(slot), phase ^ 1u);
      if (lead) {
        if (leader_cta) mbar_arrive_expect_tx(kv_full(slot), 2 * X2_SLOT_BYTES);
        const uint32_t full_l = mapa(kv_full(slot), 0);
        // all 80 rows of V, head columns [64 r, 64 r + 64): this CTA's half of the N = 128 operand
        tma_load_2d_cg2(kv_smem + slot * X2_SLOT_BYTES, &tmV, full_l, col0 + (int)cta_rank * 64, blk * X2_BN);
      }
      __syncwarp();
      if (++slot == X2_SLOTS) { slot = 0; phase ^= 1u; }
    };
    load_k(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) load_k(j + 1);
      load_v(j);
    }
   } else if ((warp == 17 || (warp == 18 && !single)) && leader_cta) {
    // ===================================================== MMA issuers (leader CTA): warp 17 tile 0, warp 18 tile 1
    const bool lead = elect_one();
    const int i = warp - 17;
    constexpr uint32_t idesc_qk = idesc_bf16(2 * X2_BM, X2_BN, 0, 0);   // M = 256, N = 80
    constexpr uint32_t idesc_pv = idesc_bf16(2 * X2_BM, X2_D, 0, 1);    // M = 256, N = 128, B MN-major
    constexpr uint64_t desc_k = smem_desc_base(/*sbo=*/1024, /*lbo=*/16);
    constexpr uint64_t desc_v = smem_desc_base(/*sbo=*/1024, /*lbo=*/X2_SLOT_BYTES);
    const uint32_t qa = q_smem + i * X2_Q_BYTES;
    const uint32_t tS = tmem_S(i), tP = tmem_P(i), tO = tmem_O(i);
    const uint32_t sfull = s_full(i), sfree = s_free(i), pfull = p_full(i), pfree = p_free(i);
    auto issue_qk = [&](uint32_t k_addr, uint32_t release_bar) {
      if (lead) {
#pragma unroll
        for (int kk = 0; kk < X2_D / 16; ++kk) {
          const uint32_t qoff = (kk >> 2) * X2_QHALF + (kk & 3) * 32;
          const uint32_t koff = (kk >> 2) * X2_KBOX + (kk & 3) * 32;
          umma_ss<2>(tS, smem_desc(desc_k, qa + qoff), smem_desc(desc_k, k_addr + koff), idesc_qk, kk != 0);
        }
        tc_commit_cg2_mc(sfull, 0x3);
        tc_commit_cg2_mc(release_bar, 0x3);
      }
      __syncwarp();
    };
    auto issue_pv = [&](uint32_t v_addr, bool first_block, uint32_t release_bar) {
      if (lead) {
#pragma unroll
        for (int kk = 0; kk < X2_BN / 16; ++kk)
          umma_ts<2>(tO, tP + kk * 8, smem_desc(desc_v, v_addr + kk * 2048), idesc_pv,
                     (first_block && kk == 0) ? 0u : 1u);
        tc_commit_cg2_mc(pfree, 0x3);
        tc_commit_cg2_mc(release_bar, 0x3);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    issue_qk(kv_smem, kv_empty(0));
    int slot = 1;
    uint32_t phase = 0;
    auto advance = [&]() { if (++slot == X2_SLOTS) { slot = 0; phase ^= 1u; } };
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) {
        mbar_wait(kv_full(slot), phase);
        mbar_wait(sfree, j & 1);
        tc_fence_after();
        issue_qk(kv_smem + slot * X2_SLOT_BYTES, kv_empty(slot));
        advance();
      }
      mbar_wait(kv_full(slot), phase);
      mbar_wait(pfull, j & 1);
      tc_fence_after();
      issue_pv(kv_smem + slot * X2_SLOT_BYTES, j == 0, kv_empty(slot));
      advance();
    }
   }
  } else if (!(single && warp >= 8)) {
    // ===================================================== softmax warpgroups (+ epilogue)
    setmaxnreg_inc<X2_SOFTMAX_REGS>();
    const int i = warp >> 3;                         // tile
    const int hf = (warp >> 2) & 1;                  // column half of the score row owned by this thread
    const int wq = warp & 3;                         // TMEM lane quarter
    const uint32_t lane = lane_id();
    const int r = wq * 32 + (int)lane;               // row inside the tile
    const uint32_t lane_off = uint32_t(wq * 32) << 16;
    const uint32_t tS = tmem_S(i) + lane_off + uint32_t(hf * X2_HC);
    const uint32_t tP = tmem_P(i) + lane_off + uint32_t(hf * X2_HC / 2);
    const uint32_t tO = tmem_O(i) + lane_off + uint32_t(hf * 64);
    const int row = q0 + i * X2_BM + r;
    const uint32_t s_free_l = mapa(s_free(i), 0), p_full_l = mapa(p_full(i), 0);   // the leader's barriers
    const int tail_valid = p.Lk - (n_kv - 1) * X2_BN - hf * X2_HC;   // valid columns of this half in the last block
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    const uint32_t x_mine = xchg_smem + uint32_t(((i * 2 + hf) * X2_BM + r) * 4);
    const uint32_t x_other = xchg_smem + uint32_t(((i * 2 + (hf ^ 1)) * X2_BM + r) * 4);
    const uint32_t tile_bar = 1 + i;                 // named barrier of the tile's 256 softmax threads
    float m_used = 0.f, l = 0.f;

    // Value held by the thread owning the other half of the row.  Slots alternate with `parity` so that a
    // thread's next write can never overtake its partner's read of the previous one.  (Measured alternatives that
    // lost on the same box: a 64-thread barrier per warp pair, -1 %; a barrier-free tagged-slot poll, -1.5 %.)
    auto exchange = [&](float mine, int parity) -> float {
      const uint32_t off = uint32_t(parity & 1) * (X2_XCHG_BYTES / 2);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(x_mine + off), "f"(mine) : "memory");
      named_bar_sync(tile_bar, 256);
      float other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(x_other + off) : "memory");
      return other;
    };

    auto kv_block = [&](const int j, auto first_tag) {
      constexpr bool kFirst = decltype(first_tag)::value;
      mbar_wait(s_full(i), j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[8];
      tmem_ld32(tS, s0);
      tmem_ld8(tS + 32, s1);
      tmem_ld_wait();
      // S(j) is in registers: the tensor core may overwrite it with S(j+1)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(s_free_l);
      if (j == n_kv - 1 && tail_valid < X2_HC) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k >= tail_valid) s0[k] = 0xFF800000u;          // -inf
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (32 + k >= tail_valid) s1[k] = 0xFF800000u;
      }
      const float hmax = fmaxf(cols_max<32>(s0), cols_max<8>(s1));
      if constexpr (kFirst) m_used = fmaxf(hmax, exchange(hmax, j)) * p.scale_log2;
      uint64_t acc[2] = {0ull, 0ull};
      uint32_t pk0[16], pk1[4];
      {
        const uint64_t negm2 = pack2(-m_used, -m_used);
        exp_cols<kEmuPairs, 32>(s0, scale2, negm2, acc, pk0);
        exp_cols<kEmuPairs, 8>(s1, scale2, negm2, acc, pk1);
      }
      if constexpr (!kFirst) {
        // true row max of this block (log2 domain); both threads of the row see the same value
        const float m_cur = fmaxf(hmax, exchange(hmax, j)) * p.scale_log2;
        // PV(j-1, i) must have drained P (and, for a rescale, O) before either is written
        mbar_wait(p_free(i), (j - 1) & 1);
        tc_fence_after();
        const bool need = m_cur > m_used + X2_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          // rare: move the reference max; this thread rescales its 64 columns of O and its partial row sum, and
          // recomputes its part of P
          const float alpha = need ? ex2_approx(m_used - m_cur) : 1.0f;
          if (need) m_used = m_cur;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[16];
            tmem_ld16(tO + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st16(tO + c * 16, o);
          }
          l *= alpha;
          acc[0] = 0ull; acc[1] = 0ull;
          const uint64_t negm2 = pack2(-m_used, -m_used);
          exp_cols<0, 32>(s0, scale2, negm2, acc, pk0);
          exp_cols<0, 8>(s1, scale2, negm2, acc, pk1);
        }
      }
      tmem_st16(tP, pk0);
      tmem_st4(tP + 16, pk1);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(p_full_l);
      float a0, a1, a2, a3;
      unpack2(acc[0], a0, a1);
      unpack2(acc[1], a2, a3);
      l += (a0 + a1) + (a2 + a3);
    };
    kv_block(0, std::true_type{});
#pragma unroll 1
    for (int j = 1; j < n_kv; ++j) kv_block(j, std::false_type{});

    // ---------------- epilogue: O / l -> bf16 -> global (this thread: 64 of the row's 128 columns)
    const float inv_l = 1.0f / (l + exchange(l, n_kv));
    mbar_wait(p_free(i), (n_kv - 1) & 1);
    tc_fence_after();
    __nv_bfloat16* orow;
    if (p.out.n_peers == 0) {
      orow = reinterpret_cast<__nv_bfloat16*>(p.out.base[0]) + (long long)row * p.out.ldo + col0 + hf * 64;
    } else {     // Ulysses return path: the row's owner receives it straight over NVLink
      const int owner = min(row / p.out.rows_per_peer, p.out.n_peers - 1);
      orow = reinterpret_cast<__nv_bfloat16*>(p.out.base[owner]) +
             (long long)(row - owner * p.out.rows_per_peer) * p.out.ldo + p.out.col_offset + col0 + hf * 64;
    }
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + c * 32, o);
      tmem_ld_wait();
      if (row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            w[k] = pack_bf16x2(__uint_as_float(o[g * 8 + 2 * k]) * inv_l, __uint_as_float(o[g * 8 + 2 * k + 1]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

template <int kEmuPairs>
static int launch80x2(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const Attn80x2Params& p,
                      cudaStream_t stream) {
  auto kern = gf_attn80x2_kernel<kEmuPairs>;
  static bool configured[64] = {};
  if (int rc = gf_set_smem_once(configured, reinterpret_cast<const void*>(kern), X2_SMEM_BYTES)) return rc;
  cudaLaunchConfig_t cfg{};
  const int items = p.q_blocks * p.heads;
  cfg.gridDim = dim3(2 * (p.n_full + 2 * (items - p.n_full)));
  cfg.blockDim = dim3(X2_THREADS);
  cfg.dynamicSmemBytes = X2_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, tmQ, tmK, tmV, p);
}

int gf_attention80x2_launch(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                            long long ldv, const AttnOut& out, int Lq, int Lk, int heads, float scale, int emu_pairs,
                            cudaStream_t stream) {
  CUtensorMap scr[3];
  int rc = 0;
  const CUtensorMap* tmQ = gf_ctx_tmap(ctx, &scr[0], Q, (uint64_t)heads * X2_D, (uint64_t)Lq, (uint64_t)ldq, 64, X2_BM, &rc);
  if (!tmQ) return rc;
  const CUtensorMap* tmK = gf_ctx_tmap(ctx, &scr[1], K, (uint64_t)heads * X2_D, (uint64_t)Lk, (uint64_t)ldk, 64, X2_BN / 2, &rc);
  if (!tmK) return rc;
  const CUtensorMap* tmV = gf_ctx_tmap(ctx, &scr[2], V, (uint64_t)heads * X2_D, (uint64_t)Lk, (uint64_t)ldv, 64, X2_BN, &rc);
  if (!tmV) return rc;
  Attn80x2Params p;
  p.out = out;
  p.Lq = Lq; p.Lk = Lk; p.heads = heads;
  p.q_blocks = (Lq + 4 * X2_BM - 1) / (4 * X2_BM);
  // tail splitting: if the items of the last partial wave fit on the SM pairs as half items, run them that way
  const int items = p.q_blocks * heads, pairs = gf_num_sms() / 2;
  const int tail = pairs > 0 ? items % pairs : 0;
  p.n_full = (tail > 0 && items > pairs && 2 * tail <= pairs) ? items - tail : items;
  p.scale_log2 = scale * 1.4426950408889634f;
  switch (emu_pairs) {
    case 0: return launch80x2<0>(*tmQ, *tmK, *tmV, p, stream);
    case 2: return launch80x2<2>(*tmQ, *tmK, *tmV, p, stream);
    case 6: return launch80x2<6>(*tmQ, *tmK, *tmV, p, stream);
    default: return launch80x2<4>(*tmQ, *tmK, *tmV, p, stream);
  }
}

}  // namespace gf
