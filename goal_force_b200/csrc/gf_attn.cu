// gf_attn.cu -- tcgen05 flash attention for sm_100a (head_dim 128, no mask, no dropout).
//
//   O[:, h] = softmax(Q[:, h] K[:, h]^T * scale) V[:, h]      replaces flash_attention(), wan_video_dit.py:28-61
//
// One CTA owns 256 query rows of one head (two 128-row tiles that ping-pong through the tensor core):
//   warps 0-3 / 4-7 : softmax warpgroup of tile 0 / tile 1.  Thread t owns query row t of its tile: it reads the
//                     whole 128-column score row from TMEM (tcgen05.ld 32x32b, no shuffles needed for the row
//                     max / sum), exponentiates in the log2 domain, packs P to bf16 and writes it back to TMEM
//                     over the score columns (tcgen05.st).
//   warp 8          : TMA producer: Q once, then K_0, V_0, K_1, V_1, ... through a 4-slot shared-memory ring.
//   warp 9          : MMA issuer (one elected thread):  S_i = Q_i K_j^T  (SS, both operands K-major in smem)
//                                                       O_i += P_i V_j   (TS, P from TMEM, V MN-major in smem)
//   TMEM (512 cols) : S0 | S1 | O0 | O1, 128 fp32 columns each; P_i aliases the first 64 columns of S_i.
// Issue order per KV block j:  PV(j,0) QK(j+1,0) PV(j,1) QK(j+1,1)  -- tcgen05.mma executes in issue order, so
// QK(j+1,i) cannot overwrite P_i(j) before PV(j,i) consumed it, and while warpgroup i runs its softmax the tensor
// core works on the other tile.
//
// The per-tile chain softmax(j) -> PV(j) -> QK(j+1) -> softmax(j+1) is what bounds the kernel (each MMA pair is
// 1024 tensor cycles, and 128 MUFU.EX2 per thread are 1024 cycles of one SMSP's MUFU), so the softmax is built to be
// short:
//   * speculative reference max: the exponentials of block j use the running reference max m_used of the previous
//     blocks, so MUFU work starts as soon as S is in registers; the true row max of block j is computed in the
//     shadow of the first half. Only if it exceeds m_used by more than 2^8 (rare after the first block) is the
//     reference moved: O and l are rescaled in TMEM by the softmax warp itself and the first half is redone.
//   * the exponent and the row sum use packed f32x2 FMA/ADD (half the issue slots);
//   * a fixed fraction of the exponentials is evaluated on the FMA pipe instead of the MUFU
//     (Cody-Waite split + degree-3 minimax polynomial, 7.5e-5 relative error, far below the bf16 rounding of P);
//   * P is handed to the MMA warp in two halves, so PV(j,i) starts while the second half is still being
//     exponentiated.
#include <cstdlib>
#include <type_traits>
#include "gf_attn_common.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int AT_D = 128;                 // head dim
constexpr int AT_BM = 128;                // query rows per tile (2 tiles per CTA)
constexpr int AT_BN = 128;                // kv rows per block
constexpr int AT_THREADS = 384;           // 2 softmax warpgroups + 1 service warpgroup (producer, MMA, 2 idle warps)
constexpr int AT_SOFTMAX_REGS = 200;      // setmaxnreg budgets: 2*128*200 + 128*96 = 63488 <= 168*384 (the CTA pool)
constexpr int AT_SERVICE_REGS = 96;
constexpr int AT_SLOTS = 4;               // K/V ring slots, 32 KB each
constexpr int AT_TILE_BYTES = 128 * 128 * 2;   // 32 KB: two [128 rows][64 cols] 128B-swizzled boxes
constexpr int AT_HALF_BYTES = 128 * 64 * 2;    // 16 KB
constexpr int AT_SMEM_BYTES = 2 * AT_TILE_BYTES + AT_SLOTS * AT_TILE_BYTES + 1024 + 256;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;   // log2 domain

struct AttnParams {
  AttnOut out;
  int Lq, Lk, heads;
  int q_blocks;          // ceil(Lq / 256)
  float scale_log2;      // softmax scale * log2(e)
};

template <int kEmuPairs>
__global__ void __launch_bounds__(AT_THREADS, 1)
gf_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;                                 // 2 tiles
  const uint32_t kv_smem = smem_base + 2 * AT_TILE_BYTES;            // AT_SLOTS tiles
  const uint32_t bar_base = kv_smem + AT_SLOTS * AT_TILE_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + AT_SLOTS + s); };
  auto s_full = [&](int i) { return bar_base + 8u * (1 + 2 * AT_SLOTS + i); };
  auto p_half = [&](int i) { return bar_base + 8u * (3 + 2 * AT_SLOTS + i); };
  auto p_full = [&](int i) { return bar_base + 8u * (5 + 2 * AT_SLOTS + i); };
  auto o_done = [&](int i) { return bar_base + 8u * (7 + 2 * AT_SLOTS + i); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (9 + 2 * AT_SLOTS);

  const int warp = threadIdx.x >> 5;
  const int head = blockIdx.x / p.q_blocks;          // consecutive CTAs share a head's K/V in L2
  const int qb = blockIdx.x % p.q_blocks;
  const int q0 = qb * 2 * AT_BM;
  const int n_kv = (p.Lk + AT_BN - 1) / AT_BN;
  const int col0 = head * AT_D;

  if (warp == 8 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 9) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < AT_SLOTS; ++s) {
        mbar_init(kv_full(s), 1);
        mbar_init(kv_empty(s), 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(s_full(i), 1);
        mbar_init(p_half(i), 4);       // one arrive per softmax warp
        mbar_init(p_full(i), 4);
        mbar_init(o_done(i), 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_ptr_smem, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  auto tmem_S = [&](int i) { return tmem_base + uint32_t(i) * 128u; };
  auto tmem_O = [&](int i) { return tmem_base + 256u + uint32_t(i) * 128u; };

  if (warp >= 8) {
   setmaxnreg_dec<AT_SERVICE_REGS>();
   if (warp == 8) {
    // ===================================================== TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * AT_TILE_BYTES);
      for (int i = 0; i < 2; ++i)
        for (int h = 0; h < 2; ++h)
          tma_load_2d(q_smem + i * AT_TILE_BYTES + h * AT_HALF_BYTES, &tmQ, q_full, col0 + h * 64, q0 + i * AT_BM);
      for (int t = 0; t < 2 * n_kv; ++t) {
        const int slot = t % AT_SLOTS;
        mbar_wait(kv_empty(slot), ((t / AT_SLOTS) & 1) ^ 1);
        mbar_arrive_expect_tx(kv_full(slot), AT_TILE_BYTES);
        const CUtensorMap* tm = (t & 1) ? &tmV : &tmK;
        const int r0 = (t >> 1) * AT_BN;
        for (int h = 0; h < 2; ++h)
          tma_load_2d(kv_smem + slot * AT_TILE_BYTES + h * AT_HALF_BYTES, tm, kv_full(slot), col0 + h * 64, r0);
      }
    }
   } else if (warp == 9) {
    // ===================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_qk = idesc_bf16(AT_BM, AT_BN, 0, 0);   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = idesc_bf16(AT_BM, AT_D, 0, 1);    // A = P (TMEM),    B = V (MN-major)
      constexpr uint64_t desc_k = smem_desc_base(/*sbo=*/1024, /*lbo=*/16);
      constexpr uint64_t desc_v = smem_desc_base(/*sbo=*/1024, /*lbo=*/AT_HALF_BYTES);
      auto issue_qk = [&](int i, uint32_t k_addr) {
        const uint32_t qa = q_smem + i * AT_TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < AT_D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * AT_HALF_BYTES + (kk & 3) * 32;
          umma_ss<1>(tmem_S(i), smem_desc(desc_k, qa + off), smem_desc(desc_k, k_addr + off), idesc_qk, kk != 0);
        }
        tc_commit(s_full(i));
      };
      mbar_wait(q_full, 0);
      mbar_wait(kv_full(0), 0);
      tc_fence_after();
      issue_qk(0, kv_smem);
      issue_qk(1, kv_smem);
      tc_commit(kv_empty(0));
      for (int j = 0; j < n_kv; ++j) {
        const int tv = 2 * j + 1, tk = 2 * j + 2;
        const int slot_v = tv % AT_SLOTS, slot_k = tk % AT_SLOTS;
        const bool more = (j + 1 < n_kv);
        mbar_wait(kv_full(slot_v), (tv / AT_SLOTS) & 1);
        const uint32_t v_addr = kv_smem + slot_v * AT_TILE_BYTES;
        const uint32_t k_addr = kv_smem + slot_k * AT_TILE_BYTES;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          // 16 kv rows per MMA: 8 packed-bf16 TMEM columns of P, 2 KB of V; first half of P arrives early
          mbar_wait(p_half(i), j & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < AT_BN / 32; ++kk)
            umma_ts(tmem_O(i), tmem_S(i) + kk * 8, smem_desc(desc_v, v_addr + kk * 2048), idesc_pv,
                    (j | kk) != 0 ? 1u : 0u);
          mbar_wait(p_full(i), j & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = AT_BN / 32; kk < AT_BN / 16; ++kk)
            umma_ts(tmem_O(i), tmem_S(i) + kk * 8, smem_desc(desc_v, v_addr + kk * 2048), idesc_pv, 1u);
          tc_commit(o_done(i));
          if (more) {
            if (i == 0) {
              mbar_wait(kv_full(slot_k), (tk / AT_SLOTS) & 1);
              tc_fence_after();
            }
            issue_qk(i, k_addr);
          }
        }
        tc_commit(kv_empty(slot_v));
        if (more) tc_commit(kv_empty(slot_k));
      }
    }
   }
  } else {
    // ===================================================== softmax warpgroups (+ epilogue)
    setmaxnreg_inc<AT_SOFTMAX_REGS>();
    const int i = warp >> 2;                         // tile
    const int wq = warp & 3;                         // TMEM lane quarter
    const uint32_t lane = lane_id();
    const uint32_t lane_off = uint32_t(wq * 32) << 16;
    const uint32_t tS = tmem_S(i) + lane_off, tO = tmem_O(i) + lane_off;
    const int row = q0 + i * AT_BM + wq * 32 + (int)lane;
    const int tail_valid = p.Lk - (n_kv - 1) * AT_BN;          // valid columns of the last kv block (1..128)
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    float m_used = 0.f, l = 0.f;
    // One kv block.  kFirst (block 0) has no reference max yet, so it computes the row max before exponentiating;
    // every later block exponentiates against the running reference and checks the true max in the shadow.
    auto kv_block = [&](const int j, auto first_tag) {
      constexpr bool kFirst = decltype(first_tag)::value;
      mbar_wait(s_full(i), j & 1);
      tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tS + c * 32, s[c]);
      tmem_ld_wait();
      if (j == n_kv - 1 && tail_valid < AT_BN) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c * 32 + k >= tail_valid) s[c][k] = 0xFF800000u;  // -inf
      }
      if constexpr (kFirst)
        m_used = fmaxf(fmaxf(chunk_max(s[0]), chunk_max(s[1])), fmaxf(chunk_max(s[2]), chunk_max(s[3]))) *
                 p.scale_log2;
      uint64_t acc[2] = {0ull, 0ull};
      uint32_t pk[16];
      {
        const uint64_t negm2 = pack2(-m_used, -m_used);
        exp_chunk<kEmuPairs>(s[0], scale2, negm2, acc, pk);
        tmem_st16(tS, pk);
        exp_chunk<kEmuPairs>(s[1], scale2, negm2, acc, pk);
        tmem_st16(tS + 16, pk);
      }
      if constexpr (!kFirst) {
        // true row max of this block (log2 domain), computed in the shadow of the first half
        const float m_cur = fmaxf(fmaxf(chunk_max(s[0]), chunk_max(s[1])), fmaxf(chunk_max(s[2]), chunk_max(s[3]))) *
                            p.scale_log2;
        const bool need = m_cur > m_used + AT_RESCALE_THRESHOLD;
        if (__any_sync(0xffffffffu, need)) {
          // rare: move the reference max. PV(j-1, i) must have landed before O is touched; the first half of P
          // (computed against the stale reference) is redone.
          const float alpha = need ? ex2_approx(m_used - m_cur) : 1.0f;
          if (need) m_used = m_cur;
          mbar_wait(o_done(i), (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
            tmem_st32(tO + c * 32, o);
          }
          l *= alpha;
          acc[0] = 0ull; acc[1] = 0ull;
          const uint64_t negm2 = pack2(-m_used, -m_used);
          exp_chunk<0>(s[0], scale2, negm2, acc, pk);
          tmem_st16(tS, pk);
          exp_chunk<0>(s[1], scale2, negm2, acc, pk);
          tmem_st16(tS + 16, pk);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_half(i));
      {
        const uint64_t negm2 = pack2(-m_used, -m_used);
        exp_chunk<kEmuPairs>(s[2], scale2, negm2, acc, pk);
        tmem_st16(tS + 32, pk);
        exp_chunk<kEmuPairs>(s[3], scale2, negm2, acc, pk);
        tmem_st16(tS + 48, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(i));
      float a0, a1, a2, a3;
      unpack2(acc[0], a0, a1);
      unpack2(acc[1], a2, a3);
      l += (a0 + a1) + (a2 + a3);
    };
    kv_block(0, std::true_type{});
#pragma unroll 1
    for (int j = 1; j < n_kv; ++j) kv_block(j, std::false_type{});
    // ---------------- epilogue: O / l -> bf16 -> global
    mbar_wait(o_done(i), (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow;
    if (p.out.n_peers == 0) {
      orow = reinterpret_cast<__nv_bfloat16*>(p.out.base[0]) + (long long)row * p.out.ldo + col0;
    } else {     // Ulysses return path: the row's owner receives it straight over NVLink
      const int owner = min(row / p.out.rows_per_peer, p.out.n_peers - 1);
      orow = reinterpret_cast<__nv_bfloat16*>(p.out.base[owner]) +
             (long long)(row - owner * p.out.rows_per_peer) * p.out.ldo + p.out.col_offset + col0;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
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
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// Kernel selection (per context, gf_ctx_set_attention; no process-wide state):
//   impl      : 80 = gf_attn80.cu (decoupled, 80-row kv blocks, four softmax warpgroups),
//               128 = this file (128-row kv blocks, P aliases S), 0 = per shape
//   emu_pairs : column pairs per 16 whose exponential runs on the FMA pipe instead of the MUFU (0, 2, 4, 6);
//               -1 = kernel default (0 for impl 80, 4 for impl 128)
// Short key sequences (cross-attention against 512 context tokens) are a handful of kv blocks per CTA: there the
// 128-row-block kernel wastes fewer padded columns (512 = 4 x 128 vs 7 x 80) and measures ~8 % faster.
// Long key sequences take the CTA-pair 80-row-block kernel (160, gf_attn80x2.cu: QK no longer bound by the
// shared-memory reads of the tensor core; +5 % at 40 heads x 32,760 tokens, +3 % at 20 heads) once the launch is at
// least 8 waves of 512-row items over the 74 SM pairs; with fewer waves (5 heads per GPU at 8 GPUs: 4.3 waves) the
// coarser tail of the pair kernel eats the gain (measured 1264 vs 1254 TFLOP/s) and the single-CTA form (80) stays.
static int attn_impl_for(const CtxTuning& t, int Lq, int Lk, int heads) {
  if (t.attn_impl == 80 || t.attn_impl == 128 || t.attn_impl == 160) return t.attn_impl;
  if (Lk <= 1024) return 128;
  const long long items = (long long)((Lq + 511) / 512) * heads;
  return items >= 8LL * (gf_num_sms() / 2) ? 160 : 80;
}
static int attn_emu_for(const CtxTuning& t, int impl) { return t.attn_emu >= 0 ? t.attn_emu : (impl == 128 ? 4 : 0); }

template <int kEmuPairs>
static int launch_attn(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p,
                       cudaStream_t stream) {
  auto kern = gf_attn_kernel<kEmuPairs>;
  static bool configured[64] = {};
  if (int rc = gf_set_smem_once(configured, reinterpret_cast<const void*>(kern), AT_SMEM_BYTES)) return rc;
  const dim3 grid(p.q_blocks * p.heads), block(AT_THREADS);
  kern<<<grid, block, AT_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  return (int)cudaGetLastError();
}

}  // namespace gf

static int attention_dispatch(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                              const gf::AttnOut& out, int Lq, int Lk, int heads, int head_dim, float scale,
                              void* stream) {
  using namespace gf;
  if (!Q || !K || !V || Lq <= 0 || Lk <= 0 || heads <= 0) return GF_ERR_BAD_ARG;
  if (head_dim != AT_D) return GF_ERR_UNSUPPORTED;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (out.ldo % 8) || (out.col_offset % 8)) return GF_ERR_BAD_ARG;
  for (int i = 0; i < (out.n_peers ? out.n_peers : 1); ++i)
    if (!out.base[i] || (reinterpret_cast<uintptr_t>(out.base[i]) & 15)) return GF_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const CtxTuning tune = gf_ctx_tuning(ctx);
  const int impl = attn_impl_for(tune, Lq, Lk, heads);
  const int emu = attn_emu_for(tune, impl);
  if (impl == 80) return gf_attention80_launch(ctx, Q, ldq, K, ldk, V, ldv, out, Lq, Lk, heads, scale, emu, s);
  if (impl == 160) return gf_attention80x2_launch(ctx, Q, ldq, K, ldk, V, ldv, out, Lq, Lk, heads, scale, emu, s);
  CUtensorMap scr[3];
  int rc = 0;
  const CUtensorMap* tmQ = gf_ctx_tmap(ctx, &scr[0], Q, (uint64_t)heads * AT_D, (uint64_t)Lq, (uint64_t)ldq, 64, AT_BM, &rc);
  if (!tmQ) return rc;
  const CUtensorMap* tmK = gf_ctx_tmap(ctx, &scr[1], K, (uint64_t)heads * AT_D, (uint64_t)Lk, (uint64_t)ldk, 64, AT_BN, &rc);
  if (!tmK) return rc;
  const CUtensorMap* tmV = gf_ctx_tmap(ctx, &scr[2], V, (uint64_t)heads * AT_D, (uint64_t)Lk, (uint64_t)ldv, 64, AT_BN, &rc);
  if (!tmV) return rc;
  AttnParams p;
  p.out = out;
  p.Lq = Lq; p.Lk = Lk; p.heads = heads;
  p.q_blocks = (Lq + 2 * AT_BM - 1) / (2 * AT_BM);
  p.scale_log2 = scale * 1.4426950408889634f;
  switch (emu) {
    case 0: return launch_attn<0>(*tmQ, *tmK, *tmV, p, s);
    case 2: return launch_attn<2>(*tmQ, *tmK, *tmV, p, s);
    case 6: return launch_attn<6>(*tmQ, *tmK, *tmV, p, s);
    default: return launch_attn<4>(*tmQ, *tmK, *tmV, p, s);
  }
}

extern "C" int gf_attention_bf16(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                                 long long ldv, void* O, long long ldo, int Lq, int Lk, int heads, int head_dim,
                                 float scale, void* stream) {
  gf::AttnOut out{};
  out.base[0] = O;
  out.ldo = ldo;
  return attention_dispatch(ctx, Q, ldq, K, ldk, V, ldv, out, Lq, Lk, heads, head_dim, scale, stream);
}

extern "C" int gf_attention_scatter_bf16(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                                         long long ldv, void* const* O_peers, int n_peers, long long ldo,
                                         int rows_per_peer, int col_offset, int Lq, int Lk, int heads, int head_dim,
                                         float scale, void* stream) {
  if (!O_peers || n_peers < 1 || n_peers > GF_MAX_PEERS || rows_per_peer <= 0 || col_offset < 0) return GF_ERR_BAD_ARG;
  if ((long long)rows_per_peer * n_peers < Lq) return GF_ERR_BAD_ARG;
  gf::AttnOut out{};
  for (int i = 0; i < n_peers; ++i) out.base[i] = O_peers[i];
  out.ldo = ldo;
  out.n_peers = n_peers;
  out.rows_per_peer = rows_per_peer;
  out.col_offset = col_offset;
  return attention_dispatch(ctx, Q, ldq, K, ldk, V, ldv, out, Lq, Lk, heads, head_dim, scale, stream);
}
