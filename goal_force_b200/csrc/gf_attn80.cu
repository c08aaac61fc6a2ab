// gf_attn80.cu -- tcgen05 flash attention, "decoupled" variant: 80-row kv blocks, S and P in separate TMEM columns,
// four softmax warpgroups.
//
// Same contract as gf_attn.cu (O = softmax(Q K^T * scale) V per head, head_dim 128; wan_video_dit.py:28-61) and the
// same CTA tile: 256 query rows of one head as two 128-row tiles.
//
// Why 80: TMEM has 512 columns; the two O accumulators take 256, leaving 128 per tile.  With 128-column kv blocks P
// (bf16, 64 columns) must alias S (fp32, 128 columns), so QK(j+1) cannot be issued before PV(j) has consumed P(j),
// and every block pays the chain  softmax -> PV -> QK -> softmax  including two mbarrier hand-offs (gf_attn.cu:
// softmax warps wait ~45 % of their time for S).  With 80-column kv blocks S (80 columns) and P (40 columns) fit side
// by side in the 128 columns, so
//   * QK(j+1, i) is issued as soon as the softmax warps have pulled S(j, i) into registers  (barrier s_free),
//   * PV(j, i) is issued whenever P(j, i) is complete                                        (barrier p_full),
//   * the softmax of block j+1 only needs PV(j) to have drained P before it STORES the new P (barrier p_free);
//     it exponentiates into registers meanwhile.
// The softmax and the MMAs then run as two decoupled streams and the kernel is bound by the slower of the two.
//
// Why four softmax warpgroups: with the streams decoupled the limiter is the softmax warps' own serial time per
// block -- about 770 cycles of fixed latency (barrier wake-up, tcgen05.ld, P store, fences) around the
// MUFU-bound exponentials.  Two threads per query row (40 columns each) put four softmax warps on every SMSP, so
// one warp's exponentials fill the MUFU while the others sit in their fixed-latency phases.  The two threads of a
// row share the reference max: per block they exchange their half-row maxima through shared memory (one 256-thread
// named barrier per tile) so that both take the same (rare) rescale decision; each keeps its own partial row sum.
//
// Warp roles (640 threads): warps 0-3 / 4-7: tile 0, columns 0-39 / 40-79;  warps 8-11 / 12-15: tile 1;
//   warp 16: TMA producer;  warp 17 / 18: MMA issuer of tile 0 / 1 (one elected thread each: per block QK(j+1, i)
//   then PV(j, i); one thread for both tiles would itself be the bottleneck: every satisfied mbarrier wait costs
//   ~90 cycles and every tcgen05.mma / commit a few tens);  warp 19: idle.
// TMEM columns:  tile i: S = [128 i, 128 i + 80), P = [128 i + 80, 128 i + 120);  O_i = [256 + 128 i, +128).
#include <type_traits>
#include "gf_attn_common.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int A8_D = 128;                       // head dim
constexpr int A8_BM = 128;                      // query rows per tile (2 tiles per CTA)
constexpr int A8_BN = 80;                       // kv rows per block
constexpr int A8_HC = A8_BN / 2;                // score columns per softmax thread
constexpr int A8_THREADS = 640;                 // 16 softmax warps + 4 service warps
constexpr int A8_SOFTMAX_REGS = 104;            // 512*104 + 128*64 = 61440 = 640*96 (the pool at launch)
constexpr int A8_SERVICE_REGS = 64;
constexpr int A8_SLOTS = 6;                     // K/V ring slots
constexpr int A8_Q_BYTES = A8_BM * A8_D * 2;    // 32 KB per tile: two [128][64] 128B-swizzled boxes
constexpr int A8_QHALF = A8_BM * 64 * 2;        // 16 KB
constexpr int A8_SLOT_BYTES = A8_BN * A8_D * 2; // 20 KB: two [80][64] boxes
constexpr int A8_HALF = A8_BN * 64 * 2;         // 10 KB
constexpr int A8_XCHG_BYTES = 2 * 2 * 2 * A8_BM * 4;  // [parity][tile][half][row] fp32 exchange slots
constexpr int A8_SMEM_BYTES = 2 * A8_Q_BYTES + A8_SLOTS * A8_SLOT_BYTES + A8_XCHG_BYTES + 1024 + 256;
constexpr float A8_RESCALE_THRESHOLD = 8.0f;    // log2 domain

// -DGF_A8_TRACE (experimental builds only, never the shipped library): CTA 0 records SM-clock timestamps of the
// softmax / issuer phases of kv blocks [A8_TRACE_J0, A8_TRACE_J0 + A8_TRACE_NJ) into a global buffer set with
// gf_debug_attn_trace(); layout [warp 0..19][block][8] of int64.
#ifdef GF_A8_TRACE
constexpr int A8_TRACE_J0 = 100, A8_TRACE_NJ = 32;
static long long* g_a8_trace = nullptr;
#define A8_TS(slot)                                                                                      \
  do {                                                                                                   \
    if (p.trace && blockIdx.x == 0 && lane_id() == 0 && j >= A8_TRACE_J0 && j < A8_TRACE_J0 + A8_TRACE_NJ) \
      p.trace[((threadIdx.x >> 5) * A8_TRACE_NJ + (j - A8_TRACE_J0)) * 8 + (slot)] = clock64();           \
  } while (0)
#else
#define A8_TS(slot) do { } while (0)
#endif

struct Attn80Params {
  long long* trace;
  AttnOut out;
  int Lq, Lk, heads;
  int q_blocks;          // ceil(Lq / 256): two-tile work items per head
  int n_full;            // CTAs [0, n_full) take a whole two-tile item; CTAs beyond take ONE tile of a tail item each
  float scale_log2;      // softmax scale * log2(e)
};

template <int kEmuPairs>
__global__ void __launch_bounds__(A8_THREADS, 1)
gf_attn80_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const Attn80Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t kv_smem = smem_base + 2 * A8_Q_BYTES;
  const uint32_t xchg_smem = kv_smem + A8_SLOTS * A8_SLOT_BYTES;
  const uint32_t bar_base = xchg_smem + A8_XCHG_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + A8_SLOTS + s); };
  auto s_full = [&](int i) { return bar_base + 8u * (1 + 2 * A8_SLOTS + i); };
  auto s_free = [&](int i) { return bar_base + 8u * (3 + 2 * A8_SLOTS + i); };
  auto p_full = [&](int i) { return bar_base + 8u * (5 + 2 * A8_SLOTS + i); };
  auto p_free = [&](int i) { return bar_base + 8u * (7 + 2 * A8_SLOTS + i); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (9 + 2 * A8_SLOTS);

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from it
  // (barrier addresses, TMEM addresses, MMA descriptors of the issuer warps) in uniform registers -- without it the
  // issuers spend ~300 cycles per QK on R2UR moves and vector-register descriptor arithmetic
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  // Work items are (head, 256 query rows).  When the last, partial wave of items would leave most SMs idle, the host
  // splits those tail items into single-tile CTAs (`single`): tile 1's warps and issuer stay passive, and the tail
  // wave runs on twice as many SMs at roughly 0.6x the per-block time.
  const bool single = (int)blockIdx.x >= p.n_full;
  const int tail_idx = single ? (int)blockIdx.x - p.n_full : 0;
  const int item = single ? p.n_full + (tail_idx >> 1) : (int)blockIdx.x;
  const int head = item / p.q_blocks;                // consecutive CTAs share a head's K/V in L2
  const int qb = item % p.q_blocks;
  const int q0 = qb * 2 * A8_BM + (tail_idx & 1) * A8_BM;
  const int n_tiles = single ? 1 : 2;
  const int n_kv = (p.Lk + A8_BN - 1) / A8_BN;
  const int col0 = head * A8_D;

  if (warp == 16 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 17) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < A8_SLOTS; ++s) {
        mbar_init(kv_full(s), 1);
        mbar_init(kv_empty(s), 2);   // both MMA issuers release a slot
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(s_full(i), 1);
        mbar_init(s_free(i), 8);     // one arrive per softmax warp of the tile
        mbar_init(p_full(i), 8);
        mbar_init(p_free(i), 1);
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
  auto tmem_P = [&](int i) { return tmem_base + uint32_t(i) * 128u + uint32_t(A8_BN); };
  auto tmem_O = [&](int i) { return tmem_base + 256u + uint32_t(i) * 128u; };

  if (warp >= 16) {
   setmaxnreg_dec<A8_SERVICE_REGS>();
   if (warp == 16) {
    // ===================================================== TMA producer: Q, K0, then K(j+1), V(j) for every block
    // (warp-wide loop for the same reason as the issuers; the elected lane issues the copies)
    const bool lead = elect_one();
    if (lead) {
      mbar_arrive_expect_tx(q_full, n_tiles * A8_Q_BYTES);
      for (int i = 0; i < n_tiles; ++i)
        for (int h = 0; h < 2; ++h)
          tma_load_2d(q_smem + i * A8_Q_BYTES + h * A8_QHALF, &tmQ, q_full, col0 + h * 64, q0 + i * A8_BM);
    }
    int slot = 0;
    uint32_t phase = 0;
    auto load = [&](const CUtensorMap* tm, int blk) {
      mbar_wait(kv_empty(slot), phase ^ 1u);
      if (lead) {
        mbar_arrive_expect_tx(kv_full(slot), A8_SLOT_BYTES);
        for (int h = 0; h < 2; ++h)
          tma_load_2d(kv_smem + slot * A8_SLOT_BYTES + h * A8_HALF, tm, kv_full(slot), col0 + h * 64, blk * A8_BN);
      }
      __syncwarp();
      if (++slot == A8_SLOTS) { slot = 0; phase ^= 1u; }
    };
    load(&tmK, 0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) load(&tmK, j + 1);
      load(&tmV, j);
    }
   } else if (warp == 17 || warp == 18) {
    // ===================================================== MMA issuers: warp 17 drives tile 0, warp 18 tile 1
    // The whole warp runs the loop (waits, ring bookkeeping) so that slot / phase / descriptors are warp-uniform and
    // live in uniform registers; only the tcgen05 instructions themselves are issued by the elected lane.
    const bool lead = elect_one();
    if (warp == 18 && single) {
      // passive tile: only keep the K/V ring turning (the slots are released by two arrivals)
      int slot = 0;
      uint32_t phase = 0;
      for (int t = 0; t < 2 * n_kv; ++t) {
        mbar_wait(kv_full(slot), phase);
        if (lead) mbar_arrive(kv_empty(slot));
        if (++slot == A8_SLOTS) { slot = 0; phase ^= 1u; }
      }
    } else {
      const int i = warp - 17;
      constexpr uint32_t idesc_qk = idesc_bf16(A8_BM, A8_BN, 0, 0);   // A = Q (K-major), B = K (K-major), N = 80
      constexpr uint32_t idesc_pv = idesc_bf16(A8_BM, A8_D, 0, 1);    // A = P (TMEM),    B = V (MN-major), N = 128
      constexpr uint64_t desc_k = smem_desc_base(/*sbo=*/1024, /*lbo=*/16);
      constexpr uint64_t desc_v = smem_desc_base(/*sbo=*/1024, /*lbo=*/A8_HALF);
      const uint32_t qa = q_smem + i * A8_Q_BYTES;
      const uint32_t tS = tmem_S(i), tP = tmem_P(i), tO = tmem_O(i);
      const uint32_t sfull = s_full(i), sfree = s_free(i), pfull = p_full(i), pfree = p_free(i);
      auto issue_qk = [&](uint32_t k_addr, uint32_t release_bar) {
        if (lead) {
#pragma unroll
          for (int kk = 0; kk < A8_D / 16; ++kk) {
            const uint32_t qoff = (kk >> 2) * A8_QHALF + (kk & 3) * 32;
            const uint32_t koff = (kk >> 2) * A8_HALF + (kk & 3) * 32;
            umma_ss<1>(tS, smem_desc(desc_k, qa + qoff), smem_desc(desc_k, k_addr + koff), idesc_qk, kk != 0);
          }
          tc_commit(sfull);
          tc_commit(release_bar);
        }
        __syncwarp();
      };
      auto issue_pv = [&](uint32_t v_addr, bool first_block, uint32_t release_bar) {
        if (lead) {
#pragma unroll
          for (int kk = 0; kk < A8_BN / 16; ++kk)    // 16 kv rows per MMA: 8 packed-bf16 TMEM columns of P, 2 KB of V
            umma_ts<1>(tO, tP + kk * 8, smem_desc(desc_v, v_addr + kk * 2048), idesc_pv,
                       (first_block && kk == 0) ? 0u : 1u);
          tc_commit(pfree);
          tc_commit(release_bar);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(kv_full(0), 0);
      tc_fence_after();
      issue_qk(kv_smem, kv_empty(0));
      int slot = 1;                   // ring position of the next load (same sequence as the producer)
      uint32_t phase = 0;
      auto advance = [&]() { if (++slot == A8_SLOTS) { slot = 0; phase ^= 1u; } };
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          mbar_wait(kv_full(slot), phase);
          A8_TS(0);
          mbar_wait(sfree, j & 1);
          A8_TS(1);
          tc_fence_after();
          issue_qk(kv_smem + slot * A8_SLOT_BYTES, kv_empty(slot));
          A8_TS(2);
          advance();
        }
        mbar_wait(kv_full(slot), phase);
        A8_TS(3);
        mbar_wait(pfull, j & 1);
        A8_TS(4);
        tc_fence_after();
        issue_pv(kv_smem + slot * A8_SLOT_BYTES, j == 0, kv_empty(slot));
        A8_TS(5);
        advance();
      }
    }
   }
  } else if (!(single && warp >= 8)) {
    // ===================================================== softmax warpgroups (+ epilogue)
    setmaxnreg_inc<A8_SOFTMAX_REGS>();
    const int i = warp >> 3;                         // tile
    const int hf = (warp >> 2) & 1;                  // column half of the score row owned by this thread
    const int wq = warp & 3;                         // TMEM lane quarter
    const uint32_t lane = lane_id();
    const int r = wq * 32 + (int)lane;               // row inside the tile
    const uint32_t lane_off = uint32_t(wq * 32) << 16;
    const uint32_t tS = tmem_S(i) + lane_off + uint32_t(hf * A8_HC);
    const uint32_t tP = tmem_P(i) + lane_off + uint32_t(hf * A8_HC / 2);
    const uint32_t tO = tmem_O(i) + lane_off + uint32_t(hf * 64);
    const int row = q0 + i * A8_BM + r;
    const int tail_valid = p.Lk - (n_kv - 1) * A8_BN - hf * A8_HC;   // valid columns of this half in the last block
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    const uint32_t x_mine = xchg_smem + uint32_t(((i * 2 + hf) * A8_BM + r) * 4);
    const uint32_t x_other = xchg_smem + uint32_t(((i * 2 + (hf ^ 1)) * A8_BM + r) * 4);
    const uint32_t tile_bar = 1 + i;                 // named barrier of the tile's 256 softmax threads
    float m_used = 0.f, l = 0.f;

    // Value held by the thread owning the other half of the row.  Slots alternate with `parity` so that a
    // thread's next write can never overtake its partner's read of the previous one.  (Measured alternatives that
    // lost on the same box: a 64-thread barrier per warp pair, -1 %; a barrier-free tagged-slot poll, -1.5 %.)
    auto exchange = [&](float mine, int parity) -> float {
      const uint32_t off = uint32_t(parity & 1) * (A8_XCHG_BYTES / 2);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(x_mine + off), "f"(mine) : "memory");
      named_bar_sync(tile_bar, 256);
      float other;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(x_other + off) : "memory");
      return other;
    };

    auto kv_block = [&](const int j, auto first_tag) {
      constexpr bool kFirst = decltype(first_tag)::value;
      A8_TS(0);
      mbar_wait(s_full(i), j & 1);
      A8_TS(1);
      tc_fence_after();
      uint32_t s0[32], s1[8];
      tmem_ld32(tS, s0);
      tmem_ld8(tS + 32, s1);
      tmem_ld_wait();
      A8_TS(2);
      // S(j) is in registers: the tensor core may overwrite it with S(j+1)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free(i));
      if (j == n_kv - 1 && tail_valid < A8_HC) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k >= tail_valid) s0[k] = 0xFF800000u;          // -inf
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (32 + k >= tail_valid) s1[k] = 0xFF800000u;
      }
      const float hmax = fmaxf(cols_max<32>(s0), cols_max<8>(s1));
      A8_TS(3);
      if constexpr (kFirst) m_used = fmaxf(hmax, exchange(hmax, j)) * p.scale_log2;
      uint64_t acc[2] = {0ull, 0ull};
      uint32_t pk0[16], pk1[4];
      {
        const uint64_t negm2 = pack2(-m_used, -m_used);
        exp_cols<kEmuPairs, 32>(s0, scale2, negm2, acc, pk0);
        exp_cols<kEmuPairs, 8>(s1, scale2, negm2, acc, pk1);
      }
      A8_TS(4);
      if constexpr (!kFirst) {
        // true row max of this block (log2 domain); both threads of the row see the same value
        const float m_cur = fmaxf(hmax, exchange(hmax, j)) * p.scale_log2;
        A8_TS(5);
        // PV(j-1, i) must have drained P (and, for a rescale, O) before either is written
        mbar_wait(p_free(i), (j - 1) & 1);
        A8_TS(6);
        tc_fence_after();
        const bool need = m_cur > m_used + A8_RESCALE_THRESHOLD;
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
      if (lane == 0) mbar_arrive(p_full(i));
      A8_TS(7);
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
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int kEmuPairs>
static int launch80(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const Attn80Params& p,
                    cudaStream_t stream) {
  auto kern = gf_attn80_kernel<kEmuPairs>;
  static bool configured[64] = {};
  if (int rc = gf_set_smem_once(configured, reinterpret_cast<const void*>(kern), A8_SMEM_BYTES)) return rc;
  const int items = p.q_blocks * p.heads;
  kern<<<dim3(p.n_full + 2 * (items - p.n_full)), dim3(A8_THREADS), A8_SMEM_BYTES, stream>>>(tmQ, tmK, tmV, p);
  return (int)cudaGetLastError();
}

int gf_attention80_launch(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                          const AttnOut& out, int Lq, int Lk, int heads, float scale, int emu_pairs,
                          cudaStream_t stream) {
  CUtensorMap scr[3];
  int rc = 0;
  const CUtensorMap* tmQ = gf_ctx_tmap(ctx, &scr[0], Q, (uint64_t)heads * A8_D, (uint64_t)Lq, (uint64_t)ldq, 64, A8_BM, &rc);
  if (!tmQ) return rc;
  const CUtensorMap* tmK = gf_ctx_tmap(ctx, &scr[1], K, (uint64_t)heads * A8_D, (uint64_t)Lk, (uint64_t)ldk, 64, A8_BN, &rc);
  if (!tmK) return rc;
  const CUtensorMap* tmV = gf_ctx_tmap(ctx, &scr[2], V, (uint64_t)heads * A8_D, (uint64_t)Lk, (uint64_t)ldv, 64, A8_BN, &rc);
  if (!tmV) return rc;
  Attn80Params p;
#ifdef GF_A8_TRACE
  p.trace = g_a8_trace;
#else
  p.trace = nullptr;
#endif
  p.out = out;
  p.Lq = Lq; p.Lk = Lk; p.heads = heads;
  p.q_blocks = (Lq + 2 * A8_BM - 1) / (2 * A8_BM);
  // tail splitting: if the items of the last partial wave fit on the SMs as single-tile CTAs, run them that way
  const int items = p.q_blocks * heads, sms = gf_num_sms();
  const int tail = sms > 0 ? items % sms : 0;
  p.n_full = (tail > 0 && items > sms && 2 * tail <= sms) ? items - tail : items;
  p.scale_log2 = scale * 1.4426950408889634f;
  switch (emu_pairs) {
    case 0: return launch80<0>(*tmQ, *tmK, *tmV, p, stream);
    case 2: return launch80<2>(*tmQ, *tmK, *tmV, p, stream);
    case 6: return launch80<6>(*tmQ, *tmK, *tmV, p, stream);
    default: return launch80<4>(*tmQ, *tmK, *tmV, p, stream);
  }
}

}  // namespace gf

#ifdef GF_A8_TRACE
extern "C" int gf_debug_attn_trace(void* buf) {
  gf::g_a8_trace = reinterpret_cast<long long*>(buf);
  return 0;
}
#endif
