// gf_attn.cu -- tcgen05 flash attention for sm_100a (head_dim 128, no mask, no dropout).
//
//   O[:, h] = softmax(Q[:, h] K[:, h]^T * scale) V[:, h]      replaces flash_attention(), wan_video_dit.py:28-61
//
// One CTA owns 256 query rows of one head (two 128-row tiles that ping-pong through the tensor core):
//   warps 0-3 / 4-7 : softmax warpgroup of tile 0 / tile 1.  Thread t owns query row t of its tile: it reads the
//                     whole 128-column score row from TMEM (tcgen05.ld 32x32b, no shuffles needed for the row
//                     max / sum), exponentiates in the log2 domain (one FFMA + one MUFU.EX2 per score), packs
//                     P to bf16 and writes it back to TMEM over the score columns (tcgen05.st).
//   warp 8          : TMA producer: Q once, then K_0, V_0, K_1, V_1, ... through a 4-slot shared-memory ring.
//   warp 9          : MMA issuer (one elected thread):  S_i = Q_i K_j^T  (SS, both operands K-major in smem)
//                                                       O_i += P_i V_j   (TS, P from TMEM, V MN-major in smem)
//   TMEM (512 cols) : S0 | S1 | O0 | O1, 128 fp32 columns each; P_i aliases the first 64 columns of S_i.
// Issue order per KV block j:  PV(j,0) QK(j+1,0) PV(j,1) QK(j+1,1)  -- tcgen05.mma executes in issue order, so
// QK(j+1,i) cannot overwrite P_i(j) before PV(j,i) consumed it, and while warpgroup i runs its softmax the tensor
// core works on the other tile.
// Online softmax with lazy rescaling: the running reference max only moves when the block max exceeds it by more
// than 2^8; O is rescaled (by the softmax warp itself, straight in TMEM) only in that rare case.
#include "gf_ptx.cuh"
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
  __nv_bfloat16* O;
  long long ldo;
  int Lq, Lk, heads;
  int q_blocks;          // ceil(Lq / 256)
  float scale_log2;      // softmax scale * log2(e)
};

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
  auto p_full = [&](int i) { return bar_base + 8u * (3 + 2 * AT_SLOTS + i); };
  auto o_done = [&](int i) { return bar_base + 8u * (5 + 2 * AT_SLOTS + i); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (7 + 2 * AT_SLOTS);

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
        mbar_init(p_full(i), 128);
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
          mbar_wait(p_full(i), j & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < AT_BN / 16; ++kk)   // 16 kv rows per MMA: 8 packed-bf16 TMEM columns of P, 2 KB of V
            umma_ts(tmem_O(i), tmem_S(i) + kk * 8, smem_desc(desc_v, v_addr + kk * 2048), idesc_pv,
                    (j | kk) != 0 ? 1u : 0u);
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
    float m_used = 0.f, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
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
      float mx = __uint_as_float(s[0][0]);
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(s[c][k]));
      const float m_cur = mx * p.scale_log2;
      bool need = false;
      float alpha = 1.0f;
      if (j == 0) {
        m_used = m_cur;
      } else if (m_cur > m_used + AT_RESCALE_THRESHOLD) {
        need = true;
        alpha = ex2_approx(m_used - m_cur);
        m_used = m_cur;
      }
      if (__any_sync(0xffffffffu, need)) {
        // rare: bring O (and l) to the new reference max. PV(j-1, i) must have landed first.
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
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[c][2 * k]), p.scale_log2, -m_used));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[c][2 * k + 1]), p.scale_log2, -m_used));
          sum += p0 + p1;
          pk[k] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tS + c * 16, pk);                  // P_i: packed bf16, columns [0, 64) of the S_i region
      }
      l += sum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(i));
    }
    // ---------------- epilogue: O / l -> bf16 -> global
    mbar_wait(o_done(i), (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.O + (long long)row * p.ldo + col0;
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

}  // namespace gf

extern "C" int gf_attention_bf16(const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                                 long long ldv, void* O, long long ldo, int Lq, int Lk, int heads, int head_dim,
                                 float scale, void* stream) {
  using namespace gf;
  if (!Q || !K || !V || !O || Lq <= 0 || Lk <= 0 || heads <= 0) return GF_ERR_BAD_ARG;
  if (head_dim != AT_D) return GF_ERR_UNSUPPORTED;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || (reinterpret_cast<uintptr_t>(O) & 15)) return GF_ERR_BAD_ARG;
  CUtensorMap tmQ, tmK, tmV;
  int rc = gf_make_tmap_2d_bf16(&tmQ, Q, (uint64_t)heads * AT_D, (uint64_t)Lq, (uint64_t)ldq, 64, AT_BM);
  if (rc) return rc;
  rc = gf_make_tmap_2d_bf16(&tmK, K, (uint64_t)heads * AT_D, (uint64_t)Lk, (uint64_t)ldk, 64, AT_BN);
  if (rc) return rc;
  rc = gf_make_tmap_2d_bf16(&tmV, V, (uint64_t)heads * AT_D, (uint64_t)Lk, (uint64_t)ldv, 64, AT_BN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gf_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  AttnParams p;
  p.O = reinterpret_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.Lq = Lq; p.Lk = Lk; p.heads = heads;
  p.q_blocks = (Lq + 2 * AT_BM - 1) / (2 * AT_BM);
  p.scale_log2 = scale * 1.4426950408889634f;
  const dim3 grid(p.q_blocks * heads), block(AT_THREADS);
  gf_attn_kernel<<<grid, block, AT_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  return (int)cudaGetLastError();
}
