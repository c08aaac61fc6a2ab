// gf_gemm.cu -- persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] . W[N,K]^T )        A, W, C bf16 row-major ("NT" GEMM == torch F.linear)
//
// Replaces every nn.Linear on the DiT hot path of the reference
//   (diffsynth/models/wan_video_dit.py:131-134,141-147,157-160,177-186,209-210,229 and the k=1 Conv1d
//    at src/goal_force/wan_video_new.py:1564-1570), with the elementwise tail fused into the epilogue:
//   EPI_BIAS        : C = acc + bias                                   (q/k/v projections, text/time MLPs)
//   EPI_BIAS_GELU   : C = gelu_tanh(acc + bias)                        (ffn.0 + nn.GELU(approximate='tanh'))
//   EPI_BIAS_SILU   : C = silu(acc + bias)                             (time_embedding.0 + SiLU)
//   EPI_F32         : C = acc + bias, stored as fp32 (ldc in floats)  (VAE attention scores, wan_video_vae.py:325-337)
//   EPI_GATE_RES    : C = R + gate[n] * (acc + bias)  (gate==null: 1)  (o-proj / ffn.2 + GateModule, cross-attn
//                                                                       residual, ControlNet zero-conv inject)
// Rounding points follow eager bf16 PyTorch (linear -> bf16, gate*y -> bf16, x+.. -> bf16) so the result tracks the
// reference's bf16 forward as closely as a different accumulation order allows.
//
// Structure (one CTA per SM, or one CTA pair per 2 SMs with cta_group::2):
//   warp 0      : TMA producer   (A tile 128x64, W tile 256x64 [or 128x64 per CTA in pair mode], 128B swizzle)
//   warp 1      : tcgen05.mma issuer (single elected thread; leader CTA only in pair mode), TMEM alloc/dealloc
//   warps 2..5  : epilogue: tcgen05.ld accumulator -> registers -> fused math -> 16-byte global stores
//   TMEM        : 2 accumulator stages x 256 fp32 columns (all 512 columns), so the epilogue of tile i overlaps the
//                 MMAs of tile i+1.
#include "gf_ptx.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int GEMM_BM = 128;       // rows per CTA
constexpr int GEMM_BN = 256;       // columns per tile (per CTA, or per CTA pair)
constexpr int GEMM_BK = 64;        // bf16 elements per k-block == one 128-byte swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_GROUP_M = 8;    // default rasterisation: tiles walk 8 m-tiles before moving along N (L2 reuse of W)

// kBN = columns per tile: 256, or 224 for shapes whose 256-wide tiling leaves the last wave of the 74 SM pairs mostly
// empty (M = 4095 rows per GPU at 8 GPUs with N = 5120: 320 tiles = 4.3 waves; 224-wide: 368 tiles = 4.97 waves).
template <int kCG, int kBN = GEMM_BN> struct GemmCfg {
  static constexpr int kBRows = kBN / kCG;                               // W rows staged by each CTA
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;                  // 16 KB
  static constexpr int kBBytes = kBRows * GEMM_BK * 2;                   // 32 KB / 16 KB (14 KB at kBN = 224)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kCG == 1) ? 4 : 6;                     // 192 KB of operand ring either way
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  __nv_bfloat16* C;
  long long ldc;
  const __nv_bfloat16* bias;      // [N] or null
  const __nv_bfloat16* gate;      // [N] or null (EPI_GATE_RES)
  const __nv_bfloat16* R;         // [M, ldr] residual (EPI_GATE_RES), may alias C
  long long ldr;
  int num_m_tiles, num_n_tiles;   // in units of (128*kCG) x 256
  int group_m;                    // rasterisation group height in m-tiles
};

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int group_m, int& m, int& n) {
  const int per_group = group_m * num_n;
  const int g = t / per_group;
  const int first_m = g * group_m;
  const int gsz = min(group_m, num_m - first_m);
  const int r = t - g * per_group;
  m = first_m + r % gsz;
  n = r / gsz;
}

template <int EPI> __device__ __forceinline__ float epi_act(float v) {
  if constexpr (EPI == GF_EPI_BIAS_GELU) {
    // nn.GELU(approximate='tanh'): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
    const float u = 0.7978845608028654f * (v + 0.044715f * v * v * v);
    return 0.5f * v * (1.0f + tanh_approx(u));
  } else if constexpr (EPI == GF_EPI_BIAS_SILU) {
    return v / (1.0f + __expf(-v));
  } else {
    return v;
  }
}

template <int kCG, int EPI, int kBN = GEMM_BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gf_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<kCG, kBN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
  // barrier layout: full[kStages] | empty[kStages] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * Cfg::kStages + 4);
  auto smem_a = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::kStageBytes + Cfg::kABytes; };

  const int warp = threadIdx.x >> 5;
  const uint32_t cta_rank = (kCG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int cluster_id = blockIdx.x / kCG;
  const int num_clusters = gridDim.x / kCG;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < Cfg::kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 4 * kCG);  // one arrive per epilogue warp of every CTA in the group
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<kCG>(tmem_ptr_smem, 512);
    tmem_relinquish<kCG>();
  }
  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mt, nt; tile_coords(t, p.num_m_tiles, p.num_n_tiles, p.group_m, mt, nt);
        const int m0 = (mt * kCG + (int)cta_rank) * GEMM_BM;
        const int n0 = nt * kBN + (int)cta_rank * Cfg::kBRows;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if constexpr (kCG == 1) {
            mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_2d(smem_a(stage), &tmA, full_bar(stage), kb * GEMM_BK, m0);
            tma_load_2d(smem_b(stage), &tmB, full_bar(stage), kb * GEMM_BK, n0);
          } else {
            // both CTAs stream their halves; all bytes are accounted on the leader's barrier
            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            const uint32_t lead_bar = mapa(full_bar(stage), 0);
            tma_load_2d_cg2(smem_a(stage), &tmA, lead_bar, kb * GEMM_BK, m0);
            tma_load_2d_cg2(smem_b(stage), &tmB, lead_bar, kb * GEMM_BK, n0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (leader && elect_one()) {
      constexpr uint32_t idesc = idesc_bf16(GEMM_BM * kCG, kBN, 0, 0);
      constexpr uint64_t dbase = smem_desc_base(/*sbo=*/1024, /*lbo=*/16);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_a(stage), b_addr = smem_b(stage);
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; ++k) {
            umma_ss<kCG>(d_tmem, smem_desc(dbase, a_addr + k * GEMM_UMMA_K * 2),
                         smem_desc(dbase, b_addr + k * GEMM_UMMA_K * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (kCG == 1) tc_commit(empty_bar(stage)); else tc_commit_cg2_mc(empty_bar(stage), 0x3);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (kCG == 1) tc_commit(tfull_bar(acc)); else tc_commit_cg2_mc(tfull_bar(acc), 0x3);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int q = warp & 3;                       // TMEM lane quarter this warp may touch
    const uint32_t lane = lane_id();
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int mt, nt; tile_coords(t, p.num_m_tiles, p.num_n_tiles, p.group_m, mt, nt);
      const int row = (mt * kCG + (int)cta_rank) * GEMM_BM + q * 32 + (int)lane;
      const int n0 = nt * kBN;
      const bool row_ok = row < p.M;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t(q * 32) << 16) + acc * GEMM_BN;
      __nv_bfloat16* crow = (EPI == GF_EPI_F32) ? p.C : p.C + (long long)row * p.ldc + n0;
      const __nv_bfloat16* rrow = (EPI == GF_EPI_GATE_RES) ? p.R + (long long)row * p.ldr + n0 : nullptr;
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(t_addr + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if constexpr (EPI == GF_EPI_F32) {
          // fp32 result (attention scores of the VAE's single-head attention, softmax'd by gf_softmax_f32_bf16)
          if (col0 < p.N && row_ok) {
            float* frow = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float4 o;
              o.x = __uint_as_float(v[g * 4 + 0]); o.y = __uint_as_float(v[g * 4 + 1]);
              o.z = __uint_as_float(v[g * 4 + 2]); o.w = __uint_as_float(v[g * 4 + 3]);
              if (p.bias) {
                o.x += __bfloat162float(p.bias[col0 + g * 4 + 0]); o.y += __bfloat162float(p.bias[col0 + g * 4 + 1]);
                o.z += __bfloat162float(p.bias[col0 + g * 4 + 2]); o.w += __bfloat162float(p.bias[col0 + g * 4 + 3]);
              }
              *reinterpret_cast<float4*>(frow + g * 4) = o;
            }
          }
        } else
        if (col0 < p.N) {                           // N is a multiple of 32 on every call site (checked on host)
          uint32_t outw[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {             // 8 columns per group == one 16-byte vector
            uint4 bv = make_uint4(0, 0, 0, 0), gv = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
            uint4 rv = make_uint4(0, 0, 0, 0);
            if (p.bias) bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + g * 8));
            if constexpr (EPI == GF_EPI_GATE_RES) {
              if (p.gate) gv = __ldg(reinterpret_cast<const uint4*>(p.gate + col0 + g * 8));
              if (row_ok) rv = *reinterpret_cast<const uint4*>(rrow + c * 32 + g * 8);
            }
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
            const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float x0 = __uint_as_float(v[g * 8 + 2 * j]) + bf16_lo(bw[j]);
              float x1 = __uint_as_float(v[g * 8 + 2 * j + 1]) + bf16_hi(bw[j]);
              if constexpr (EPI == GF_EPI_GATE_RES) {
                x0 = round_bf16(round_bf16(x0) * bf16_lo(gw[j]));
                x1 = round_bf16(round_bf16(x1) * bf16_hi(gw[j]));
                x0 += bf16_lo(rw[j]);
                x1 += bf16_hi(rw[j]);
              } else if constexpr (EPI != GF_EPI_BIAS) {
                x0 = epi_act<EPI>(round_bf16(x0));
                x1 = epi_act<EPI>(round_bf16(x1));
              }
              outw[g * 4 + j] = pack_bf16x2(x0, x1);
            }
          }
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(crow + c * 32 + g * 8) =
                  make_uint4(outw[g * 4], outw[g * 4 + 1], outw[g * 4 + 2], outw[g * 4 + 3]);
          }
        }
      }
      // accumulator stage drained: hand it back to the MMA issuer (in pair mode: the leader's barrier)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCG == 1 || leader) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cluster(mapa(tempty_bar(acc), 0));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  // ------------------------------------------------------- teardown
  tc_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
template <int kCG, int EPI, int kBN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<kCG, kBN>;
  auto kern = gf_gemm_kernel<kCG, EPI, kBN>;
  static bool configured[64] = {};
  if (int rc = gf_set_smem_once(configured, reinterpret_cast<const void*>(kern), Cfg::kSmemBytes)) return rc;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  int clusters = gf_num_sms() / kCG;
  if (clusters > tiles) clusters = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * kCG);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
}

template <int kCG, int kBN>
static int dispatch_epi(int epi, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, cudaStream_t s) {
  switch (epi) {
    case GF_EPI_BIAS: return launch_gemm<kCG, GF_EPI_BIAS, kBN>(a, b, p, s);
    case GF_EPI_BIAS_GELU: return launch_gemm<kCG, GF_EPI_BIAS_GELU, kBN>(a, b, p, s);
    case GF_EPI_BIAS_SILU: return launch_gemm<kCG, GF_EPI_BIAS_SILU, kBN>(a, b, p, s);
    case GF_EPI_GATE_RES: return launch_gemm<kCG, GF_EPI_GATE_RES, kBN>(a, b, p, s);
    case GF_EPI_F32: return launch_gemm<kCG, GF_EPI_F32, kBN>(a, b, p, s);
    default: return GF_ERR_BAD_ARG;
  }
}

// Tile width for the pair GEMM.  Tiles run in waves of one per SM pair; a mostly empty last wave costs more than its
// share (measured under the sustained power cap, tools/gemm_tile_bench.py, 256- vs 224-wide, alternating samples):
//   M = 4095 (8-way sequence parallel), N = 5120: 320 tiles = 4.3 waves vs 368 = 4.97: 224 is 22 % (K = 5120) and 26 %
//   (K = 13824) faster; M = 8190, N = 5120: 8.65 vs 9.95 waves: +4 %; N = 13824 / 15360 at M = 4095: 11.7 / 12.97 waves,
//   equal; M = 32760: 224 is 5 % slower (its smaller tile moves 134 B/clk through shared memory per MMA cycle, the
//   256-wide one exactly the 128 B/clk the SM has).  So: 224 only for short launches whose wave efficiency it lifts.
static int choose_bn(int M, int N, int pairs) {
  if (pairs <= 0) return GEMM_BN;
  const long long m_tiles = (M + 255) / 256;
  const long long t256 = m_tiles * ((N + 255) / 256), t224 = m_tiles * ((N + 223) / 224);
  const long long w256 = (t256 + pairs - 1) / pairs, w224 = (t224 + pairs - 1) / pairs;
  const double e256 = (double)M * N / ((double)w256 * pairs * 256.0 * 256.0);
  const double e224 = (double)M * N / ((double)w224 * pairs * 256.0 * 224.0);
  return (N > 256 && t256 >= pairs && w256 <= 12 && e224 > e256 + 0.02) ? 224 : GEMM_BN;
}

}  // namespace gf

extern "C" int gf_gemm_bf16(gf_ctx* ctx, const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M,
                            int N, int K, const void* bias, int epi, const void* gate, const void* R, long long ldr,
                            int cta_group, void* stream) {
  using namespace gf;
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0) return GF_ERR_BAD_ARG;
  if ((N % 32) || (K % 8) || (lda % 8) || (ldw % 8) || (ldc % (epi == GF_EPI_F32 ? 4 : 8))) return GF_ERR_BAD_ARG;
  if (epi == GF_EPI_GATE_RES && (!R || (ldr % 8))) return GF_ERR_BAD_ARG;
  if (cta_group != 1 && cta_group != 2) return GF_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) & 15)
    return GF_ERR_BAD_ARG;
  CUtensorMap scrA, scrB;
  int rc = 0;
  const CUtensorMap* tmA = gf_ctx_tmap(ctx, &scrA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, GEMM_BM, &rc);
  if (!tmA) return rc;
  const int forced_bn = gf_ctx_tuning(ctx).gemm_bn;
  const int bn = cta_group == 2 ? (forced_bn > 0 ? forced_bn : choose_bn(M, N, gf_num_sms() / 2)) : GEMM_BN;
  const CUtensorMap* tmB =
      gf_ctx_tmap(ctx, &scrB, W, (uint64_t)K, (uint64_t)N, (uint64_t)ldw, GEMM_BK, bn / cta_group, &rc);
  if (!tmB) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.C = reinterpret_cast<__nv_bfloat16*>(C); p.ldc = ldc;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.gate = reinterpret_cast<const __nv_bfloat16*>(gate);
  p.R = reinterpret_cast<const __nv_bfloat16*>(R); p.ldr = ldr;
  p.num_m_tiles = (M + GEMM_BM * cta_group - 1) / (GEMM_BM * cta_group);
  p.num_n_tiles = (N + bn - 1) / bn;
  // Rasterisation: tiles walk `group_m` m-tiles before moving along N.  Measured under sustained load at M = 32760
  // (tools/gpu_check.py gemm_sustained): 16 is best for K = 5120 (each operand slab of a tile is 2.6 MB), 8 for
  // K = 13824 (7 MB slabs: a taller group no longer fits L2 next to the W columns in flight); 4 and 32 lose 6-10 %.
  const int forced = gf_ctx_tuning(ctx).gemm_group_m;
  p.group_m = forced > 0 ? forced : (K <= 8192 ? 2 * GEMM_GROUP_M : GEMM_GROUP_M);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cta_group == 1) return dispatch_epi<1, GEMM_BN>(epi, *tmA, *tmB, p, s);
  return bn == 224 ? dispatch_epi<2, 224>(epi, *tmA, *tmB, p, s) : dispatch_epi<2, GEMM_BN>(epi, *tmA, *tmB, p, s);
}
