// gf_rowwise.cu -- the HBM-bound kernels of the DiT block: adaLN LayerNorm and full-row RMSNorm + 3-D RoPE.
// A row (d bf16) is split over TPR threads (128 for d = 5120): every thread keeps NVT 128-bit vectors of the row
// in registers, so the row is read once (streaming 128-bit loads) and written once (128-bit stores); statistics are
// reduced with warp shuffles plus one shared-memory exchange between the warps of a row.  Few registers per thread
// (about 50) keep 40+ warps resident per SM, which is what hides HBM latency; a 256-thread CTA carries 256/TPR rows.
// Algorithmic traffic = 2 * rows * d * 2 bytes.
#include "gf_ptx.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int ROW_THREADS = 256;  // threads per CTA

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// Sum `v` over the TPR threads that share a row. `slot` is this reduction's private shared array
// [rows per CTA][warps per row]; one __syncthreads per call (all threads of the CTA must call it).
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* slot) {
  v = warp_sum(v);
  if constexpr (TPR == 32) {
    return v;
  } else {
    constexpr int WPR = TPR / 32;
    const int w = threadIdx.x >> 5, r = threadIdx.x / TPR;
    if ((threadIdx.x & 31) == 0) slot[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < WPR; ++i) t += slot[r * WPR + i];
    return t;
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// weight == nullptr: y = bf16(bf16(bf16(LN(x)) * bf16(1 + scale)) + shift)   (eager-PyTorch rounding chain of
//                    modulate(norm(x), shift, scale), wan_video_dit.py:64-65)
// weight != nullptr: y = bf16(LN(x) * weight + bias)                          (nn.LayerNorm with affine, fp32 math)
template <int NVT, int TPR>  // d = NVT * TPR * 8
__global__ void __launch_bounds__(ROW_THREADS)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                 int rows, float eps, const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
                 const __nv_bfloat16* __restrict__ weight, const __nv_bfloat16* __restrict__ bias) {
  constexpr int d = NVT * TPR * 8;
  constexpr int RPC = ROW_THREADS / TPR;
  __shared__ float red[2][ROW_THREADS / 32];
  const int t = threadIdx.x % TPR;
  const int row = blockIdx.x * RPC + threadIdx.x / TPR;
  const bool live = row < rows;                      // dead rows still take part in the CTA barriers
  const __nv_bfloat16* xr = x + (long long)(live ? row : 0) * ldx;
  uint4 v[NVT];
#pragma unroll
  for (int i = 0; i < NVT; ++i) v[i] = ld_stream(xr + (i * TPR + t) * 8);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NVT; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) s += bf16_lo(w[j]) + bf16_hi(w[j]);
  }
  const float mean = row_sum<TPR>(s, red[0]) * (1.0f / d);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVT; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(w[j]) - mean, b = bf16_hi(w[j]) - mean;
      ss += a * a + b * b;
    }
  }
  const float rstd = rsqrtf(row_sum<TPR>(ss, red[1]) * (1.0f / d) + eps);
  if (!live) return;
  __nv_bfloat16* yr = y + (long long)row * ldy;
#pragma unroll
  for (int i = 0; i < NVT; ++i) {
    const int col = (i * TPR + t) * 8;
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    uint32_t o[4];
    if (weight) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(weight + col));
      const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + col));
      const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = pack_bf16x2((bf16_lo(w[j]) - mean) * rstd * bf16_lo(ww[j]) + bf16_lo(bw[j]),
                           (bf16_hi(w[j]) - mean) * rstd * bf16_hi(ww[j]) + bf16_hi(bw[j]));
    } else {
      // bf16(bf16(LN) * bf16(1 + scale)) + shift with packed bf16x2 arithmetic: HADD2 / HMUL2 on bf16 compute in
      // fp32 and round once, which is exactly what eager PyTorch does for each of these bf16 tensor ops (the _rn forms
      // keep the compiler from contracting the multiply and the add into one fused, singly rounded HFMA2)
      const uint4 sc = __ldg(reinterpret_cast<const uint4*>(scale + col));
      const uint4 sh = __ldg(reinterpret_cast<const uint4*>(shift + col));
      const uint32_t cw[4] = {sc.x, sc.y, sc.z, sc.w}, hw[4] = {sh.x, sh.y, sh.z, sh.w};
      const __nv_bfloat162 one2 = __floats2bfloat162_rn(1.0f, 1.0f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 n2 = __floats2bfloat162_rn((bf16_lo(w[j]) - mean) * rstd, (bf16_hi(w[j]) - mean) * rstd);
        const __nv_bfloat162 s1p = __hadd2_rn(one2, *reinterpret_cast<const __nv_bfloat162*>(&cw[j]));
        const __nv_bfloat162 r2 = __hadd2_rn(__hmul2_rn(n2, s1p), *reinterpret_cast<const __nv_bfloat162*>(&hw[j]));
        o[j] = *reinterpret_cast<const uint32_t*>(&r2);
      }
    }
    *reinterpret_cast<uint4*>(yr + col) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// (a, b) = bf16( bf16(x * r) * weight ) for one packed pair, as fp32 values: the normalised pair is rounded to bf16
// once (F2FP), the multiply by the bf16 weight is one HMUL2.BF16 (fp32 product, one rounding) -- the two roundings of
// RMSNorm.forward (wan_video_dit.py:106-111) in eager PyTorch.
__device__ __forceinline__ void rms_scale_pair(uint32_t x2, float r, uint32_t w2, float& a, float& b) {
  const __nv_bfloat162 n2 = __floats2bfloat162_rn(bf16_lo(x2) * r, bf16_hi(x2) * r);
  const __nv_bfloat162 m2 = __hmul2_rn(n2, *reinterpret_cast<const __nv_bfloat162*>(&w2));
  const uint32_t m = *reinterpret_cast<const uint32_t*>(&m2);
  a = bf16_lo(m);
  b = bf16_hi(m);
}

// ---------------------------------------------------------------------------------------------- RMSNorm + RoPE
// x <- rope( bf16( bf16(x * rsqrt(mean(x^2) + eps)) * weight ) )       in place
// RoPE acts on interleaved pairs (2i, 2i+1) of every head: (a + ib) * (cos + i sin), one rounding to bf16
// (wan_video_dit.py:92-97 does the product in complex128; fp32 FMA on an fp32 table built from the float64
//  angles differs from that by < 1e-7 relative before the bf16 rounding).
// Thread t of a row always sees pairs 4t%64 .. 4t%64+3 of a head (head_dim 128, 8 elements per vector and TPR*8 a
// multiple of 128), so its 4 (cos, sin) pairs are loaded once per row.
// blockIdx.y selects a segment: segment g normalises columns [g*seg_stride, g*seg_stride + d) with weight[g]
// (q and k of a fused qkv row in one launch; each segment has its own row statistic).
template <int NVT, int TPR>
__global__ void __launch_bounds__(ROW_THREADS)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, long long ldx, long long seg_stride, int rows, float eps,
                    const __nv_bfloat16* __restrict__ weight0, const __nv_bfloat16* __restrict__ weight1,
                    const float* __restrict__ cos_sin, int half_dim) {
  constexpr int d = NVT * TPR * 8;
  constexpr int RPC = ROW_THREADS / TPR;
  __shared__ float red[ROW_THREADS / 32];
  const int t = threadIdx.x % TPR;
  const int row = blockIdx.x * RPC + threadIdx.x / TPR;
  const bool live = row < rows;
  const __nv_bfloat16* weight = blockIdx.y ? weight1 : weight0;
  __nv_bfloat16* xr = x + (long long)(live ? row : 0) * ldx + (long long)blockIdx.y * seg_stride;
  uint4 v[NVT];
#pragma unroll
  for (int i = 0; i < NVT; ++i) v[i] = *reinterpret_cast<const uint4*>(xr + (i * TPR + t) * 8);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NVT; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) ss += bf16_lo(w[j]) * bf16_lo(w[j]) + bf16_hi(w[j]) * bf16_hi(w[j]);
  }
  const float r = rsqrtf(row_sum<TPR>(ss, red) * (1.0f / d) + eps);
  if (!live) return;
  float cs[4], sn[4];
  if (cos_sin) {
    // head_dim == 128 => pair index of this thread's first pair inside a head: (t*4) % 64
    const float4* tb = reinterpret_cast<const float4*>(cos_sin + ((long long)row * half_dim + (t * 4) % half_dim) * 2);
    const float4 t0 = __ldg(tb), t1 = __ldg(tb + 1);
    cs[0] = t0.x; sn[0] = t0.y; cs[1] = t0.z; sn[1] = t0.w;
    cs[2] = t1.x; sn[2] = t1.y; cs[3] = t1.z; sn[3] = t1.w;
  }
#pragma unroll
  for (int i = 0; i < NVT; ++i) {
    const int col = (i * TPR + t) * 8;
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(weight + col));
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a, b;
      rms_scale_pair(w[j], r, ww[j], a, b);
      if (cos_sin) {
        const float ra = a * cs[j] - b * sn[j];
        const float rb = a * sn[j] + b * cs[j];
        a = ra; b = rb;
      }
      o[j] = pack_bf16x2(a, b);
    }
    *reinterpret_cast<uint4*>(xr + col) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------- Ulysses scatter
// Same math as rmsnorm_rope_kernel for q (segment 0) and k (segment 1), v (segment 2) is passed through; the result
// is not written back but stored into the receive buffers of the ranks that own the heads: thread columns
// [col, col + 8) of segment g belong to rank col / w (w = d / n_peers columns per rank, a multiple of 128) and land
// there at row rank*rows + row, column g*w + col % w.  Remote stores are 16 bytes per thread, 512 contiguous bytes
// per warp: NVLink-friendly.
struct ScatterDst {
  __nv_bfloat16* recv[GF_MAX_PEERS];
  long long ld_recv;
  int n_peers, rank;
};

template <int NVT, int TPR>
__global__ void __launch_bounds__(ROW_THREADS)
qkv_scatter_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int rows, float eps,
                   const __nv_bfloat16* __restrict__ weight_q, const __nv_bfloat16* __restrict__ weight_k,
                   const float* __restrict__ cos_sin, int half_dim, const ScatterDst dst) {
  constexpr int d = NVT * TPR * 8;
  constexpr int RPC = ROW_THREADS / TPR;
  __shared__ float red[ROW_THREADS / 32];
  const int t = threadIdx.x % TPR;
  const int row = blockIdx.x * RPC + threadIdx.x / TPR;
  const bool live = row < rows;
  const int seg = blockIdx.y;
  const __nv_bfloat16* xr = qkv + (long long)(live ? row : 0) * ld + (long long)seg * d;
  uint4 v[NVT];
#pragma unroll
  for (int i = 0; i < NVT; ++i) v[i] = ld_stream(xr + (i * TPR + t) * 8);
  float r = 1.0f;
  if (seg < 2) {   // uniform per CTA: every thread reaches the barrier inside row_sum
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NVT; ++i) {
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) ss += bf16_lo(w[j]) * bf16_lo(w[j]) + bf16_hi(w[j]) * bf16_hi(w[j]);
    }
    r = rsqrtf(row_sum<TPR>(ss, red) * (1.0f / d) + eps);
  }
  if (!live) return;
  const int w_cols = d / dst.n_peers;
  const long long drow = (long long)dst.rank * rows + row;
  float cs[4], sn[4];
  if (seg < 2) {
    const float4* tb = reinterpret_cast<const float4*>(cos_sin + ((long long)row * half_dim + (t * 4) % half_dim) * 2);
    const float4 t0 = __ldg(tb), t1 = __ldg(tb + 1);
    cs[0] = t0.x; sn[0] = t0.y; cs[1] = t0.z; sn[1] = t0.w;
    cs[2] = t1.x; sn[2] = t1.y; cs[3] = t1.z; sn[3] = t1.w;
  }
  const __nv_bfloat16* weight = seg ? weight_k : weight_q;
#pragma unroll
  for (int i = 0; i < NVT; ++i) {
    const int col = (i * TPR + t) * 8;
    uint4 o = v[i];
    if (seg < 2) {
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(weight + col));
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t oo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a, b;
        rms_scale_pair(w[j], r, ww[j], a, b);
        oo[j] = pack_bf16x2(a * cs[j] - b * sn[j], a * sn[j] + b * cs[j]);
      }
      o = make_uint4(oo[0], oo[1], oo[2], oo[3]);
    }
    const int owner = col / w_cols;
    *reinterpret_cast<uint4*>(dst.recv[owner] + drow * dst.ld_recv + seg * w_cols + (col - owner * w_cols)) = o;
  }
}

template <int NVT, int TPR>
static void launch_ln(cudaStream_t s, const __nv_bfloat16* X, long long ldx, __nv_bfloat16* Y, long long ldy, int rows,
                      float eps, const __nv_bfloat16* SH, const __nv_bfloat16* SC, const __nv_bfloat16* W,
                      const __nv_bfloat16* B) {
  constexpr int RPC = ROW_THREADS / TPR;
  layernorm_kernel<NVT, TPR><<<(rows + RPC - 1) / RPC, ROW_THREADS, 0, s>>>(X, ldx, Y, ldy, rows, eps, SH, SC, W, B);
}

template <int NVT, int TPR>
static void launch_rms(cudaStream_t s, __nv_bfloat16* X, long long ldx, long long seg_stride, int segs, int rows,
                       float eps, const __nv_bfloat16* W0, const __nv_bfloat16* W1, const float* cos_sin, int half) {
  constexpr int RPC = ROW_THREADS / TPR;
  const dim3 grid((rows + RPC - 1) / RPC, segs);
  rmsnorm_rope_kernel<NVT, TPR><<<grid, ROW_THREADS, 0, s>>>(X, ldx, seg_stride, rows, eps, W0, W1, cos_sin, half);
}

static int rms_dispatch(cudaStream_t s, __nv_bfloat16* X, long long ldx, long long seg_stride, int segs, int rows, int d,
                        float eps, const __nv_bfloat16* W0, const __nv_bfloat16* W1, const float* cos_sin, int half) {
  switch (d) {
    case 5120: launch_rms<5, 128>(s, X, ldx, seg_stride, segs, rows, eps, W0, W1, cos_sin, half); break;
    case 1536: launch_rms<3, 64>(s, X, ldx, seg_stride, segs, rows, eps, W0, W1, cos_sin, half); break;
    case 512: launch_rms<2, 32>(s, X, ldx, seg_stride, segs, rows, eps, W0, W1, cos_sin, half); break;
    case 256: launch_rms<1, 32>(s, X, ldx, seg_stride, segs, rows, eps, W0, W1, cos_sin, half); break;
    default: return GF_ERR_UNSUPPORTED;
  }
  return (int)cudaGetLastError();
}

}  // namespace gf

extern "C" int gf_layernorm_bf16(const void* x, long long ldx, void* y, long long ldy, int rows, int d, float eps,
                                 const void* shift, const void* scale, const void* weight, const void* bias,
                                 void* stream) {
  using namespace gf;
  if (!x || !y || rows <= 0 || (ldx % 8) || (ldy % 8)) return GF_ERR_BAD_ARG;
  if (weight ? !bias : (!shift || !scale)) return GF_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto Y = reinterpret_cast<__nv_bfloat16*>(y);
  auto SH = reinterpret_cast<const __nv_bfloat16*>(shift), SC = reinterpret_cast<const __nv_bfloat16*>(scale);
  auto W = reinterpret_cast<const __nv_bfloat16*>(weight), B = reinterpret_cast<const __nv_bfloat16*>(bias);
  switch (d) {
    case 5120: launch_ln<5, 128>(s, X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    case 1536: launch_ln<3, 64>(s, X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    case 512: launch_ln<2, 32>(s, X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    case 256: launch_ln<1, 32>(s, X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    default: return GF_ERR_UNSUPPORTED;
  }
  return (int)cudaGetLastError();
}

extern "C" int gf_rmsnorm_rope_bf16(void* x, long long ldx, int rows, int d, const void* weight, float eps,
                                    const float* cos_sin, int head_dim, void* stream) {
  using namespace gf;
  if (!x || !weight || rows <= 0 || (ldx % 8)) return GF_ERR_BAD_ARG;
  if (cos_sin && head_dim != 128) return GF_ERR_UNSUPPORTED;
  auto W = reinterpret_cast<const __nv_bfloat16*>(weight);
  return rms_dispatch(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<__nv_bfloat16*>(x), ldx, 0, 1, rows, d,
                      eps, W, W, cos_sin, head_dim / 2);
}

extern "C" int gf_qk_rmsnorm_rope_bf16(void* qkv, long long ld, int rows, int d, const void* weight_q,
                                       const void* weight_k, float eps, const float* cos_sin, int head_dim,
                                       void* stream) {
  using namespace gf;
  if (!qkv || !weight_q || !weight_k || rows <= 0 || (ld % 8) || ld < 2LL * d) return GF_ERR_BAD_ARG;
  if (cos_sin && head_dim != 128) return GF_ERR_UNSUPPORTED;
  return rms_dispatch(reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<__nv_bfloat16*>(qkv), ld, d, 2, rows, d,
                      eps, reinterpret_cast<const __nv_bfloat16*>(weight_q),
                      reinterpret_cast<const __nv_bfloat16*>(weight_k), cos_sin, head_dim / 2);
}

extern "C" int gf_qkv_rmsnorm_rope_scatter_bf16(const void* qkv, long long ld, int rows, int d, const void* weight_q,
                                                const void* weight_k, float eps, const float* cos_sin, int head_dim,
                                                void* const* recv_peers, int n_peers, int rank, long long ld_recv,
                                                void* stream) {
  using namespace gf;
  if (!qkv || !weight_q || !weight_k || !cos_sin || !recv_peers || rows <= 0 || (ld % 8) || ld < 3LL * d)
    return GF_ERR_BAD_ARG;
  if (head_dim != 128) return GF_ERR_UNSUPPORTED;
  if (n_peers < 1 || n_peers > GF_MAX_PEERS || rank < 0 || rank >= n_peers || d % (n_peers * 128) || (ld_recv % 8) ||
      ld_recv < 3LL * (d / n_peers))
    return GF_ERR_BAD_ARG;
  ScatterDst dst{};
  for (int i = 0; i < n_peers; ++i) {
    if (!recv_peers[i] || (reinterpret_cast<uintptr_t>(recv_peers[i]) & 15)) return GF_ERR_BAD_ARG;
    dst.recv[i] = reinterpret_cast<__nv_bfloat16*>(recv_peers[i]);
  }
  dst.ld_recv = ld_recv;
  dst.n_peers = n_peers;
  dst.rank = rank;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto X = reinterpret_cast<const __nv_bfloat16*>(qkv);
  auto WQ = reinterpret_cast<const __nv_bfloat16*>(weight_q), WK = reinterpret_cast<const __nv_bfloat16*>(weight_k);
  switch (d) {
    case 5120: qkv_scatter_kernel<5, 128><<<dim3((rows + 1) / 2, 3), ROW_THREADS, 0, s>>>(X, ld, rows, eps, WQ, WK,
                                                                                          cos_sin, 64, dst); break;
    case 1536: qkv_scatter_kernel<3, 64><<<dim3((rows + 3) / 4, 3), ROW_THREADS, 0, s>>>(X, ld, rows, eps, WQ, WK,
                                                                                         cos_sin, 64, dst); break;
    case 512: qkv_scatter_kernel<2, 32><<<dim3((rows + 7) / 8, 3), ROW_THREADS, 0, s>>>(X, ld, rows, eps, WQ, WK,
                                                                                        cos_sin, 64, dst); break;
    case 256: qkv_scatter_kernel<1, 32><<<dim3((rows + 7) / 8, 3), ROW_THREADS, 0, s>>>(X, ld, rows, eps, WQ, WK,
                                                                                        cos_sin, 64, dst); break;
    default: return GF_ERR_UNSUPPORTED;
  }
  return (int)cudaGetLastError();
}
