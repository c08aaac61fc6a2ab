// gf_rowwise.cu -- the HBM-bound kernels of the DiT block: adaLN LayerNorm and full-row RMSNorm + 3-D RoPE.
// One warp owns one row: the row (d bf16, d % 256 == 0) is read once with 128-bit loads, kept packed in registers,
// reduced with warp shuffles, and written once with 128-bit stores.  Algorithmic traffic = 2 * rows * d * 2 bytes.
#include "gf_ptx.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int ROW_WARPS = 8;  // warps (rows) per CTA

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

// ---------------------------------------------------------------------------------------------- LayerNorm
// weight == nullptr: y = bf16(bf16(bf16(LN(x)) * bf16(1 + scale)) + shift)   (eager-PyTorch rounding chain of
//                    modulate(norm(x), shift, scale), wan_video_dit.py:64-65)
// weight != nullptr: y = bf16(LN(x) * weight + bias)                          (nn.LayerNorm with affine, fp32 math)
template <int NV>  // uint4 vectors per lane: d = NV * 256
__global__ void __launch_bounds__(ROW_WARPS * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                 int rows, float eps, const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
                 const __nv_bfloat16* __restrict__ weight, const __nv_bfloat16* __restrict__ bias) {
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int d = NV * 256;
  const __nv_bfloat16* xr = x + (long long)row * ldx;
  uint4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = ld_stream(xr + (i * 32 + lane) * 8);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) s += bf16_lo(w[j]) + bf16_hi(w[j]);
  }
  const float mean = warp_sum(s) * (1.0f / d);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(w[j]) - mean, b = bf16_hi(w[j]) - mean;
      ss += a * a + b * b;
    }
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / d) + eps);
  __nv_bfloat16* yr = y + (long long)row * ldy;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
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
      const uint4 sc = __ldg(reinterpret_cast<const uint4*>(scale + col));
      const uint4 sh = __ldg(reinterpret_cast<const uint4*>(shift + col));
      const uint32_t cw[4] = {sc.x, sc.y, sc.z, sc.w}, hw[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float n0 = round_bf16((bf16_lo(w[j]) - mean) * rstd), n1 = round_bf16((bf16_hi(w[j]) - mean) * rstd);
        const float m0 = round_bf16(n0 * round_bf16(1.0f + bf16_lo(cw[j])));
        const float m1 = round_bf16(n1 * round_bf16(1.0f + bf16_hi(cw[j])));
        o[j] = pack_bf16x2(m0 + bf16_lo(hw[j]), m1 + bf16_hi(hw[j]));
      }
    }
    *reinterpret_cast<uint4*>(yr + col) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------- RMSNorm + RoPE
// x <- rope( bf16( bf16(x * rsqrt(mean(x^2) + eps)) * weight ) )       in place, one warp per row
// RoPE acts on interleaved pairs (2i, 2i+1) of every head: (a + ib) * (cos + i sin), one rounding to bf16
// (wan_video_dit.py:92-97 does the product in complex128; fp32 FMA on an fp32 table built from the float64
//  angles differs from that by < 1e-7 relative before the bf16 rounding).
// Lane l always sees pairs 4l%64 .. 4l%64+3 of a head (row layout d = heads*128, 8 elements per lane per vector,
// 256 elements per warp-wide vector), so the 4 (cos, sin) pairs are loaded once per row.
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, long long ldx, int rows, float eps,
                    const __nv_bfloat16* __restrict__ weight, const float* __restrict__ cos_sin, int half_dim) {
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int d = NV * 256;
  __nv_bfloat16* xr = x + (long long)row * ldx;
  uint4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const uint4*>(xr + (i * 32 + lane) * 8);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) ss += bf16_lo(w[j]) * bf16_lo(w[j]) + bf16_hi(w[j]) * bf16_hi(w[j]);
  }
  const float r = rsqrtf(warp_sum(ss) * (1.0f / d) + eps);
  float cs[4], sn[4];
  if (cos_sin) {
    // head_dim == 128 => pair index of this lane's first pair inside a head: (lane*4) % 64
    const float4* t = reinterpret_cast<const float4*>(cos_sin + ((long long)row * half_dim + (lane * 4) % half_dim) * 2);
    const float4 t0 = __ldg(t), t1 = __ldg(t + 1);
    cs[0] = t0.x; sn[0] = t0.y; cs[1] = t0.z; sn[1] = t0.w;
    cs[2] = t1.x; sn[2] = t1.y; cs[3] = t1.z; sn[3] = t1.w;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 8;
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(weight + col));
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = round_bf16(round_bf16(bf16_lo(w[j]) * r) * bf16_lo(ww[j]));
      float b = round_bf16(round_bf16(bf16_hi(w[j]) * r) * bf16_hi(ww[j]));
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

}  // namespace gf

extern "C" int gf_layernorm_bf16(const void* x, long long ldx, void* y, long long ldy, int rows, int d, float eps,
                                 const void* shift, const void* scale, const void* weight, const void* bias,
                                 void* stream) {
  using namespace gf;
  if (!x || !y || rows <= 0 || (ldx % 8) || (ldy % 8)) return GF_ERR_BAD_ARG;
  if (weight ? !bias : (!shift || !scale)) return GF_ERR_BAD_ARG;
  const dim3 grid((rows + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto Y = reinterpret_cast<__nv_bfloat16*>(y);
  auto SH = reinterpret_cast<const __nv_bfloat16*>(shift), SC = reinterpret_cast<const __nv_bfloat16*>(scale);
  auto W = reinterpret_cast<const __nv_bfloat16*>(weight), B = reinterpret_cast<const __nv_bfloat16*>(bias);
  switch (d) {
    case 5120: layernorm_kernel<20><<<grid, block, 0, s>>>(X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    case 1536: layernorm_kernel<6><<<grid, block, 0, s>>>(X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    case 256: layernorm_kernel<1><<<grid, block, 0, s>>>(X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    case 512: layernorm_kernel<2><<<grid, block, 0, s>>>(X, ldx, Y, ldy, rows, eps, SH, SC, W, B); break;
    default: return GF_ERR_UNSUPPORTED;
  }
  return (int)cudaGetLastError();
}

extern "C" int gf_rmsnorm_rope_bf16(void* x, long long ldx, int rows, int d, const void* weight, float eps,
                                    const float* cos_sin, int head_dim, void* stream) {
  using namespace gf;
  if (!x || !weight || rows <= 0 || (ldx % 8)) return GF_ERR_BAD_ARG;
  if (cos_sin && head_dim != 128) return GF_ERR_UNSUPPORTED;
  const dim3 grid((rows + ROW_WARPS - 1) / ROW_WARPS), block(ROW_WARPS * 32);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto X = reinterpret_cast<__nv_bfloat16*>(x);
  auto W = reinterpret_cast<const __nv_bfloat16*>(weight);
  switch (d) {
    case 5120: rmsnorm_rope_kernel<20><<<grid, block, 0, s>>>(X, ldx, rows, eps, W, cos_sin, head_dim / 2); break;
    case 1536: rmsnorm_rope_kernel<6><<<grid, block, 0, s>>>(X, ldx, rows, eps, W, cos_sin, head_dim / 2); break;
    case 256: rmsnorm_rope_kernel<1><<<grid, block, 0, s>>>(X, ldx, rows, eps, W, cos_sin, head_dim / 2); break;
    case 512: rmsnorm_rope_kernel<2><<<grid, block, 0, s>>>(X, ldx, rows, eps, W, cos_sin, head_dim / 2); break;
    default: return GF_ERR_UNSUPPORTED;
  }
  return (int)cudaGetLastError();
}
