// gf_misc.cu -- small HBM-bound kernels around the DiT blocks: patch gather / unpatchify (index shuffles, bit-exact),
// modulation add, SiLU, CFG + Euler update, Ulysses pack/unpack.  All are pure streaming kernels; grids are sized to
// a few waves of 148 SMs and every global access is 4-16 bytes per thread, coalesced on the output side.
#include "gf_ptx.cuh"
#include "gf_api_internal.h"

namespace gf {

// out[(f*H2+h)*W2 + w, c*4 + kh*2 + kw] = src[c, f, 2h+kh, 2w+kw]
// One thread produces the 4 bf16 (kh,kw) of one (token, channel): two 4-byte loads (rows 2h, 2h+1), one 8-byte store.
__global__ void patch_gather_kernel(const __nv_bfloat16* __restrict__ s0, int C0, const __nv_bfloat16* __restrict__ s1,
                                    int C1, __nv_bfloat16* __restrict__ out, long long ldo, int F, int H, int W) {
  const int H2 = H >> 1, W2 = W >> 1, C = C0 + C1;
  const long long total = (long long)F * H2 * W2 * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long tok = i / C;
    const int w = (int)(tok % W2);
    const int h = (int)((tok / W2) % H2);
    const int f = (int)(tok / ((long long)W2 * H2));
    const __nv_bfloat16* src = (c < C0) ? s0 + (long long)c * F * H * W : s1 + (long long)(c - C0) * F * H * W;
    const __nv_bfloat16* p = src + ((long long)f * H + 2 * h) * W + 2 * w;
    const uint32_t r0 = *reinterpret_cast<const uint32_t*>(p);
    const uint32_t r1 = *reinterpret_cast<const uint32_t*>(p + W);
    *reinterpret_cast<uint2*>(out + tok * ldo + c * 4) = make_uint2(r0, r1);
  }
}

// out[c, f, 2h+kh, 2w+kw] = tokens[(f*H2+h)*W2 + w, (kh*2+kw)*C + c]      (out is (C, F, H, W))
__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ t, long long ldt, __nv_bfloat16* __restrict__ out,
                                  int C, int F, int H, int W) {
  const int H2 = H >> 1, W2 = W >> 1;
  const long long total = (long long)C * F * H * W2;  // one thread per output pair (kw = 0, 1)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W2);
    const int y = (int)((i / W2) % H);
    const int f = (int)((i / ((long long)W2 * H)) % F);
    const int c = (int)(i / ((long long)W2 * H * F));
    const int h = y >> 1, kh = y & 1;
    const __nv_bfloat16* row = t + (((long long)f * H2 + h) * W2 + w) * ldt;
    const __nv_bfloat16 a = row[(kh * 2 + 0) * C + c];
    const __nv_bfloat16 b = row[(kh * 2 + 1) * C + c];
    __nv_bfloat162 v; v.x = a; v.y = b;
    *reinterpret_cast<__nv_bfloat162*>(out + (((long long)c * F + f) * H + y) * W + 2 * w) = v;
  }
}

__global__ void add_rows_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ y, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i % cols]));
}

// y = a + b elementwise (one bf16 rounding, as torch's bf16 add); y may alias a or b, so no __restrict__.
// n8 = number of 8-element (16-byte) vectors; the scalar tail is handled by the last threads.
__global__ void add_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* y, long long n) {
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 va = reinterpret_cast<const uint4*>(a)[i];
    const uint4 vb = reinterpret_cast<const uint4*>(b)[i];
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = pack_bf16x2(bf16_lo(wa[k]) + bf16_lo(wb[k]), bf16_hi(wa[k]) + bf16_hi(wb[k]));
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (long long i = (n8 << 3) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16_rn(__bfloat162float(a[i]) + __bfloat162float(b[i]));
}

__global__ void silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = __bfloat162float(x[i]);
    y[i] = __float2bfloat16_rn(v / (1.0f + expf(-v)));
  }
}

// pred = nega + s*(posi - nega) ; out = lat + pred*dsigma, bf16 rounding after each torch op of the reference
__global__ void cfg_euler_kernel(const __nv_bfloat16* __restrict__ posi, const __nv_bfloat16* __restrict__ nega,
                                 const __nv_bfloat16* lat, __nv_bfloat16* out /* may alias lat */, float s,
                                 float dsigma, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float p = __bfloat162float(posi[i]);
    if (nega) {
      const float q = __bfloat162float(nega[i]);
      const float diff = round_bf16(p - q);
      const float sc = round_bf16(s * diff);
      p = round_bf16(q + sc);
    }
    const float upd = round_bf16(p * dsigma);
    out[i] = __float2bfloat16_rn(__bfloat162float(lat[i]) + upd);
  }
}

// x[rows, heads, hd] (row pitch ldx) -> out[P][rows][ldo >= heads/P*hd]; 16-byte vectors
__global__ void ulysses_pack_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ out,
                                    long long ldo, int rows, int heads, int hd, int P) {
  const int vec_per_row = heads * hd / 8;
  const int hp = heads / P;
  const int vec_per_head = hd / 8;
  const long long total = (long long)rows * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int vcol = (int)(i % vec_per_row);
    const long long row = i / vec_per_row;
    const int head = vcol / vec_per_head, j = vcol % vec_per_head;
    const int dst = head / hp, hl = head % hp;
    const uint4 v = *reinterpret_cast<const uint4*>(x + row * ldx + (long long)vcol * 8);
    *reinterpret_cast<uint4*>(out + ((long long)dst * rows + row) * ldo + hl * hd + j * 8) = v;
  }
}
// in[P][rows][heads/P][hd] -> y[rows, heads, hd] (row pitch ldy)
__global__ void ulysses_unpack_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ y,
                                      long long ldy, int rows, int heads, int hd, int P) {
  const int vec_per_row = heads * hd / 8;
  const int hp = heads / P;
  const int vec_per_head = hd / 8;
  const long long total = (long long)rows * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int vcol = (int)(i % vec_per_row);
    const long long row = i / vec_per_row;
    const int head = vcol / vec_per_head, j = vcol % vec_per_head;
    const int src = head / hp, hl = head % hp;
    const uint4 v = *reinterpret_cast<const uint4*>(in + ((((long long)src * rows + row) * hp + hl) * hd) + j * 8);
    *reinterpret_cast<uint4*>(y + row * ldy + (long long)vcol * 8) = v;
  }
}

// sinusoidal_embedding_1d (wan_video_dit.py:68-72): float64 angles t * 10000^(-i/half), [cos | sin], cast to bf16
__global__ void timestep_embedding_kernel(const __nv_bfloat16* __restrict__ t, __nv_bfloat16* __restrict__ out, int B,
                                          int dim) {
  const int half = dim / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * half; i += gridDim.x * blockDim.x) {
    const int b = i / half, k = i % half;
    const double pos = (double)__bfloat162float(t[b]);
    const double ang = pos * pow(10000.0, -((double)k / (double)half));
    out[(long long)b * dim + k] = __float2bfloat16_rn((float)cos(ang));
    out[(long long)b * dim + half + k] = __float2bfloat16_rn((float)sin(ang));
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)gf_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace gf

using namespace gf;
typedef __nv_bfloat16 bf16;
#define GF_STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int gf_patch_gather_bf16(const void* src0, int C0, const void* src1, int C1, void* out, long long ldo,
                                    int F, int H, int W, void* stream) {
  if (!src0 || !out || C0 <= 0 || C1 < 0 || (C1 > 0 && !src1) || (H & 1) || (W & 1) || (ldo % 4)) return GF_ERR_BAD_ARG;
  const long long total = (long long)F * (H / 2) * (W / 2) * (C0 + C1);
  patch_gather_kernel<<<grid_for(total, 256), 256, 0, GF_STREAM(stream)>>>(
      (const bf16*)src0, C0, (const bf16*)src1, C1, (bf16*)out, ldo, F, H, W);
  return (int)cudaGetLastError();
}

extern "C" int gf_unpatchify_bf16(const void* tokens, long long ldt, void* out, int C, int F, int H, int W,
                                  void* stream) {
  if (!tokens || !out || C <= 0 || (H & 1) || (W & 1)) return GF_ERR_BAD_ARG;
  const long long total = (long long)C * F * H * (W / 2);
  unpatchify_kernel<<<grid_for(total, 256), 256, 0, GF_STREAM(stream)>>>((const bf16*)tokens, ldt, (bf16*)out, C, F, H,
                                                                         W);
  return (int)cudaGetLastError();
}

extern "C" int gf_add_rows_bf16(const void* a, const void* b, void* y, int rows, int cols, void* stream) {
  if (!a || !b || !y || rows <= 0 || cols <= 0) return GF_ERR_BAD_ARG;
  add_rows_kernel<<<grid_for((long long)rows * cols, 256), 256, 0, GF_STREAM(stream)>>>((const bf16*)a, (const bf16*)b,
                                                                                       (bf16*)y, rows, cols);
  return (int)cudaGetLastError();
}

extern "C" int gf_add_bf16(const void* a, const void* b, void* y, long long n, void* stream) {
  if (!a || !b || !y || n <= 0) return GF_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15)
    return GF_ERR_BAD_ARG;
  add_kernel<<<grid_for((n + 7) / 8, 256), 256, 0, GF_STREAM(stream)>>>((const bf16*)a, (const bf16*)b, (bf16*)y, n);
  return (int)cudaGetLastError();
}

extern "C" int gf_silu_bf16(const void* x, void* y, long long n, void* stream) {
  if (!x || !y || n <= 0) return GF_ERR_BAD_ARG;
  silu_kernel<<<grid_for(n, 256), 256, 0, GF_STREAM(stream)>>>((const bf16*)x, (bf16*)y, n);
  return (int)cudaGetLastError();
}

extern "C" int gf_timestep_embedding_bf16(const void* timestep, void* out, int B, int dim, void* stream) {
  if (!timestep || !out || B <= 0 || dim <= 0 || (dim & 1)) return GF_ERR_BAD_ARG;
  timestep_embedding_kernel<<<grid_for((long long)B * dim / 2, 128), 128, 0, GF_STREAM(stream)>>>(
      (const bf16*)timestep, (bf16*)out, B, dim);
  return (int)cudaGetLastError();
}

extern "C" int gf_cfg_euler_bf16(const void* posi, const void* nega, const void* latents, void* latents_out,
                                 float cfg_scale, float dsigma, long long n, void* stream) {
  if (!posi || !latents || !latents_out || n <= 0) return GF_ERR_BAD_ARG;
  cfg_euler_kernel<<<grid_for(n, 256), 256, 0, GF_STREAM(stream)>>>((const bf16*)posi, (const bf16*)nega,
                                                                   (const bf16*)latents, (bf16*)latents_out, cfg_scale,
                                                                   dsigma, n);
  return (int)cudaGetLastError();
}

extern "C" int gf_ulysses_pack_bf16(const void* x, long long ldx, void* out, long long ldo, int rows, int heads,
                                    int head_dim, int P, void* stream) {
  if (!x || !out || rows <= 0 || P <= 0 || heads % P || head_dim % 8 || (ldx % 8) || (ldo % 8)) return GF_ERR_BAD_ARG;
  if (ldo < (long long)(heads / P) * head_dim) return GF_ERR_BAD_ARG;
  const long long total = (long long)rows * heads * head_dim / 8;
  ulysses_pack_kernel<<<grid_for(total, 256), 256, 0, GF_STREAM(stream)>>>((const bf16*)x, ldx, (bf16*)out, ldo, rows,
                                                                          heads, head_dim, P);
  return (int)cudaGetLastError();
}

extern "C" int gf_ulysses_unpack_bf16(const void* in, void* y, long long ldy, int rows, int heads, int head_dim, int P,
                                      void* stream) {
  if (!in || !y || rows <= 0 || P <= 0 || heads % P || head_dim % 8 || (ldy % 8)) return GF_ERR_BAD_ARG;
  const long long total = (long long)rows * heads * head_dim / 8;
  ulysses_unpack_kernel<<<grid_for(total, 256), 256, 0, GF_STREAM(stream)>>>((const bf16*)in, (bf16*)y, ldy, rows,
                                                                            heads, head_dim, P);
  return (int)cudaGetLastError();
}
