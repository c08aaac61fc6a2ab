// gf_t5.cu -- the pieces of the umT5-XXL prompt encoder that the DiT kernels do not already cover
// (diffsynth/models/wan_video_text_encoder.py).  The encoder runs twice per video (positive / negative
// prompt, 512 tokens each, ~5 TFLOP), far off the denoising hot path; its linears go through gf_gemm_bf16, and this
// file adds the row-wise and attention parts as plain CUDA-core kernels -- at 512 x 512 x 64 per head there is no
// tensor-core-sized problem to feed:
//   gf_embedding_bf16     token_embedding(ids)                                     :236
//   gf_t5_rmsnorm_bf16    T5LayerNorm: bf16(x * rsqrt(mean(x^2) + eps)) * weight    :18-30   (out of place)
//   gf_t5_attention_bf16  softmax(q k^T + pos_bias + mask) v, head_dim 64, no 1/sqrt(d) scaling   :47-78,
//                         pos_bias = embedding[bucket(j - i)][head]  (T5RelativeEmbedding, :136-175)
//   gf_mul_bf16           fc1(x) * gelu(gate(x))  (the product; GELU is the gate GEMM's epilogue)  :95-100
#include "gf_ptx.cuh"
#include "gf_api_internal.h"

namespace gf {

constexpr int T5_HD = 64;          // head dim
constexpr int T5_MAX_LK = 512;     // text_len of the Wan prompter
constexpr int T5_ROWS = 32;        // query rows per CTA
constexpr int T5_THREADS = 256;    // 8 warps, 4 query rows each
constexpr int T5_KT_LD = T5_MAX_LK + 2;   // padded row of the transposed K tile (bank-conflict-free transposing stores)
constexpr int T5_SMEM = T5_HD * T5_KT_LD * 2 + T5_MAX_LK * T5_HD * 2 + (T5_THREADS / 32) * T5_MAX_LK * 4;

__global__ void embedding_kernel(const long long* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                                 __nv_bfloat16* __restrict__ out, int rows, int dim, long long vocab) {
  const int vec = dim / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)rows * vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec), c = (int)(i % vec);
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    reinterpret_cast<uint4*>(out + (long long)r * dim)[c] = __ldg(reinterpret_cast<const uint4*>(table + id * dim) + c);
  }
}

// one CTA per row; any d % 8 == 0
__global__ void __launch_bounds__(256)
t5_rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                  const __nv_bfloat16* __restrict__ w, int d, float eps) {
  __shared__ float red[8];
  const __nv_bfloat16* xr = x + (long long)blockIdx.x * ldx;
  __nv_bfloat16* yr = y + (long long)blockIdx.x * ldy;
  float ss = 0.f;
  for (int c = threadIdx.x * 8; c < d; c += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) ss += bf16_lo(u[j]) * bf16_lo(u[j]) + bf16_hi(u[j]) * bf16_hi(u[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += red[k];
  const float r = rsqrtf(tot / (float)d + eps);
  for (int c = threadIdx.x * 8; c < d; c += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + c);
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + c));
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)      // bf16(x * r) then * weight with one more rounding, as the reference's two ops
      o[j] = pack_bf16x2(round_bf16(bf16_lo(u[j]) * r) * bf16_lo(ww[j]), round_bf16(bf16_hi(u[j]) * r) * bf16_hi(ww[j]));
    *reinterpret_cast<uint4*>(yr + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void mul_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* y, long long n) {
  const long long n8 = n >> 3, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 va = reinterpret_cast<const uint4*>(a)[i], vb = reinterpret_cast<const uint4*>(b)[i];
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = pack_bf16x2(bf16_lo(wa[k]) * bf16_lo(wb[k]), bf16_hi(wa[k]) * bf16_hi(wb[k]));
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (long long i = (n8 << 3) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16_rn(__bfloat162float(a[i]) * __bfloat162float(b[i]));
}

// grid (ceil(Lq / 32), heads, batch).  q/k/v/o: [batch*L, ld] with head h at columns [64 h, 64 h + 64).
// bias_table: [num_buckets, heads] bf16 (T5RelativeEmbedding.embedding.weight); bucket_of: int32 [Lq + Lk - 1] indexed
// by (j - i) + (Lq - 1), built on the host with the reference's own formula; key_mask: [batch, Lk] int32 (0 = padding)
// or null.  Scores and softmax in fp32; the probabilities are rounded to bf16 before the PV product like the
// reference's `.type_as(attn)`.
__global__ void __launch_bounds__(T5_THREADS)
t5_attention_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ k,
                    long long ldk, const __nv_bfloat16* __restrict__ v, long long ldv, __nv_bfloat16* __restrict__ o,
                    long long ldo, int Lq, int Lk, const __nv_bfloat16* __restrict__ bias_table, int heads,
                    const int* __restrict__ bucket_of, const int* __restrict__ key_mask) {
  extern __shared__ uint8_t smem_raw[];
  __nv_bfloat16* kt = reinterpret_cast<__nv_bfloat16*>(smem_raw);                  // [64][T5_KT_LD]
  __nv_bfloat16* vs = kt + T5_HD * T5_KT_LD;                                       // [Lk][64]
  float* ps = reinterpret_cast<float*>(vs + T5_MAX_LK * T5_HD);                     // [8 warps][512]
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long kv0 = (long long)b * Lk, q0 = (long long)b * Lq;
  // stage K (transposed) and V of this head: thread t copies 8 elements of a row at a time
  for (int i = threadIdx.x; i < Lk * (T5_HD / 8); i += T5_THREADS) {
    const int j = i / (T5_HD / 8), c = (i % (T5_HD / 8)) * 8;
    const uint4 kv = *reinterpret_cast<const uint4*>(k + (kv0 + j) * ldk + h * T5_HD + c);
    const __nv_bfloat16* ke = reinterpret_cast<const __nv_bfloat16*>(&kv);
#pragma unroll
    for (int e = 0; e < 8; ++e) kt[(c + e) * T5_KT_LD + j] = ke[e];
    *reinterpret_cast<uint4*>(vs + j * T5_HD + c) = *reinterpret_cast<const uint4*>(v + (kv0 + j) * ldv + h * T5_HD + c);
  }
  __syncthreads();
  float* pw = ps + warp * T5_MAX_LK;
  const int nt = (Lk + 31) / 32;
  for (int rr = warp; rr < T5_ROWS; rr += T5_THREADS / 32) {
    const int i = blockIdx.x * T5_ROWS + rr;
    if (i >= Lq) break;                                   // warp-uniform
    float qv[T5_HD];
    const __nv_bfloat16* qr = q + (q0 + i) * ldq + h * T5_HD;
#pragma unroll
    for (int c = 0; c < T5_HD; c += 8) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(qr + c));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { qv[c + 2 * e] = bf16_lo(w[e]); qv[c + 2 * e + 1] = bf16_hi(w[e]); }
    }
    float s[T5_MAX_LK / 32];
    float m = -3.0e38f;
#pragma unroll
    for (int t = 0; t < T5_MAX_LK / 32; ++t) {
      const int j = t * 32 + lane;
      float acc = -3.0e38f;
      if (t < nt && j < Lk) {
        acc = 0.f;
#pragma unroll
        for (int d = 0; d < T5_HD; ++d) acc += qv[d] * __bfloat162float(kt[d * T5_KT_LD + j]);
        acc += __bfloat162float(bias_table[(long long)bucket_of[j - i + Lq - 1] * heads + h]);
        if (key_mask && key_mask[(long long)b * Lk + j] == 0) acc = -3.3895314e38f;    // finfo(bf16).min
      }
      s[t] = acc;
      m = fmaxf(m, acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < T5_MAX_LK / 32; ++t) {
      const int j = t * 32 + lane;
      const float e = (t < nt && j < Lk) ? __expf(s[t] - m) : 0.f;
      s[t] = e;
      sum += e;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int t = 0; t < T5_MAX_LK / 32; ++t) {
      const int j = t * 32 + lane;
      if (t < nt && j < Lk) pw[j] = round_bf16(s[t] * inv);
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Lk; ++j) {
      const float p = pw[j];
      const uint32_t vv = *reinterpret_cast<const uint32_t*>(vs + j * T5_HD + 2 * lane);
      o0 += p * bf16_lo(vv);
      o1 += p * bf16_hi(vv);
    }
    *reinterpret_cast<uint32_t*>(o + (q0 + i) * ldo + h * T5_HD + 2 * lane) = pack_bf16x2(o0, o1);
    __syncwarp();
  }
}

}  // namespace gf

using namespace gf;
#define GF_STREAM(s) reinterpret_cast<cudaStream_t>(s)

static inline int t5_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

extern "C" int gf_embedding_bf16(const long long* ids, const void* table, void* out, int rows, int dim,
                                 long long vocab, void* stream) {
  if (!ids || !table || !out || rows <= 0 || dim <= 0 || (dim % 8) || vocab <= 0) return GF_ERR_BAD_ARG;
  embedding_kernel<<<t5_grid((long long)rows * dim / 8, 256), 256, 0, GF_STREAM(stream)>>>(
      ids, (const __nv_bfloat16*)table, (__nv_bfloat16*)out, rows, dim, vocab);
  return (int)cudaGetLastError();
}

extern "C" int gf_t5_rmsnorm_bf16(const void* x, long long ldx, void* y, long long ldy, int rows, int d,
                                  const void* weight, float eps, void* stream) {
  if (!x || !y || !weight || rows <= 0 || d <= 0 || (d % 8) || (ldx % 8) || (ldy % 8)) return GF_ERR_BAD_ARG;
  t5_rmsnorm_kernel<<<rows, 256, 0, GF_STREAM(stream)>>>((const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy,
                                                         (const __nv_bfloat16*)weight, d, eps);
  return (int)cudaGetLastError();
}

extern "C" int gf_mul_bf16(const void* a, const void* b, void* y, long long n, void* stream) {
  if (!a || !b || !y || n <= 0) return GF_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15)
    return GF_ERR_BAD_ARG;
  mul_kernel<<<t5_grid((n + 7) / 8, 256), 256, 0, GF_STREAM(stream)>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                        (__nv_bfloat16*)y, n);
  return (int)cudaGetLastError();
}

extern "C" int gf_t5_attention_bf16(const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                                    long long ldv, void* O, long long ldo, int batch, int Lq, int Lk, int heads,
                                    int head_dim, const void* bias_table, const int* bucket_of, const int* key_mask,
                                    void* stream) {
  if (!Q || !K || !V || !O || !bias_table || !bucket_of || batch <= 0 || Lq <= 0 || Lk <= 0 || heads <= 0)
    return GF_ERR_BAD_ARG;
  if (head_dim != T5_HD || Lk > T5_MAX_LK) return GF_ERR_UNSUPPORTED;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return GF_ERR_BAD_ARG;
  static bool configured[64] = {};
  if (int rc = gf_set_smem_once(configured, reinterpret_cast<const void*>(t5_attention_kernel), T5_SMEM)) return rc;
  const dim3 grid((Lq + T5_ROWS - 1) / T5_ROWS, heads, batch);
  t5_attention_kernel<<<grid, T5_THREADS, T5_SMEM, GF_STREAM(stream)>>>(
      (const __nv_bfloat16*)Q, ldq, (const __nv_bfloat16*)K, ldk, (const __nv_bfloat16*)V, ldv, (__nv_bfloat16*)O, ldo,
      Lq, Lk, (const __nv_bfloat16*)bias_table, heads, bucket_of, key_mask);
  return (int)cudaGetLastError();
}
