// gf_vae.cu -- the HBM-bound kernels of the Wan video VAE around the implicit-GEMM convolutions (gf_conv.cu) and the
// GEMMs (gf_gemm.cu): channel-wise RMS_norm (+ SiLU) on channels-last rows, nearest-exact 2x spatial upsampling with
// the temporal frame interleave of `upsample3d` folded into the gather, softmax over fp32 score rows, layout changes
// between the reference's (C, T, H, W) tensors and the channels-last clips (with the latent mean / std affine), and
// the tile blending of WanVideoVAE.tiled_decode / tiled_encode.  Streaming kernels: 16-byte accesses, grid-stride
// loops over a few waves of the SMs.  Reference: diffsynth/models/wan_video_vae.py.
#include "gf_ptx.cuh"
#include "gf_api_internal.h"

namespace gf {

static inline int stream_grid(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)gf_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------------- RMS_norm (+SiLU)
// RMS_norm.forward (:55-70): F.normalize(x, dim=channel) * sqrt(C) * gamma, then nn.SiLU in ResidualBlock / head.
// G lanes share one row (G = power of two <= 32, G * 8 * kMaxVec >= C); 32 / G rows per warp.
// NV = 16-byte vectors per lane, U = independent row groups a warp keeps in flight per iteration: with one 16-byte
// load per lane outstanding the kernel sat at 2.4 TB/s (latency-bound: 32 warps x 384 B per SM in flight); U x NV = 4
// loads per lane are issued before the first reduction.
constexpr int VAE_NORM_MAXVEC = 4;
template <int NV, int U>
__global__ void __launch_bounds__(256)
vae_rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                   long long rows, int C, const __nv_bfloat16* __restrict__ gamma, int silu, int g_shift) {
  const int G = 1 << g_shift;
  const int rows_per_warp = 32 >> g_shift;
  const int lane = threadIdx.x & 31;
  const int sub = lane & (G - 1);
  const int nvec = C >> 3;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float sqrt_c = sqrtf((float)C);
  uint4 gv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = sub + i * G;
    gv[i] = vi < nvec ? __ldg(reinterpret_cast<const uint4*>(gamma + vi * 8)) : make_uint4(0, 0, 0, 0);
  }
  for (long long base = warp_global * rows_per_warp * U; base < rows; base += num_warps * rows_per_warp * U) {
    uint4 v[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = base + u * rows_per_warp + (lane >> g_shift);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = sub + i * G;
        v[u][i] = make_uint4(0, 0, 0, 0);
        if (row < rows && vi < nvec) v[u][i] = *reinterpret_cast<const uint4*>(x + row * ldx + vi * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = base + u * rows_per_warp + (lane >> g_shift);
      float ss = 0.0f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const uint32_t w[4] = {v[u][i].x, v[u][i].y, v[u][i].z, v[u][i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) ss += bf16_lo(w[j]) * bf16_lo(w[j]) + bf16_hi(w[j]) * bf16_hi(w[j]);
      }
      for (int o = G >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rinv = sqrt_c / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = sub + i * G;
        if (row < rows && vi < nvec) {
          const uint32_t w[4] = {v[u][i].x, v[u][i].y, v[u][i].z, v[u][i].w};
          const uint32_t gw[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a = bf16_lo(w[j]) * rinv * bf16_lo(gw[j]);
            float b = bf16_hi(w[j]) * rinv * bf16_hi(gw[j]);
            if (silu) {
              a = silu_fast(a);
              b = silu_fast(b);
            }
            o[j] = pack_bf16x2(a, b);
          }
          *reinterpret_cast<uint4*>(y + row * ldy + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- 2x upsample
// Upsample(scale_factor=(2,2), mode='nearest-exact') per frame (:74-80,92-101): out[f, 2h+a, 2w+b, :] = src_f[h, w, :].
// With `rest` != null this is `upsample3d` (:138-160): output frame 0 reads frame 0 of `first`; output frame f >= 1 reads
// the time_conv result rest[(f-1) >> 1] at channel offset ((f-1) & 1) * C  (reshape(b, 2, c, t, h, w) + stack).
__global__ void vae_upsample2x_kernel(const __nv_bfloat16* __restrict__ first, long long ld_first,
                                      const __nv_bfloat16* __restrict__ rest, long long ld_rest,
                                      __nv_bfloat16* __restrict__ out, long long ldo, int F, int H, int W, int C) {
  const int nvec = C >> 3;
  const long long total = (long long)F * H * W * nvec;     // one thread per source vector, four stores
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int vi = (int)(i % nvec);
    const long long pos = i / nvec;
    const int w = (int)(pos % W);
    const int h = (int)((pos / W) % H);
    const int f = (int)(pos / ((long long)W * H));
    const __nv_bfloat16* src;
    if (rest == nullptr) {
      src = first + (((long long)f * H + h) * W + w) * ld_first + vi * 8;
    } else if (f == 0) {
      src = first + ((long long)h * W + w) * ld_first + vi * 8;
    } else {
      const int tf = (f - 1) >> 1, half = (f - 1) & 1;
      src = rest + (((long long)tf * H + h) * W + w) * ld_rest + half * C + vi * 8;
    }
    const uint4 v = *reinterpret_cast<const uint4*>(src);
    __nv_bfloat16* o = out + (((long long)f * 2 * H + 2 * h) * (2 * W) + 2 * w) * ldo + vi * 8;
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + ldo) = v;
    *reinterpret_cast<uint4*>(o + (long long)2 * W * ldo) = v;
    *reinterpret_cast<uint4*>(o + (long long)2 * W * ldo + ldo) = v;
  }
}

// ---------------------------------------------------------------------------------------------------- softmax
// P[r, :L] = softmax(S[r, :L] * scale) as bf16, P[r, L:Lp] = 0.  One CTA per row; the row is staged in shared memory.
__global__ void vae_softmax_kernel(const float* __restrict__ S, long long lds, __nv_bfloat16* __restrict__ P,
                                   long long ldp, int L, int Lp, float scale_log2e) {
  extern __shared__ float row_s[];
  __shared__ float red[32];
  const long long r = blockIdx.x;
  const float* s = S + r * lds;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float v = s[i] * scale_log2e;
    row_s[i] = v;
    m = fmaxf(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.0f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float e = exp2f(row_s[i] - m);
    row_s[i] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.0f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.0f / sum;
  __nv_bfloat16* p = P + r * ldp;
  for (int i = threadIdx.x; i < Lp; i += blockDim.x)
    p[i] = __float2bfloat16_rn(i < L ? row_s[i] * inv : 0.0f);
}

// ---------------------------------------------------------------------------------------------------- layout changes
// (C, N) planes -> channels-last rows [N][ldo] with channels [C, Cp) zero.  mode 0: copy; mode 1: the decode-side
// un-normalisation z / inv_std + mean (VideoVAE_.decode :1014-1018: z / scale[1] + scale[0], bf16 after each op).
__global__ void vae_planes_to_cl_kernel(const __nv_bfloat16* __restrict__ src, long long N, int C,
                                        __nv_bfloat16* __restrict__ dst, long long ldo, int Cp,
                                        const float* __restrict__ mean, const float* __restrict__ inv_std, int mode,
                                        int row_w, int wpad) {
  const int nvec = Cp >> 3;
  const long long total = N * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / nvec;                  // consecutive threads: consecutive positions of one channel group
    const int vi = (int)(i % nvec);
    // destination position: dense, or rows of row_w positions with wpad untouched positions on either side
    long long nd = n;
    if (row_w > 0) {                               // 32-bit division (N < 2^31 is checked on the host)
      const unsigned row = (unsigned)n / (unsigned)row_w;
      nd = (long long)row * (row_w + 2 * wpad) + ((unsigned)n - row * (unsigned)row_w) + wpad;
    }
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = vi * 8 + j;
      float v = 0.0f;
      if (c < C) {
        v = __bfloat162float(src[(long long)c * N + n]);
        if (mode == 1) v = round_bf16(round_bf16(v / round_bf16(inv_std[c])) + round_bf16(mean[c]));
      }
      x[j] = v;
    }
    *reinterpret_cast<uint4*>(dst + nd * ldo + vi * 8) =
        make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
  }
}

// channels-last rows [N][ld] -> (C, N) planes.  mode 0: copy; mode 1: the encode-side normalisation
// (mu - mean) * inv_std (VideoVAE_.encode :1003-1007), bf16 after each op.
__global__ void vae_cl_to_planes_kernel(const __nv_bfloat16* __restrict__ src, long long ld, long long N, int C,
                                        __nv_bfloat16* __restrict__ dst, const float* __restrict__ mean,
                                        const float* __restrict__ inv_std, int mode) {
  const long long total = N * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / N);
    const long long n = i - (long long)c * N;
    float v = __bfloat162float(src[n * ld + c]);
    if (mode == 1) v = round_bf16(round_bf16(v - round_bf16(mean[c])) * round_bf16(inv_std[c]));
    dst[i] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------------- 3-channel head
// The decoder head is a 3x3x3 convolution to 3 channels (:801-803): as an implicit GEMM its N would be 3 padded to 32.
// It runs instead as a (3,1,1) convolution whose output channels are the 9 spatial taps x (3 channels + 1 pad):
//   D[t, h, w, tap*4 + co] = sum_{dt, ci} W[co, ci, dt, dh, dw] * X[t + dt - 2, h, w, ci],   tap = dh*3 + dw
// and this kernel gathers  out[co, t, h, w] = bias[co] + sum_tap D[t, h + dh - 1, w + dw - 1, tap*4 + co]  (positions
// outside the frame contribute zero: the convolution's zero padding), fp32 sum of the bf16 partials, one rounding.
__global__ void vae_head_gather_kernel(const __nv_bfloat16* __restrict__ D, long long ld, const float* __restrict__ bias,
                                       __nv_bfloat16* __restrict__ out, int C, int T, int H, int W) {
  const long long plane = (long long)T * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const long long t = i / ((long long)W * H);
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int hh = h + dh - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        const int ww = w + dw - 1;
        if (ww < 0 || ww >= W) continue;
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(D + ((t * H + hh) * W + ww) * ld + (dh * 3 + dw) * 4));
        acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      }
    }
    for (int c = 0; c < C; ++c) out[c * plane + i] = __float2bfloat16_rn(acc[c] + bias[c]);
  }
}

// ---------------------------------------------------------------------------------------------------- tile blending
// WanVideoVAE.tiled_decode / tiled_encode (:1133-1153,1184-1204): values[:, :, h0:h0+th, w0:w0+tw] += tile * mask with
// the bf16 rounding of the reference's two torch ops (mul, then add).  tile: (C, T, th, tw); mask: [th][tw] bf16.
__global__ void vae_blend_kernel(__nv_bfloat16* __restrict__ values, int C, int T, int H, int W,
                                 const __nv_bfloat16* __restrict__ tile, int th, int tw, int h0, int w0,
                                 const __nv_bfloat16* __restrict__ mask) {
  const long long total = (long long)C * T * th * tw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % tw);
    const int y = (int)((i / tw) % th);
    const long long ct = i / ((long long)tw * th);
    const float prod = round_bf16(__bfloat162float(tile[i]) * __bfloat162float(mask[y * tw + x]));
    __nv_bfloat16* v = values + (ct * H + (h0 + y)) * W + (w0 + x);
    *v = __float2bfloat16_rn(__bfloat162float(*v) + prod);
  }
}

// values / weight (weight: [H][W] bf16, the same for every channel and frame), optionally clamped to [-1, 1].
__global__ void vae_blend_finish_kernel(__nv_bfloat16* __restrict__ values, long long planes, int H, int W,
                                        const __nv_bfloat16* __restrict__ weight, int clamp) {
  const long long hw = (long long)H * W;
  const long long total = planes * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float v = round_bf16(__bfloat162float(values[i]) / __bfloat162float(weight[i % hw]));
    if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
    values[i] = __float2bfloat16_rn(v);
  }
}

__global__ void vae_clamp_kernel(__nv_bfloat16* __restrict__ x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __float2bfloat16_rn(fminf(fmaxf(__bfloat162float(x[i]), -1.0f), 1.0f));
}

}  // namespace gf

using namespace gf;

extern "C" int gf_vae_rmsnorm_bf16(const void* x, long long ldx, void* y, long long ldy, long long rows, int C,
                                   const void* gamma, int silu, void* stream) {
  if (!x || !y || !gamma || rows <= 0 || C <= 0 || (C % 8) || (ldx % 8) || (ldy % 8)) return GF_ERR_BAD_ARG;
  if (C > 32 * 8 * VAE_NORM_MAXVEC) return GF_ERR_UNSUPPORTED;
  const int nvec = C / 8;
  int g_shift = 0;
  while ((1 << g_shift) < nvec && g_shift < 5) ++g_shift;
  const int rows_per_warp = 32 >> g_shift;
  const int nv = (nvec + (1 << g_shift) - 1) >> g_shift;            // vectors per lane: 1 (C <= 256), 2 (<= 512), 4
  const int unroll = nv == 1 ? 4 : (nv == 2 ? 2 : 1);
  const long long warps = (rows + (long long)rows_per_warp * unroll - 1) / ((long long)rows_per_warp * unroll);
  const int threads = 256;
  const int grid = stream_grid(warps * 32, threads);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(gamma);
  if (nv == 1) vae_rmsnorm_kernel<1, 4><<<grid, threads, 0, s>>>(xp, ldx, yp, ldy, rows, C, gp, silu, g_shift);
  else if (nv == 2) vae_rmsnorm_kernel<2, 2><<<grid, threads, 0, s>>>(xp, ldx, yp, ldy, rows, C, gp, silu, g_shift);
  else vae_rmsnorm_kernel<4, 1><<<grid, threads, 0, s>>>(xp, ldx, yp, ldy, rows, C, gp, silu, g_shift);
  return (int)cudaGetLastError();
}

extern "C" int gf_vae_upsample2x_bf16(const void* first, long long ld_first, const void* rest, long long ld_rest,
                                      void* out, long long ldo, int F, int H, int W, int C, void* stream) {
  if (!first || !out || F <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8) || (ld_first % 8) || (ldo % 8) ||
      (rest && (ld_rest % 8)))
    return GF_ERR_BAD_ARG;
  const long long total = (long long)F * H * W * (C / 8);
  vae_upsample2x_kernel<<<stream_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(first), ld_first, reinterpret_cast<const __nv_bfloat16*>(rest), ld_rest,
      reinterpret_cast<__nv_bfloat16*>(out), ldo, F, H, W, C);
  return (int)cudaGetLastError();
}

extern "C" int gf_softmax_f32_bf16(const float* S, long long lds, void* P, long long ldp, int rows, int L, int Lp,
                                   float scale, void* stream) {
  if (!S || !P || rows <= 0 || L <= 0 || Lp < L || ldp < Lp || lds < L) return GF_ERR_BAD_ARG;
  if ((size_t)L * 4 > 200 * 1024) return GF_ERR_UNSUPPORTED;
  const int smem = L * 4;
  static bool configured[64] = {};
  if (int e = gf_set_smem_once(configured, reinterpret_cast<const void*>(vae_softmax_kernel), 200 * 1024)) return e;
  vae_softmax_kernel<<<rows, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      S, lds, reinterpret_cast<__nv_bfloat16*>(P), ldp, L, Lp, scale * 1.4426950408889634f);
  return (int)cudaGetLastError();
}

extern "C" int gf_vae_planes_to_cl_bf16(const void* src, long long N, int C, void* dst, long long ldo, int Cp,
                                        const float* mean, const float* inv_std, int mode, int row_w, int wpad,
                                        void* stream) {
  if (!src || !dst || N <= 0 || C <= 0 || Cp < C || (Cp % 8) || (ldo % 8) || ldo < Cp) return GF_ERR_BAD_ARG;
  if (row_w < 0 || wpad < 0 || (row_w == 0 && wpad != 0) || (row_w > 0 && (N % row_w || N >= (1ll << 31))))
    return GF_ERR_BAD_ARG;
  if (mode != 0 && mode != 1) return GF_ERR_BAD_ARG;
  if (mode == 1 && (!mean || !inv_std)) return GF_ERR_BAD_ARG;
  const long long total = N * (Cp / 8);
  vae_planes_to_cl_kernel<<<stream_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), N, C, reinterpret_cast<__nv_bfloat16*>(dst), ldo, Cp, mean, inv_std,
      mode, row_w, wpad);
  return (int)cudaGetLastError();
}

extern "C" int gf_vae_cl_to_planes_bf16(const void* src, long long ld, long long N, int C, void* dst, const float* mean,
                                        const float* inv_std, int mode, void* stream) {
  if (!src || !dst || N <= 0 || C <= 0 || ld < C) return GF_ERR_BAD_ARG;
  if (mode != 0 && mode != 1) return GF_ERR_BAD_ARG;
  if (mode == 1 && (!mean || !inv_std)) return GF_ERR_BAD_ARG;
  vae_cl_to_planes_kernel<<<stream_grid(N * C, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), ld, N, C, reinterpret_cast<__nv_bfloat16*>(dst), mean, inv_std, mode);
  return (int)cudaGetLastError();
}

extern "C" int gf_vae_blend_bf16(void* values, int C, int T, int H, int W, const void* tile, int th, int tw, int h0,
                                 int w0, const void* mask, void* stream) {
  if (!values || !tile || !mask || C <= 0 || T <= 0 || th <= 0 || tw <= 0 || h0 < 0 || w0 < 0 || h0 + th > H ||
      w0 + tw > W)
    return GF_ERR_BAD_ARG;
  const long long total = (long long)C * T * th * tw;
  vae_blend_kernel<<<stream_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(values), C, T, H, W, reinterpret_cast<const __nv_bfloat16*>(tile), th, tw, h0,
      w0, reinterpret_cast<const __nv_bfloat16*>(mask));
  return (int)cudaGetLastError();
}

extern "C" int gf_vae_blend_finish_bf16(void* values, long long planes, int H, int W, const void* weight, int clamp,
                                        void* stream) {
  if (!values || planes <= 0 || H <= 0 || W <= 0) return GF_ERR_BAD_ARG;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long total = planes * H * W;
  if (weight)
    vae_blend_finish_kernel<<<stream_grid(total, 256), 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(values), planes, H,
                                                                    W, reinterpret_cast<const __nv_bfloat16*>(weight),
                                                                    clamp);
  else if (clamp)
    vae_clamp_kernel<<<stream_grid(total, 256), 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(values), total);
  return (int)cudaGetLastError();
}

extern "C" int gf_vae_head_gather_bf16(const void* D, long long ld, const float* bias, void* out, int C, int T, int H,
                                       int W, void* stream) {
  if (!D || !bias || !out || C < 1 || C > 4 || T <= 0 || H <= 0 || W <= 0 || (ld % 8) || ld < 36) return GF_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(D) & 15)) return GF_ERR_BAD_ARG;
  const long long plane = (long long)T * H * W;
  vae_head_gather_kernel<<<stream_grid(plane, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(D), ld, bias, reinterpret_cast<__nv_bfloat16*>(out), C, T, H, W);
  return (int)cudaGetLastError();
}
