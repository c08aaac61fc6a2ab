// gf_api_internal.h -- shared host-side helpers for the kernels behind include/goalforce_b200.h
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/goalforce_b200.h"

namespace gf {

// Number of SMs on the current device (cached per device).
int gf_num_sms();

// Encode a 2-D bf16 row-major tensor map with 128-byte swizzle.
//   inner/outer: logical extent (elements / rows); ld: row pitch in elements; box_inner x box_outer: TMA box.
// Out-of-range box elements read as zero. Returns 0 or a GF_ERR_* / cudaError code.
int gf_make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                         uint32_t box_inner, uint32_t box_outer);

// 4-D bf16 tensor map with 128-byte swizzle: dims[0] is the contiguous one, strides_bytes[i] is the pitch of dims[i+1]
// (multiples of 16), box[0] * 2 == 128.  Used by the implicit-GEMM convolution (gf_conv.cu); not cached.
int gf_make_tmap_4d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4]);

// Tensor map for (base, dims, pitch, box): looked up in / inserted into the context's cache when ctx != nullptr,
// otherwise encoded into `scratch`.  Returns nullptr and sets *rc on failure.  A tensor map depends only on the
// address and the geometry, never on the contents, so a cached entry stays valid when a buffer is freed and another
// one with the same geometry lands on the same address.
const CUtensorMap* gf_ctx_tmap(gf_ctx* ctx, CUtensorMap* scratch, const void* base, uint64_t inner, uint64_t outer,
                               uint64_t ld, uint32_t box_inner, uint32_t box_outer, int* rc);

// Per-context tuning (0 / negative = library default).
struct CtxTuning {
  int attn_impl = 0;       // 0: per shape (80 / 160 for long key sequences, 128 for Lk <= 1024); 80 / 128 / 160: forced
                           // (160 = CTA-pair variant of 80)
  int attn_emu = -1;       // -1: kernel default; 0, 2, 4, 6: column pairs per 16 with exp2 on the FMA pipe
  int gemm_group_m = 0;    // 0: per shape; > 0: rasterisation group height in m-tiles
  int gemm_bn = 0;         // pair GEMM tile width: 0 = per shape (cost model), 224 or 256 = forced
  int conv_impl = 0;       // 0: per shape (halo form, as a CTA pair, for 3x3 windows with Cout <= 128); 1: always
                           // tap-by-tap; 2: halo form on single CTAs
};
CtxTuning gf_ctx_tuning(const gf_ctx* ctx);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): `done` is a per-kernel static array.
inline int gf_set_smem_once(bool (&done)[64], const void* kern, int bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return GF_ERR_BAD_ARG;
  if (done[dev]) return 0;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return (int)e;
  done[dev] = true;
  return 0;
}

// Where an attention kernel stores its output.  n_peers == 0: plain [Lq, ldo] buffer at base[0].  Otherwise the
// Ulysses return path: global query row g goes to base[g / rows_per_peer] at row g % rows_per_peer (pitch ldo),
// column col_offset + head*head_dim.
struct AttnOut {
  void* base[GF_MAX_PEERS];
  long long ldo;
  int n_peers;
  int rows_per_peer;
  int col_offset;
};

// gf_attn80.cu: the decoupled 80-row-block attention kernel behind gf_attention_bf16 (arguments as the C ABI).
int gf_attention80_launch(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                          const AttnOut& out, int Lq, int Lk, int heads, float scale, int emu_pairs,
                          cudaStream_t stream);

// gf_attn80x2.cu: CTA-pair (cta_group::2) variant of the same kernel; work items of 512 query rows.
int gf_attention80x2_launch(gf_ctx* ctx, const void* Q, long long ldq, const void* K, long long ldk, const void* V,
                            long long ldv, const AttnOut& out, int Lq, int Lk, int heads, float scale, int emu_pairs,
                            cudaStream_t stream);

}  // namespace gf
