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
int gf_attention80_launch(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv,
                          const AttnOut& out, int Lq, int Lk, int heads, float scale, int emu_pairs,
                          cudaStream_t stream);

}  // namespace gf
