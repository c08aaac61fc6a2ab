// gf_api_internal.h -- shared host-side helpers for the kernels behind include/goalforce_b200.h
#pragma once
#include <cstdint>
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

}  // namespace gf
