// gf_runtime.cu -- host-side plumbing shared by the kernels: device query and TMA tensor-map encoding.
// libcuda is resolved at run time through cudaGetDriverEntryPoint so the library loads (and exports its symbols)
// on a machine without a driver; compute entry points then fail with GF_ERR_NO_DRIVER instead of crashing.
#include <mutex>
#include "gf_api_internal.h"

namespace gf {

int gf_num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int gf_make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                         uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = resolve_encode();
  if (!enc) return GF_ERR_NO_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16 || box_inner * 2 != 128 || box_outer > 256)
    return GF_ERR_BAD_ARG;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};  // bytes, dims 1..rank-1
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : GF_ERR_TMAP;
}

}  // namespace gf

extern "C" int gf_abi_version(void) { return GF_ABI_VERSION; }
extern "C" int gf_device_sms(void) { return gf::gf_num_sms(); }
