// gf_runtime.cu -- host-side plumbing shared by the kernels: device query and TMA tensor-map encoding.
// libcuda is resolved at run time through cudaGetDriverEntryPoint so the library loads (and exports its symbols)
// on a machine without a driver; compute entry points then fail with GF_ERR_NO_DRIVER instead of crashing.
#include <mutex>
#include <new>
#include <unordered_map>
#include "gf_api_internal.h"

// The opaque context of the C ABI: cached TMA descriptors and per-context tuning.  No process-wide state: two
// contexts (e.g. two pipelines in one process) never see each other's settings.
struct gf_ctx {
  struct Key {
    const void* base;
    uint64_t inner, outer, ld;
    uint32_t box_inner, box_outer;
    bool operator==(const Key& o) const {
      return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
             box_outer == o.box_outer;
    }
  };
  struct Hash {
    size_t operator()(const Key& k) const {
      uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
      h ^= (k.inner + 0x632BE59BD9B4E019ull) + (h << 6) + (h >> 2);
      h ^= (k.outer * 0xD6E8FEB86659FD93ull) + (h << 6) + (h >> 2);
      h ^= (k.ld * 31 + k.box_inner * 7 + k.box_outer) + (h << 6) + (h >> 2);
      return (size_t)h;
    }
  };
  std::mutex mu;
  std::unordered_map<Key, CUtensorMap, Hash> tmaps;
  long long hits = 0, misses = 0;
  gf::CtxTuning tuning;
};

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

int gf_make_tmap_4d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                         const uint32_t box[4]) {
  EncodeTiledFn enc = resolve_encode();
  if (!enc) return GF_ERR_NO_DRIVER;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || box[0] * 2 != 128) return GF_ERR_BAD_ARG;
  cuuint64_t d[4];
  cuuint64_t st[3];
  cuuint32_t bx[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    if (dims[i] == 0 || box[i] == 0 || box[i] > 256) return GF_ERR_BAD_ARG;
    d[i] = dims[i];
    bx[i] = box[i];
  }
  for (int i = 0; i < 3; ++i) {
    if (strides_bytes[i] % 16) return GF_ERR_BAD_ARG;
    st[i] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, st, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : GF_ERR_TMAP;
}

const CUtensorMap* gf_ctx_tmap(gf_ctx* ctx, CUtensorMap* scratch, const void* base, uint64_t inner, uint64_t outer,
                               uint64_t ld, uint32_t box_inner, uint32_t box_outer, int* rc) {
  *rc = 0;
  if (!ctx) {
    *rc = gf_make_tmap_2d_bf16(scratch, base, inner, outer, ld, box_inner, box_outer);
    return *rc ? nullptr : scratch;
  }
  const gf_ctx::Key key{base, inner, outer, ld, box_inner, box_outer};
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto it = ctx->tmaps.find(key);
  if (it != ctx->tmaps.end()) {
    ++ctx->hits;
    return &it->second;          // unordered_map never moves its nodes: the pointer stays valid until clear()
  }
  ++ctx->misses;
  *rc = gf_make_tmap_2d_bf16(scratch, base, inner, outer, ld, box_inner, box_outer);
  if (*rc) return nullptr;
  if (ctx->tmaps.size() >= 16384) return scratch;   // bounded: beyond this, behave like the stateless path
  return &ctx->tmaps.emplace(key, *scratch).first->second;
}

CtxTuning gf_ctx_tuning(const gf_ctx* ctx) { return ctx ? ctx->tuning : CtxTuning{}; }

}  // namespace gf

extern "C" int gf_ctx_create(gf_ctx** ctx) {
  if (!ctx) return GF_ERR_BAD_ARG;
  *ctx = new (std::nothrow) gf_ctx();
  return *ctx ? 0 : GF_ERR_BAD_ARG;
}

extern "C" int gf_ctx_destroy(gf_ctx* ctx) {
  if (!ctx) return GF_ERR_BAD_ARG;
  delete ctx;
  return 0;
}

extern "C" int gf_ctx_set_attention(gf_ctx* ctx, int impl, int emu_pairs) {
  if (!ctx) return GF_ERR_BAD_ARG;
  if ((impl != 0 && impl != 80 && impl != 128 && impl != 160) ||
      (emu_pairs != -1 && emu_pairs != 0 && emu_pairs != 2 && emu_pairs != 4 && emu_pairs != 6))
    return GF_ERR_BAD_ARG;
  ctx->tuning.attn_impl = impl;
  ctx->tuning.attn_emu = emu_pairs;
  return 0;
}

extern "C" int gf_ctx_set_gemm_raster(gf_ctx* ctx, int group_m) {
  if (!ctx || group_m < 0 || group_m > 1024) return GF_ERR_BAD_ARG;
  ctx->tuning.gemm_group_m = group_m;
  return 0;
}

extern "C" int gf_ctx_set_gemm_tile(gf_ctx* ctx, int bn) {
  if (!ctx || (bn != 0 && bn != 224 && bn != 256)) return GF_ERR_BAD_ARG;
  ctx->tuning.gemm_bn = bn;
  return 0;
}

extern "C" int gf_ctx_set_conv(gf_ctx* ctx, int impl) {
  if (!ctx || impl < 0 || impl > 2) return GF_ERR_BAD_ARG;
  ctx->tuning.conv_impl = impl;
  return 0;
}

extern "C" int gf_ctx_stats(gf_ctx* ctx, long long* tmap_entries, long long* tmap_hits, long long* tmap_misses) {
  if (!ctx) return GF_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (tmap_entries) *tmap_entries = (long long)ctx->tmaps.size();
  if (tmap_hits) *tmap_hits = ctx->hits;
  if (tmap_misses) *tmap_misses = ctx->misses;
  return 0;
}

extern "C" int gf_abi_version(void) { return GF_ABI_VERSION; }
extern "C" int gf_device_sms(void) { return gf::gf_num_sms(); }
