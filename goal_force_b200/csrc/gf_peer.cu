// gf_peer.cu -- peer-memory plumbing for the fused Ulysses exchange: device buffers that the other GPUs of the node
// map through CUDA IPC, and a flag barrier between the ranks that runs on the compute stream.
//
// The reference delegates the head<->sequence exchange to xfuser's all-to-all over NCCL
// (diffsynth/distributed/xdit_context_parallel.py:121-126).  Here the producing kernels store straight into the
// consumer GPU's buffers over NVLink (gf_qkv_rmsnorm_rope_scatter_bf16, gf_attention_scatter_bf16); what remains of
// the collective is "everybody has finished writing into my buffer", which is this barrier.
#include "gf_api_internal.h"

namespace gf {

struct PeerPtrs {
  void* p[GF_MAX_PEERS];
};

// flags of rank r: uint32 slot[s] = last epoch rank s has signalled to r.
// Thread s < n: (1) make this GPU's earlier writes visible system-wide, (2) release-store `epoch` into rank s's
// slot[rank], (3) spin until rank s has stored `epoch` (or later) into my slot[s].
__global__ void peer_barrier_kernel(PeerPtrs flags, int n, int rank, unsigned epoch) {
  const int s = threadIdx.x;
  if (s >= n) return;
  __threadfence_system();
  unsigned* remote = reinterpret_cast<unsigned*>(flags.p[s]) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const unsigned* mine = reinterpret_cast<const unsigned*>(flags.p[rank]) + s;
  unsigned v;
  unsigned long long spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (++spins > (1ull << 26)) __trap();      // seconds: a missing peer traps instead of hanging the GPU
  } while ((int)(v - epoch) < 0);
}

}  // namespace gf

using namespace gf;

extern "C" int gf_peer_alloc(void** ptr, long long bytes) {
  if (!ptr || bytes <= 0) return GF_ERR_BAD_ARG;
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(*ptr, 0, (size_t)bytes);
}

extern "C" int gf_peer_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : GF_ERR_BAD_ARG; }

extern "C" int gf_peer_export(void* ptr, void* handle64) {
  if (!ptr || !handle64) return GF_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == GF_PEER_HANDLE_BYTES, "IPC handle size");
  return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
}

extern "C" int gf_peer_import(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return GF_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int gf_peer_unimport(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : GF_ERR_BAD_ARG; }

extern "C" int gf_peer_barrier(void* const* flag_peers, int n_peers, int rank, unsigned epoch, void* stream) {
  if (!flag_peers || n_peers < 1 || n_peers > GF_MAX_PEERS || rank < 0 || rank >= n_peers) return GF_ERR_BAD_ARG;
  PeerPtrs t{};
  for (int i = 0; i < n_peers; ++i) {
    if (!flag_peers[i]) return GF_ERR_BAD_ARG;
    t.p[i] = flag_peers[i];
  }
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, n_peers, rank, epoch);
  return (int)cudaGetLastError();
}
