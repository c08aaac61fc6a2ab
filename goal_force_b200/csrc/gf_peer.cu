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

// flags of rank r (uint32 words): slot[s], s < GF_MAX_PEERS = last epoch rank s has signalled to r;
// slot[GF_PEER_EPOCH_SLOT] = number of barriers rank r has passed (touched only by r's own barrier kernel, so the
// epoch lives on the device and a captured CUDA graph can replay the barrier).
// Thread s < n: (1) make this GPU's earlier writes visible system-wide, (2) release-store the new epoch into rank s's
// slot[rank], (3) wait until rank s has stored that epoch (or a later one) into my slot[s].
// The wait is bounded by wall-clock time (%globaltimer), not by iterations: a slow peer (host-side skew while
// another rank converts an expert or runs a callback) is waited for; a peer that never shows up within timeout_ns
// makes the kernel record `1 + s` in *status (host-mapped memory the host polls without a sync) and return, so the
// failure surfaces as a Python exception at the next check instead of a trap that destroys every rank's context.
constexpr int GF_PEER_EPOCH_SLOT = 16;

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_barrier_kernel(PeerPtrs flags, int n, int rank, unsigned long long timeout_ns,
                                    unsigned* status) {
  const int s = threadIdx.x;
  unsigned* own = reinterpret_cast<unsigned*>(flags.p[rank]);
  const unsigned epoch = own[GF_PEER_EPOCH_SLOT] + 1u;
  __syncwarp();
  if (s < n) {
    __threadfence_system();
    unsigned* remote = reinterpret_cast<unsigned*>(flags.p[s]) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
    const unsigned* mine = own + s;
    unsigned v;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if ((++spins & 0x3FFu) == 0) {             // look at the clock every 1024 polls
        const unsigned long long now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > timeout_ns) {
          if (status) {
            *reinterpret_cast<volatile unsigned*>(status) = 1u + (unsigned)s;
            __threadfence_system();
          }
          break;
        }
      }
    }
  }
  __syncwarp();
  if (s == 0) own[GF_PEER_EPOCH_SLOT] = epoch;
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

extern "C" int gf_peer_barrier(void* const* flag_peers, int n_peers, int rank, long long timeout_ms, unsigned* status,
                               void* stream) {
  if (!flag_peers || n_peers < 1 || n_peers > GF_MAX_PEERS || rank < 0 || rank >= n_peers) return GF_ERR_BAD_ARG;
  PeerPtrs t{};
  for (int i = 0; i < n_peers; ++i) {
    if (!flag_peers[i]) return GF_ERR_BAD_ARG;
    t.p[i] = flag_peers[i];
  }
  const unsigned long long timeout_ns = (timeout_ms > 0 ? (unsigned long long)timeout_ms : 60000ull) * 1000000ull;
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, n_peers, rank, timeout_ns, status);
  return (int)cudaGetLastError();
}

// Host-mapped status word for gf_peer_barrier: *host_ptr is readable by the CPU at any time (no stream sync), *dev_ptr
// is what the kernels write.  Zero at allocation.
extern "C" int gf_peer_status_alloc(unsigned** host_ptr, unsigned** dev_ptr) {
  if (!host_ptr || !dev_ptr) return GF_ERR_BAD_ARG;
  void* h = nullptr;
  cudaError_t e = cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) return (int)e;
  memset(h, 0, 64);
  void* d = nullptr;
  e = cudaHostGetDevicePointer(&d, h, 0);
  if (e != cudaSuccess) { cudaFreeHost(h); return (int)e; }
  *host_ptr = reinterpret_cast<unsigned*>(h);
  *dev_ptr = reinterpret_cast<unsigned*>(d);
  return 0;
}

extern "C" int gf_peer_status_free(unsigned* host_ptr) {
  return host_ptr ? (int)cudaFreeHost(host_ptr) : GF_ERR_BAD_ARG;
}
