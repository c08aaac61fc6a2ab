// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, cta_group::1, M = 128, K = 16) as a function
// of N, for A in shared memory (SS) and A in tensor memory (TS).  One CTA per SM, one issuing thread, ITER
// back-to-back MMAs into the same accumulator, one commit at the end.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I goal_force_b200/csrc -o mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include "gf_ptx.cuh"
using namespace gf;

template <int N, bool kTS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 32768, bar = base + 32768 + 65536, tptr = bar + 8;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (elect_one()) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc<1>(tptr, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc = idesc_bf16(128, N, 0, 0);
    constexpr uint64_t dk = smem_desc_base(1024, 16);
    const uint64_t da = smem_desc(dk, a_smem), db = smem_desc(dk, b_smem);
    // warm-up
    for (int i = 0; i < 16; ++i) {
      if (kTS) umma_ts<1>(tmem, tmem + 256, db, idesc, 1u); else umma_ss<1>(tmem, da, db, idesc, 1u);
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (kTS) umma_ts<1>(tmem, tmem + 256, db, idesc, 1u); else umma_ss<1>(tmem, da, db, idesc, 1u);
    }
    tc_commit(bar);
    mbar_wait(bar, 1);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

template <int N, bool kTS> void run(long long* d, int iters) {
  auto kern = k<N, kTS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  kern<<<148, 128, 110 * 1024>>>(d, iters);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = double(h) / iters;
  printf("%s M=128 N=%3d K=16: %7.2f clk/MMA  -> %6.1f MAC/clk/SM  (linear-in-N model %5.1f clk)  err=%s\n", kTS ? "TS" : "SS", N,
         per, 128.0 * N * 16 / per, 128.0 * N / 256, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int iters = 4000;
  run<16, false>(d, iters); run<32, false>(d, iters); run<48, false>(d, iters); run<64, false>(d, iters);
  run<80, false>(d, iters); run<96, false>(d, iters); run<112, false>(d, iters); run<128, false>(d, iters);
  run<160, false>(d, iters); run<192, false>(d, iters); run<256, false>(d, iters);
  run<64, true>(d, iters); run<80, true>(d, iters); run<128, true>(d, iters); run<256, true>(d, iters);
  return 0;
}
