// Microbenchmark: does the tcgen05.mma floor for small N come from the issuing thread or from the tensor pipe /
// shared-memory operand read?  NW warps issue independent MMA streams (own accumulator columns) concurrently.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../segmentation-networks-benchmark_b200/csrc/sm100_ptx.cuh"
using namespace snb;

template <int N, int NW>
__global__ void __launch_bounds__(32 * (NW + 1), 1) mma_dual_kernel(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == NW) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc05_fence_before();
  __syncthreads();
  tc05_fence_after();
  const uint32_t tmem = tptr;
  long long t0 = clock64();
  if (warp < NW && elect_one()) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t a0 = smem_u32(smem + warp * 16384), b0 = smem_u32(smem + 65536 + warp * 8192);
    const uint64_t ad = make_kmajor_desc<128>(a0, 8 * 128), bd = make_kmajor_desc<128>(b0, 8 * 128);
    for (int i = 0; i < iters; ++i) umma_bf16_ss(ad, bd, tmem + warp * 128, idesc, 1u);
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    out[blockIdx.x * 4 + warp] = clock64() - t0;
  }
  tc05_fence_before();
  __syncthreads();
  if (warp == NW) { tc05_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int NW>
void run() {
  const int grid = 148, iters = 4096, smem = 100 * 1024;
  long long* d; cudaMalloc(&d, grid * 4 * sizeof(long long)); cudaMemset(d, 0, grid * 4 * sizeof(long long));
  cudaFuncSetAttribute(mma_dual_kernel<N, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) mma_dual_kernel<N, NW><<<grid, 32 * (NW + 1), smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148 * 4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < grid * 4; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d issuing warps=%d: %.1f cycles per MMA per warp -> %.1f cycles per MMA aggregate (ideal %d) [%s]\n", N, NW,
         mx / iters, mx / iters / NW, N / 2, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<32, 1>(); run<32, 2>(); run<32, 4>();
  run<64, 1>(); run<64, 2>(); run<64, 4>();
  run<128, 1>(); run<128, 2>();
  return 0;
}
