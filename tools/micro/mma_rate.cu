// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS) as a function of N, operands resident in smem.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../segmentation-networks-benchmark_b200/csrc/sm100_ptx.cuh"
using namespace snb;

template <int M, int N, int SBO_ROWS, bool FIXED>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters, int distinct) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc05_fence_before();
  __syncthreads();
  tc05_fence_after();
  const uint32_t tmem = tptr;
  if (warp == 0 && elect_one()) {
    constexpr uint32_t idesc = make_idesc_bf16(M, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = FIXED ? 0 : (i % distinct);   // rotate K slices / stages like a real main loop
      const uint64_t ad = make_kmajor_desc<128>(a0 + (s >> 2) * 8192 * 0 + (s & 3) * 32, SBO_ROWS * 128);
      const uint64_t bd = make_kmajor_desc<128>(b0 + (s & 3) * 32, 8 * 128);
      umma_bf16_ss(ad, bd, tmem, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc05_fence_before();
  __syncthreads();
  if (warp == 1) { tc05_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int M, int N, int SBO_ROWS, bool FIXED>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, grid * sizeof(long long));
  const int iters = 4096;
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel<M, N, SBO_ROWS, FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<M, N, SBO_ROWS, FIXED><<<grid, 128, smem>>>(d, iters, 4);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mx = 0, mn = 1e30;
  for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("%-20s grid=%3d M=%3d N=%3d fixed=%d: %.1f .. %.1f cycles/MMA (ideal %d)  [%s]\n", name, grid, M, N, (int)FIXED,
         mn / iters, mx / iters, M * N / 256, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {148}) {
    run<128, 32, 8, true>("M128 fixed desc", grid);
    run<128, 64, 8, true>("M128 fixed desc", grid);
    run<128, 128, 8, true>("M128 fixed desc", grid);
    run<128, 256, 8, true>("M128 fixed desc", grid);
    run<128, 256, 8, false>("M128", grid);
    run<64, 32, 8, true>("M64", grid);
    run<64, 64, 8, true>("M64", grid);
    run<64, 128, 8, true>("M64", grid);
    run<64, 256, 8, true>("M64", grid);
    run<64, 256, 8, false>("M64", grid);
  }
  return 0;
}
