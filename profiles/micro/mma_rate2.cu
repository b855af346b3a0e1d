// Microbenchmark 2: tcgen05.mma (bf16, M=128, cta_group::1) issued from warp-uniform code (uniform registers,
// no per-instruction election loop): cycles per instruction for fixed issue patterns.
//   pattern 0: one accumulator, every instruction              (long chain)
//   pattern 1: switch accumulator every 4 instructions, same A/B tiles
//   pattern 2: switch accumulator every instruction
//   pattern 3: like 1, and the A tile alternates with the accumulator (two row tiles)
//   pattern 4: like 1 with 4 accumulators
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lcn_pose_b200/csrc -o mma_rate2 mma_rate2.cu
#include <cuda_runtime.h>
#include "lcn_tc_ptx.cuh"

template <int PATTERN>
__global__ void __launch_bounds__(128) k_mma(int N, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (32768 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (sbase - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  if (warp == 1) {
    const uint64_t desc_hi = (uint64_t)((1024u >> 4) & 0x3FFF) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16);
    const uint64_t ad = desc_hi | (uint64_t)((sbase >> 4) & 0x3FFF), bd = desc_hi | (uint64_t)(((sbase + 32768) >> 4) & 0x3FFF);
    const uint32_t idesc = umma_idesc(N, 0, 0);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t d = tm;
          uint64_t a = ad + 2 * (j & 3);
          if (PATTERN == 1) d = tm + ((j >> 2) & 1) * 256;
          if (PATTERN == 2) d = tm + (j & 1) * 256;
          if (PATTERN == 3) { d = tm + ((j >> 2) & 1) * 256; a += ((j >> 2) & 1) * (16384 >> 4); }
          if (PATTERN == 4) d = tm + ((j >> 2) & 3) * 128;
          umma_f16(d, a, bd + 2 * (j & 3), idesc, 1u);
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int P>
void run(int g, int N, long long* d, long long* h) {
  const int iters = 128;
  cudaFuncSetAttribute(k_mma<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  k_mma<P><<<g, 128, 70 * 1024>>>(N, iters, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  cudaMemcpy(h, d, g * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < g; ++i) mean += (double)h[i]; mean /= g;
  printf("%d,%d,%d,%.1f,%d\n", g, N, P, mean / (iters * 16), N / 2);
}

int main() {
  long long* d; cudaMalloc(&d, 4096 * sizeof(long long));
  long long h[4096];
  printf("grid,N,pattern,cycles_per_mma,nominal\n");
  for (int g : {1, 148})
    for (int N : {64, 128, 192, 256}) {
      run<0>(g, N, d, h);
      run<1>(g, N, d, h);
      run<2>(g, N, d, h);
      run<3>(g, N, d, h);
      if (N <= 128) run<4>(g, N, d, h);
    }
  return 0;
}
