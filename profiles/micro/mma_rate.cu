// Microbenchmark: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M=128, cta_group::1) issue patterns.
//   cycles per instruction for N in {64,128,192,256}, K16 steps walking a K=64 SW128 K-major tile, when
//   (a) all instructions accumulate into ONE TMEM tile, (b) consecutive instructions rotate over R tiles,
//   (c) chains of C instructions per tile before switching.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lcn_pose_b200/csrc -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include "lcn_tc_ptx.cuh"

__global__ void __launch_bounds__(128) k_mma(int N, int R, int C, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (sbase - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_base_s), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tm = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint64_t desc_hi = (uint64_t)((1024u >> 4) & 0x3FFF) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16);
    uint64_t ad = desc_hi | (uint64_t)((sbase >> 4) & 0x3FFF), bd = desc_hi | (uint64_t)(((sbase + 16384) >> 4) & 0x3FFF);
    uint32_t idesc = umma_idesc(N, 0, 0);
    long long t0 = clock64();
    int n = 0;
    for (int it = 0; it < iters; ++it)
      for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c, ++n)
          umma_f16(tm + r * (512 / R), ad + 2 * (c & 3), bd + 2 * (c & 3), idesc, 1u);
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 4096 * sizeof(long long));
  long long h[4096];
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  printf("grid,N,R(tiles rotated),C(chain per tile),cycles_per_mma,nominal\n");
  int grids[] = {1, 148};
  for (int g : grids)
    for (int N : {64, 128, 192, 256})
      for (int R : {1, 2, 4})
        for (int C : {1, 4, 16, 64}) {
          if (R * N > 512) continue;
          int iters = 2048 / (R * C);
          k_mma<<<g, 128, 50 * 1024>>>(N, R, C, iters, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          cudaMemcpy(h, d, g * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
          double mean = 0; for (int i = 0; i < g; ++i) mean += (double)h[2 * i] / h[2 * i + 1]; mean /= g;
          printf("%d,%d,%d,%d,%.1f,%d\n", g, N, R, C, mean, N / 2);
        }
  return 0;
}
