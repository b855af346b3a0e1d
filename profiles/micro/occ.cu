// Which kernel property limits CTAs/SM?  Occupancy queries for: plain kernel, +tcgen05.alloc, +12 KB of parameters, +cluster barrier.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct Big { int v[2992]; };
struct Small { int v[16]; };
__global__ void __launch_bounds__(192) k_plain(Small p, int* out) { extern __shared__ int sm[]; sm[threadIdx.x] = p.v[0]; __syncthreads(); out[threadIdx.x] = sm[(threadIdx.x + 1) % 192]; }
__global__ void __launch_bounds__(192) k_big(const __grid_constant__ Big p, int* out) { extern __shared__ int sm[]; sm[threadIdx.x] = p.v[threadIdx.x]; __syncthreads(); out[threadIdx.x] = sm[(threadIdx.x + 1) % 192]; }
__global__ void __launch_bounds__(192) k_tmem(Small p, int* out) {
  extern __shared__ int sm[];
  __shared__ uint32_t base;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  sm[threadIdx.x] = p.v[0] + base;
  __syncthreads();
  out[threadIdx.x] = sm[(threadIdx.x + 1) % 192];
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(256) : "memory");
}
__global__ void __launch_bounds__(192) k_cluster(Small p, int* out) {
  extern __shared__ int sm[];
  sm[threadIdx.x] = p.v[0];
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
  out[threadIdx.x] = sm[(threadIdx.x + 1) % 192];
}
template <typename K> void q(const char* name, K k) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 114688);
  for (size_t s : {16384, 65536, 110000}) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, 192, s);
    printf("%s smem %zu -> %d CTAs/SM (%s)\n", name, s, n, cudaGetErrorString(e));
  }
}
int main() {
  q("plain", k_plain); q("bigparam", k_big); q("tmem", k_tmem); q("cluster", k_cluster);
  return 0;
}
