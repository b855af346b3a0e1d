// Microbenchmark 3: period of the producer <-> MMA-issuer mbarrier handshake of a S-stage pipeline with no
// payload (no loads, no MMAs), by signalling variant.
//   variant 0: consumer releases a stage with tcgen05.commit (as in the kernels)
//   variant 1: consumer releases with a plain mbarrier.arrive
//   variant 2: like 0, with N=128 MMAs (4 per stage) in the consumer
//   variant 3: like 0, producer issues one 16 KB bulk copy per stage (L2 resident)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lcn_pose_b200/csrc -o handshake handshake.cu
#include <cuda_runtime.h>
#include "lcn_tc_ptx.cuh"

__device__ __forceinline__ void mbar_wait_plain(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_plain(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int VARIANT>
__global__ void __launch_bounds__(128) k_hs(int S, int iters, const uint8_t* src, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  __shared__ uint32_t tmem_base_s;
  uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_s), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  long long t0 = clock64();
  if (warp == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1u;
      mbar_wait_plain(empty0 + 8 * s, ph ^ 1u);
      if (elect_one()) {
        if (VARIANT == 3) {
          mbar_expect_tx(full0 + 8 * s, 16384);
          bulk_g2s(sbase + s * 16384, src + (size_t)((it * 37 + blockIdx.x * 11) % 1024) * 16384, 16384, full0 + 8 * s);
        } else {
          mbar_expect_tx(full0 + 8 * s, 0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint64_t desc_hi = (uint64_t)((1024u >> 4) & 0x3FFF) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16);
    const uint32_t idesc = umma_idesc(128, 0, 0);
    for (int it = 0; it < iters; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1u;
      mbar_wait_plain(full0 + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        if (VARIANT == 2) {
          const uint64_t ad = desc_hi | (uint64_t)(((sbase + s * 16384) >> 4) & 0x3FFF);
          const uint64_t bd = desc_hi | (uint64_t)(((sbase + 65536) >> 4) & 0x3FFF);
          for (int k = 0; k < 4; ++k) umma_f16(tm, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
        if (VARIANT == 1) mbar_arrive_plain(empty0 + 8 * s); else umma_commit(empty0 + 8 * s);
      }
      __syncwarp();
    }
    // drain
    for (int s = 0; s < S; ++s) { /* last phases complete on their own */ }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tm, 512);
}

template <int V>
void run(int g, int S, const uint8_t* src, long long* d, long long* h) {
  const int iters = 2000;
  cudaFuncSetAttribute(k_hs<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k_hs<V><<<g, 128, 90 * 1024>>>(S, iters, src, d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  cudaMemcpy(h, d, g * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < g; ++i) mean += (double)h[i]; mean /= g;
  printf("%d,%d,%d,%.1f\n", g, V, S, mean / iters);
}

int main() {
  long long* d; cudaMalloc(&d, 4096 * sizeof(long long));
  uint8_t* src; cudaMalloc(&src, 16 << 20); cudaMemset(src, 0, 16 << 20);
  long long h[4096];
  printf("grid,variant,stages,cycles_per_iteration\n");
  for (int g : {1, 148})
    for (int S : {1, 2, 3, 4}) {
      run<0>(g, S, src, d, h);
      run<1>(g, S, src, d, h);
      run<2>(g, S, src, d, h);
      run<3>(g, S, src, d, h);
    }
  return 0;
}
