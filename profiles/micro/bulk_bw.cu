// Microbenchmark: L2 -> shared-memory throughput per SM of (a) 1-D bulk copies (cp.async.bulk), by copy size and
// pipeline depth, and (b) 2-D tiled TMA (cp.async.bulk.tensor) with a 64x128 SW128 box (16 KB), on an L2-resident
// source.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}

// each CTA: one thread keeps `depth` stages of `stage_bytes` in flight; a stage = stage_bytes/copy_bytes copies
__global__ void k_bulk(const uint8_t* src, size_t src_bytes, int copy_bytes, int stage_bytes, int depth, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[16];
  uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < depth; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    uint32_t n_stage_src = (uint32_t)(src_bytes / stage_bytes);
    uint32_t rng = blockIdx.x * 2654435761u + 12345u;
    for (int it = 0; it < iters + depth; ++it) {
      int s = it % depth;
      uint32_t bar = smem_u32(&bars[s]);
      if (it >= depth) mbar_wait(bar, ((it / depth) - 1) & 1u);
      if (it < iters) {
        rng = rng * 1664525u + 1013904223u;
        const uint8_t* p = src + (size_t)((rng >> 8) % (uint32_t)n_stage_src) * stage_bytes;
        mbar_expect_tx(bar, stage_bytes);
        for (int o = 0; o < stage_bytes; o += copy_bytes) bulk_g2s(sb + s * stage_bytes + o, p + o, copy_bytes, bar);
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

__global__ void k_tma2d(const __grid_constant__ CUtensorMap map, int rows_total, int boxes_per_stage, int depth, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[16];
  uint32_t sb = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < depth; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    uint32_t rng = blockIdx.x * 2654435761u + 12345u;
    const int stage_bytes = boxes_per_stage * 16384;
    for (int it = 0; it < iters + depth; ++it) {
      int s = it % depth;
      uint32_t bar = smem_u32(&bars[s]);
      if (it >= depth) mbar_wait(bar, ((it / depth) - 1) & 1u);
      if (it < iters) {
        mbar_expect_tx(bar, stage_bytes);
        for (int b = 0; b < boxes_per_stage; ++b) {
          rng = rng * 1664525u + 1013904223u;
          int y = (rng % (rows_total / 128)) * 128;
          int x = ((rng >> 20) % 17) * 64;
          tma_2d(sb + s * stage_bytes + b * 16384, &map, x, y, bar);
        }
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t src_bytes = 48u << 20;      // L2 resident
  uint8_t* src;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  long long* cyc;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  long long h[1024];
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k_tma2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 400;
  printf("kind,copy_bytes,stage_bytes,depth,ctas_per_sm,B_per_clk_per_SM,chip_B_per_clk,GBps\n");
  int cfgs[][4] = {{16384, 16384, 1, 1}, {16384, 49152, 1, 1}, {8192, 8192, 1, 1}, {2048, 2048, 1, 1}, {16384, 49152, 2, 1}, {16384, 49152, 4, 1},{16384, 16384, 4, 1}, {16384, 16384, 8, 1}, {16384, 49152, 3, 1}, {8192, 49152, 3, 1}, {4096, 49152, 3, 1},
                   {2048, 49152, 3, 1}, {16384, 32768, 3, 2}, {16384, 32768, 5, 1}, {4096, 16384, 8, 1}, {1024, 16384, 8, 1}, {16384, 16384, 11, 1}};
  for (auto& c : cfgs) {
    int copy = c[0], stage = c[1], depth = c[2], occ = c[3];
    size_t smem = (size_t)stage * depth + 1024;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      k_bulk<<<sms * occ, 32, smem>>>(src, src_bytes, copy, stage, depth, iters, cyc);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaMemcpy(h, cyc, sms * occ * sizeof(long long), cudaMemcpyDeviceToHost);
      double mean = 0; for (int i = 0; i < sms * occ; ++i) mean += h[i]; mean /= sms * occ;
      if (rep) printf("bulk1d,%d,%d,%d,%d,%.1f,%.0f,%.0f,cycles_per_stage=%.0f\n", copy, stage, depth, occ, (double)stage * iters * occ / mean,
                      (double)stage * iters * occ / mean * sms, (double)stage * iters * occ * sms / (ms * 1e6), mean / iters);
    }
  }
  // 2-D tiled TMA over a [rows][1088] bf16 row-major matrix, box 64 cols x 128 rows, SWIZZLE_128B
  {
    const int cols = 1088, rows = (int)(src_bytes / (cols * 2)) / 128 * 128;
    CUtensorMap map;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, gdim, gstr, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", (int)r); return 1; }
    int cfg2[][3] = {{1, 4, 1}, {1, 8, 1}, {3, 3, 1}, {2, 3, 2}, {1, 11, 1}};
    for (auto& c : cfg2) {
      int bps = c[0], depth = c[1], occ = c[2];
      size_t smem = (size_t)bps * 16384 * depth + 1024;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_tma2d<<<sms * occ, 32, smem>>>(map, rows, bps, depth, iters, cyc);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, sms * occ * sizeof(long long), cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < sms * occ; ++i) mean += h[i]; mean /= sms * occ;
        double bytes = (double)bps * 16384 * iters * occ;
        if (rep) printf("tma2d,%d,%d,%d,%d,%.1f,%.0f,%.0f\n", 16384, bps * 16384, depth, occ, bytes / mean, bytes / mean * sms, bytes * sms / (ms * 1e6));
      }
    }
  }
  return 0;
}
