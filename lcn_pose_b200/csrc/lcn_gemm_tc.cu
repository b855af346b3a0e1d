// bf16 tensor-core kernels of the LCN mid layers for sm_100a: tcgen05.mma with TMEM accumulators,
// operands staged by 1-D bulk TMA (cp.async.bulk) through an mbarrier pipeline.
//
//   lcn_tc_gemm  : Y[128-row tile, group of <=4 output chunks] = sum over the input chunks that have a
//                  nonzero 64x64 block into the group of  A[tile, chunk] (128x64, K-major SW128)  x
//                  Wp(panel of present blocks) -- all-zero joint-pair blocks of the mask are never
//                  loaded nor multiplied (network/models_att.py:576-586 multiplies them densely).
//                  Forward: + bias, BatchNorm (mean, M2) partials per column (models_att.py:588-612),
//                  transposed (dgrad): + residual gradient addend.
//   lcn_tc_wgrad : dWm block(i,j) = A[:, i]^T dZ[:, j] for the nonzero blocks only, K = batch rows,
//                  both operands MN-major straight out of the tile-major activation layout.
//
// Data layout contracts: activations are tile-major SW128 (lcn_internal.cuh, lcn_off<bf16>), packed
// weights are per-64x64-block SW128 K-major images written by k_pack_mid (lcn_kernels.cu) in the order
// the panels are consumed, so every operand is one contiguous bulk copy -- no tensor maps needed.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..5 = epilogue (TMEM -> registers -> swizzled smem tile -> bulk store).  Two CTAs are
// co-resident per SM (256 TMEM columns and ~97 KB smem each) so one CTA's epilogue overlaps the other's
// main loop.
#include <stdlib.h>

#include "lcn_internal.cuh"

#define TC_G 4                      // max output chunks (of 64 columns) per CTA -> 256 TMEM columns
#define TC_STAGES 2
#define TC_A_BYTES 16384            // 128 rows x 64 bf16
#define TC_B_BYTES 8192             // one 64x64 bf16 block
#define TC_STAGE_BYTES (TC_A_BYTES + TC_G * TC_B_BYTES)
#define TC_THREADS 192

struct TcParams {
  uint32_t kmask[LCN_J];            // K-side joint -> bitmask of N-side joints with a block
  int FC, NC, n_groups, P;
  int bn_group, gstride;
  int transposed;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) {   // watchdog: turn a protocol bug into an error, not a hang
      printf("lcn_tc: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread = TMEM lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout=2 (SW128)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, M=128, N; a_major/b_major: 0 = K-major, 1 = MN-major
__device__ __forceinline__ uint32_t umma_idesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void group_range(int g, int NC, int n_groups, int* oc0, int* G) {
  int base = NC / n_groups, extra = NC % n_groups;
  *G = base + (g < extra ? 1 : 0);
  *oc0 = g * base + min(g, extra);
}

// present N-side chunks of the group for K-side chunk kc: bit q set <=> block (kc -> oc0+q) exists
__device__ __forceinline__ uint32_t present_bits(const TcParams& p, int kc, int oc0, int G) {
  uint32_t km = p.kmask[kc / p.FC], bits = 0;
  for (int q = 0; q < G; ++q)
    if ((km >> ((oc0 + q) / p.FC)) & 1u) bits |= 1u << q;
  return bits;
}
// index (in 64x64 blocks) of block (kc -> oc) inside the packed weight buffer of one layer
__device__ __forceinline__ int panel_slot(const TcParams& p, int kc, int oc) {
  int ka = kc / p.FC, hk = kc % p.FC, nb = oc / p.FC, hn = oc % p.FC;
  int base = 0;
  for (int q = 0; q < ka; ++q) base += __popc(p.kmask[q]);
  uint32_t km = p.kmask[ka];
  return p.FC * p.FC * base + hk * (__popc(km) * p.FC) + __popc(km & ((1u << nb) - 1u)) * p.FC + hn;
}

// ---------------------------------------------------------------------------------------------
// forward / dgrad GEMM
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS) k_tc_gemm(const __nv_bfloat16* __restrict__ A,
                                                        const __nv_bfloat16* __restrict__ Wp,
                                                        const float* __restrict__ bias,
                                                        const __nv_bfloat16* __restrict__ addend,
                                                        __nv_bfloat16* __restrict__ Y, float* __restrict__ part,
                                                        TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_STAGES + 1];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[TC_G * 64];

  // 1024-byte aligned operand area (SWIZZLE_128B atoms are 1024 B)
  uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x, tile = blockIdx.y;
  int oc0, G;
  group_range(g, p.NC, p.n_groups, &oc0, &G);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_STAGES]), tfull = smem_u32(&bars[2 * TC_STAGES]);
  const uint32_t tmem_cols = G <= 1 ? 64u : (G == 2 ? 128u : 256u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), tmem_cols);
  if (!p.transposed)
    for (int c = threadIdx.x; c < G * 64; c += TC_THREADS) bias_s[c] = bias[oc0 * 64 + c];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int kc = 0; kc < p.NC; ++kc) {
        uint32_t bits = present_bits(p, kc, oc0, G);
        if (!bits) continue;
        int s = it % TC_STAGES;
        uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        int cnt = __popc(bits);
        int first = oc0 + (__ffs(bits) - 1);
        uint32_t sa = sbase + s * TC_STAGE_BYTES;
        mbar_expect_tx(full0 + 8 * s, TC_A_BYTES + cnt * TC_B_BYTES);
        bulk_g2s(sa, A + ((size_t)tile * p.NC + kc) * 8192, TC_A_BYTES, full0 + 8 * s);
        bulk_g2s(sa + TC_A_BYTES, Wp + (size_t)panel_slot(p, kc, first) * 4096, cnt * TC_B_BYTES, full0 + 8 * s);
        ++it;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int it = 0;
      uint32_t written = 0;
      for (int kc = 0; kc < p.NC; ++kc) {
        uint32_t bits = present_bits(p, kc, oc0, G);
        if (!bits) continue;
        int s = it % TC_STAGES;
        uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        uint32_t sa = sbase + s * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
        int rank = 0, q = 0;
        while (q < G) {
          if (!((bits >> q) & 1u)) { ++q; continue; }
          // maximal run of present chunks with the same accumulate state
          uint32_t acc = (written >> q) & 1u;
          int len = 1;
          while (q + len < G && ((bits >> (q + len)) & 1u) && (((written >> (q + len)) & 1u) == acc)) ++len;
          uint32_t idesc = umma_idesc(64 * len, 0, 0);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            uint64_t ad = umma_desc_sw128(sa + k4 * 32, 16, 1024);
            uint64_t bd = umma_desc_sw128(sb + rank * TC_B_BYTES + k4 * 32, 16, 1024);
            umma_f16(tmem_base + q * 64, ad, bd, idesc, (acc | (uint32_t)(k4 > 0)));
          }
          written |= ((1u << len) - 1u) << q;
          rank += len;
          q += len;
        }
        umma_commit(empty0 + 8 * s);     // frees the smem stage when these MMAs have read it
        ++it;
      }
      umma_commit(tfull);                // accumulators complete
    }
  } else {
    // ===== epilogue: 4 warps, warp's TMEM lane quarter = warp id % 4 =====
    const int lq = warp & 3;
    const int row = lq * 32 + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    uint32_t written = 0;
    for (int kc = 0; kc < p.NC; ++kc) written |= present_bits(p, kc, oc0, G);
    for (int q = 0; q < G; ++q) {
      uint8_t* tile_s = sgen + q * TC_A_BYTES;
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        if ((written >> q) & 1u) {
          tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + q * 64 + h * 32, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0u;
        }
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (!p.transposed) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] += bias_s[q * 64 + h * 32 + i];
        } else if (addend != nullptr) {
          const uint8_t* arow = reinterpret_cast<const uint8_t*>(addend) +
                                (((size_t)tile * p.NC + oc0 + q) * 128 + row) * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 u = *reinterpret_cast<const uint4*>(arow + (((h * 4 + c) ^ (row & 7)) << 4));
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 t = __bfloat1622float2(hp[e]);
              f[c * 8 + 2 * e] += t.x;
              f[c * 8 + 2 * e + 1] += t.y;
            }
          }
        }
        // pack to bf16 and store the row's 4 sixteen-byte chunks into the swizzled staging tile
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          __nv_bfloat162 b0 = __floats2bfloat162_rn(f[c * 8 + 0], f[c * 8 + 1]);
          __nv_bfloat162 b1 = __floats2bfloat162_rn(f[c * 8 + 2], f[c * 8 + 3]);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(f[c * 8 + 4], f[c * 8 + 5]);
          __nv_bfloat162 b3 = __floats2bfloat162_rn(f[c * 8 + 6], f[c * 8 + 7]);
          u.x = *reinterpret_cast<uint32_t*>(&b0);
          u.y = *reinterpret_cast<uint32_t*>(&b1);
          u.z = *reinterpret_cast<uint32_t*>(&b2);
          u.w = *reinterpret_cast<uint32_t*>(&b3);
          *reinterpret_cast<uint4*>(tile_s + row * 128 + (((h * 4 + c) ^ (row & 7)) << 4)) = u;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();                                   // generic smem writes -> bulk-store (async proxy) reads
    asm volatile("bar.sync 1, 128;" ::: "memory");         // the 4 epilogue warps
    if (warp == 2 && lane == 0) {
      for (int q = 0; q < G; ++q)
        bulk_s2g(Y + ((size_t)tile * p.NC + oc0 + q) * 8192, sbase + q * TC_A_BYTES, TC_A_BYTES);
      bulk_commit_wait();
    }
    if (!p.transposed && part != nullptr) {
      // BatchNorm partials of this tile: per column (mean, M2) over the valid rows, from the bf16 values
      int q = warp - 2;
      if (q < G) {
        int tig = tile % (p.gstride / LCN_TILE);
        int nvalid = min(LCN_TILE, p.bn_group - tig * LCN_TILE);
        const uint8_t* tile_s = sgen + q * TC_A_BYTES;
        float sh0 = 0.f, sh1 = 0.f, s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
        for (int r = 0; r < nvalid; ++r) {
          uint32_t w = *reinterpret_cast<const uint32_t*>(tile_s + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
          float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
          if (r == 0) { sh0 = t.x; sh1 = t.y; }
          float d0 = t.x - sh0, d1 = t.y - sh1;
          s1a += d0; s2a = fmaf(d0, d0, s2a);
          s1b += d1; s2b = fmaf(d1, d1, s2b);
        }
        float n = (float)nvalid;
        size_t o = ((size_t)tile * p.P + (oc0 + q) * 64 + lane * 2) * 2;
        float4 out = make_float4(sh0 + s1a / n, fmaxf(s2a - s1a * s1a / n, 0.f), sh1 + s1b / n,
                                 fmaxf(s2b - s1b * s1b / n, 0.f));
        *reinterpret_cast<float4*>(part + o) = out;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool lcn_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LCN_DISABLE_TC");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

int lcn_tc_gemm(const lcn_model* m, const WsLayout& lay, int mid_index, int transposed, const __nv_bfloat16* A,
                const char* wpacked, const float* bias, const __nv_bfloat16* addend, __nv_bfloat16* Y, float* part,
                cudaStream_t st) {
  (void)mid_index;
  TcParams p;
  for (int a = 0; a < LCN_J; ++a) p.kmask[a] = transposed ? m->sup.col[a] : m->sup.row[a];
  p.FC = m->FC;
  p.NC = LCN_J * m->FC;
  p.n_groups = (p.NC + TC_G - 1) / TC_G;
  p.P = m->P;
  p.bn_group = lay.bn_group;
  p.gstride = lay.gstride;
  p.transposed = transposed;
  size_t smem = (size_t)TC_STAGES * TC_STAGE_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    LCN_CHECK_CUDA(cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  k_tc_gemm<<<dim3(p.n_groups, lay.tiles), TC_THREADS, smem, st>>>(
      A, reinterpret_cast<const __nv_bfloat16*>(wpacked), bias, addend, Y, part, p);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

// ---------------------------------------------------------------------------------------------
// weight gradient of the nonzero blocks: dWm[(i,hi) chunk, (j,ho) chunk] += A[rows, ic]^T dZ[rows, oc]
//   D[M=64 input channels, N=64*len output channels] = sum_k A_op[m][k] * B_op[n][k], k = batch row.
//   Both operands are MN-major SW128 tiles exactly as stored in HBM (K = the 128 rows of a tile).
//   unit = (input chunk, group of <=4 of its present output chunks); grid.y splits the batch rows.
//   M=64 accumulators use the half-subpartition TMEM layout: row m -> lane 32*(m/16) + m%16.
// ---------------------------------------------------------------------------------------------
#define TCW_STAGE_BYTES ((1 + TC_G) * TC_A_BYTES)
#define TCW_PITCH 260   // floats per staged output row (1040 B: 16-byte aligned, bank-conflict free)

struct TcwParams {
  uint32_t row[LCN_J];     // outputs of input joint i
  int FC, NC, P, tiles, tiles_per_cta;
};

__device__ __forceinline__ void bulk_reduce_add_f32(float* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src),
               "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(TC_THREADS) k_tc_wgrad(const __nv_bfloat16* __restrict__ A,
                                                         const __nv_bfloat16* __restrict__ dZ,
                                                         float* __restrict__ dW, TcwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_STAGES + 1];
  __shared__ uint32_t tmem_base_s;
  uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // decode the unit: input chunk ic and its gi-th group of present output chunks
  int ic = 0, gi = 0;
  {
    int u = blockIdx.x;
    for (ic = 0; ic < p.NC; ++ic) {
      int cnt = __popc(p.row[ic / p.FC]) * p.FC;
      int ng = (cnt + TC_G - 1) / TC_G;
      if (u < ng) { gi = u; break; }
      u -= ng;
    }
  }
  int ocs[TC_G];
  int len = 0;
  {
    uint32_t bits = p.row[ic / p.FC];
    int e = 0;
    for (int j = 0; j < LCN_J; ++j) {
      if (!((bits >> j) & 1u)) continue;
      for (int ho = 0; ho < p.FC; ++ho, ++e)
        if (e >= gi * TC_G && e < gi * TC_G + TC_G) ocs[len++] = j * p.FC + ho;
    }
  }
  const int t0 = blockIdx.y * p.tiles_per_cta;
  const int t1 = min(t0 + p.tiles_per_cta, p.tiles);
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TC_STAGES]), tfull = smem_u32(&bars[2 * TC_STAGES]);
  const uint32_t tmem_cols = len <= 1 ? 64u : (len == 2 ? 128u : 256u);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        int s = it % TC_STAGES;
        uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(empty0 + 8 * s, ph ^ 1u);
        uint32_t sa = sbase + s * TCW_STAGE_BYTES;
        mbar_expect_tx(full0 + 8 * s, (1 + len) * TC_A_BYTES);
        bulk_g2s(sa, A + ((size_t)t * p.NC + ic) * 8192, TC_A_BYTES, full0 + 8 * s);
        for (int q = 0; q < len; ++q)
          bulk_g2s(sa + (1 + q) * TC_A_BYTES, dZ + ((size_t)t * p.NC + ocs[q]) * 8192, TC_A_BYTES, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // M = 64, N = 64*len, both operands MN-major
      uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)((64 * len) >> 3) << 17) |
                       ((uint32_t)(64 >> 4) << 24);
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        int s = it % TC_STAGES;
        uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        tc_fence_after();
        uint32_t sa = sbase + s * TCW_STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
        for (int k16 = 0; k16 < 8; ++k16) {
          uint64_t ad = umma_desc_sw128(sa + k16 * 2048, TC_A_BYTES, 1024);
          uint64_t bd = umma_desc_sw128(sb + k16 * 2048, TC_A_BYTES, 1024);
          umma_f16(tmem_base, ad, bd, idesc, (uint32_t)(it > 0 || k16 > 0));
        }
        umma_commit(empty0 + 8 * s);
      }
      umma_commit(tfull);
    }
  } else {
    const int lq = warp & 3;
    mbar_wait(tfull, 0);
    tc_fence_after();
    float* out_s = reinterpret_cast<float*>(sgen);
    const int m = lq * 16 + lane;                    // valid for lane < 16
    for (int q = 0; q < len; ++q)
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + q * 64 + h * 32, v);
        if (lane < 16) {
          float* dst = out_s + m * TCW_PITCH + q * 64 + h * 32;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(dst + c * 4) = make_uint4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
        }
      }
    tc_fence_before();
    fence_proxy_async();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (lane < 16) {
      for (int q = 0; q < len; ++q)
        bulk_reduce_add_f32(dW + (size_t)(ic * 64 + m) * p.P + ocs[q] * 64, smem_u32(out_s + m * TCW_PITCH + q * 64), 256);
      bulk_commit_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

int lcn_tc_wgrad(const lcn_model* m, const WsLayout& lay, const __nv_bfloat16* A, const __nv_bfloat16* dZ, float* dW,
                 cudaStream_t st) {
  TcwParams p;
  int units = 0;
  for (int i = 0; i < LCN_J; ++i) {
    p.row[i] = m->sup.row[i];
    int cnt = __builtin_popcount(m->sup.row[i]) * m->FC;
    units += m->FC * ((cnt + TC_G - 1) / TC_G);
  }
  p.FC = m->FC;
  p.NC = LCN_J * m->FC;
  p.P = m->P;
  p.tiles = lay.tiles;
  p.tiles_per_cta = lay.tiles >= 16 ? 4 : (lay.tiles >= 4 ? 2 : 1);
  int splits = (lay.tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  size_t smem = (size_t)TC_STAGES * TCW_STAGE_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    LCN_CHECK_CUDA(cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  k_tc_wgrad<<<dim3(units, splits), TC_THREADS, smem, st>>>(A, dZ, dW, p);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}
