// Fused inference kernel of the whole LCN layer stack (cgcnn._inference_lcn, network/models_att.py:707-775,
// as base_model.predict runs it, :79-132) for sm_100a: one persistent thread-block CLUSTER per BatchNorm
// group of `batch_size` poses.
//
// Why a cluster per BN group: the reference's BatchNormalization uses batch statistics at inference too
// (SURVEY 9-Q2), so every layer needs a reduction over the batch_size x 17 rows of its group before its
// activation can be applied.  batch_size <= 256 rows are 1-2 row tiles of 128; the NS CTAs of a cluster split
// the 17 output joints (64-channel chunks) of those tiles between them, keep their slice of the layer's
// pre-BN output Z as fp32 accumulators in TMEM (<= 512 columns), exchange per-column (mean, M2) through
// distributed shared memory, and apply BN + LeakyReLU(0.2) + residual straight out of TMEM.  Activations
// never touch HBM: a layer's output goes to a per-cluster L2-resident scratch (bulk store) from which the
// next layer's A operand is streamed back with bulk TMA, one multicast copy per K chunk for the whole cluster.
//
//   layer 0   : X[rows, 17*in_F] fp32 -> bf16 hi + lo tiles built in shared memory, K padded to 64 (x = hi + lo
//               keeps ~16 mantissa bits of the 2D input), one dense 64x64 block per output chunk
//   mid layers: block-sparse X * (W.M): only joint pairs inside the mask support are loaded and multiplied
//   head      : 17*F -> 51 (N padded to 64) + xy skip connection, fp32 rows to the caller's output
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2.. = epilogue (EW = 4 or 8 warps:
// one or two per TMEM lane quarter; with two, each takes one 32-column half of a chunk).
//
// Two BN groups per SM.  A kernel that allocates TMEM is limited to ONE resident CTA per SM (measured:
// profiles/micro/occ.cu), and the per-layer chain MMA -> statistics -> exchange -> BN epilogue -> store -> barrier of
// one group is serial.  So a CTA hosts NV = 2 "virtual CTAs" (its own producer, MMA and epilogue warps, pipeline
// stages, mbarriers and half of the TMEM columns each) that belong to two different virtual clusters working on two
// different BN groups: one group's exchange / epilogue / cluster synchronisation overlaps the other group's MMA main
// loop.  The virtual clusters synchronise with mbarriers (remote arrive, release/acquire at cluster scope); the
// hardware cluster barrier is only used at kernel start and end.  256 TMEM columns per virtual CTA = 2 output chunks
// x 2 row tiles, hence 9 CTAs per 256-row group (non-portable cluster size).  The (EW = 8, NV = 1) instantiation is the
// one-group-per-SM variant (up to 384 columns, 6 CTAs per group).
#include <stdlib.h>
#include <string.h>

#include "lcn_internal.cuh"
#include "lcn_tc_ptx.cuh"

#define ST_MAX_STAGES 4
#define ST_A_BYTES 16384
#define ST_B_BYTES 8192
#define ST_MAX_COLS 384        // output columns per CTA: chunks*64
#define ST_MAX_RUN 4           // chunks per MMA (N <= 256)
#define ST_MAX_NS 9            // CTAs per cluster

// optional timeline (clock64) of cluster 0 / CTA 0 over its first group: enabled with -DLCN_TC_PROFILE
#ifdef LCN_TC_PROFILE
__device__ unsigned long long g_st_prof[512];
#define ST_STAMP(i) do { if (blockIdx.x == 0 && vc == 0 && grp_cnt == 1) g_st_prof[(i)] = clock64(); } while (0)
extern "C" int lcn_debug_read_stack_prof(unsigned long long* h_out) {
  return cudaMemcpyFromSymbol(h_out, g_st_prof, sizeof(g_st_prof)) == cudaSuccess ? 0 : -2;
}
#else
#define ST_STAMP(i) do {} while (0)
#endif

struct StackParams {
  const float* x;
  float* out;
  const float* params;
  const __nv_bfloat16* wf16;    // first layer: [17 chunks][64 n][64 k] SW128 K-major
  const __nv_bfloat16* wp16f;   // mid layers:  [n_mid][nnz blocks] (k_pack_mid forward order)
  const __nv_bfloat16* wl16f;   // head:        [17 K chunks][64 n (51 used)][64 k]
  __nv_bfloat16* scratch;       // [clusters][2][TPG][17][128x64] activation ping-pong, tile-major SW128
  __nv_bfloat16* taps;          // optional parity tap: every A_l, [n_bn][tiles][17][128x64] tile-major SW128
  int64_t taps_stride;          // elements per layer of `taps`
  int64_t b_off[LCN_MAX_LIN], gamma_off[LCN_MAX_LIN], beta_off[LCN_MAX_LIN];
  int8_t res[LCN_MAX_LIN];      // 1: layer output += A_{l-2}   (models_att.py:704)
  uint32_t kmask[LCN_J];        // input joint -> bitmask of output joints with a block
  // MMA program of a mid layer per cluster rank: per K chunk the runs of present output chunks with equal
  // accumulate state -> (TMEM column offset, B descriptor offset, instruction descriptor, accumulate)
  uint4 prog[ST_MAX_NS][LCN_J][ST_MAX_RUN];
  uint8_t prog_cnt[ST_MAX_NS][LCN_J];
  uint8_t sch_cnt[ST_MAX_NS][LCN_J];    // per rank and K chunk: number of present blocks of the rank's output range ...
  uint16_t sch_slot[ST_MAX_NS][LCN_J];  // ... and the slot of the first one in the packed layer (k_pack_mid forward order)
  int8_t oc_start[ST_MAX_NS + 1];       // output chunks of CTA r: [oc_start[r], oc_start[r+1])  (balanced by block count)
  int n_lin, in_F, TPG, stages, mc, nnz, tmem_cols;
  int stage_bytes;              // TPG*16 KB of A + Gmax*8 KB of weight blocks
  int vc_bytes;                 // shared memory of one virtual CTA (stages + statistics), multiple of 1024
  int cols;                     // Gmax*64: output columns per CTA
  int dbg;                      // LCN_STACK_DBG experiments: 1 = no MMAs, 2 = no A loads, 4 = no B loads (wrong results)
  int64_t n_rows;
  int bn_group, n_groups;
};

__device__ __forceinline__ void st_range(int r, int NC, int n, int* c0, int* G) {
  int base = NC / n, extra = NC % n;
  *G = base + (r < extra ? 1 : 0);
  *c0 = r * base + min(r, extra);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// named barrier of the epilogue warps of virtual CTA `vc`
template <int EW>
__device__ __forceinline__ void epi_bar(int vc) { asm volatile("bar.sync %0, %1;" ::"r"(vc + 1), "n"(EW * 32) : "memory"); }
// arrive (release, cluster scope) on the same mbarrier of CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  const uint32_t remote = mapa_shared(bar, rank);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a local mbarrier whose arrivals come from other CTAs (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) {
      printf("lcn_stack: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// column sums over the 32 rows (lanes) of a warp for 32 columns held one row per lane:
// recursive halving, 31 shuffles; on return lane L holds the total of column L in a[0].
__device__ __forceinline__ float warp_colsum32(float* a, int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    bool up = lane & 16;
    float send = up ? a[i] : a[i + 16];
    float keep = up ? a[i + 16] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bool up = lane & 8;
    float send = up ? a[i] : a[i + 8];
    float keep = up ? a[i + 8] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bool up = lane & 4;
    float send = up ? a[i] : a[i + 4];
    float keep = up ? a[i + 4] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    bool up = lane & 2;
    float send = up ? a[i] : a[i + 2];
    float keep = up ? a[i + 2] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    bool up = lane & 1;
    float send = up ? a[0] : a[1];
    float keep = up ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return a[0];
}

template <int EW, int NV>
__global__ void __launch_bounds__(NV * (64 + 32 * EW), 1) k_lcn_stack(const __grid_constant__ StackParams p) {
  constexpr int ST_EPI_THREADS = 32 * EW;
  constexpr int VT = 64 + 32 * EW;       // threads of one virtual CTA
  constexpr int HSTEP = EW / 4;          // epilogue warps per TMEM lane quarter
  constexpr int NBAR = 2 * ST_MAX_STAGES + 5;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars_all[NV * NBAR];
  __shared__ uint32_t tmem_base_s;

  const int vc = __shfl_sync(0xffffffffu, (int)(threadIdx.x / VT), 0);     // virtual CTA of this warp
  const int vtid = (int)threadIdx.x - vc * VT;
  uint64_t* bars = bars_all + vc * NBAR;
  const uint32_t sbase = ((smem_u32(smem_raw) + 1023u) & ~1023u) + (uint32_t)vc * (uint32_t)p.vc_bytes;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int ST_STAGE_BYTES = p.stage_bytes;
  const int NCOL = p.cols;
  // dedicated (never aliased with the pipeline stages: stat_all is written by the other CTAs of the cluster at any time)
  float2* stat_all = reinterpret_cast<float2*>(sgen + (size_t)p.stages * ST_STAGE_BYTES);   // per (joint, channel): (mean, M2)
  float* colA = reinterpret_cast<float*>(stat_all + LCN_J * 64);
  float* colB = colA + NCOL;
  // epilogue scratch inside the (then idle) pipeline stages: per-warp 32x33 transpose tiles, then the per-lane-quarter
  // column sums (fixed-order sum: deterministic)
  // -- placed at the END of the stage area: the front holds the output tiles (and the prefetched residual tiles)
  float* scr0 = reinterpret_cast<float*>(sgen + (size_t)p.stages * ST_STAGE_BYTES) - (EW * (32 * 33) + 8 * NCOL);
  float* colS = scr0 + EW * (32 * 33);
  float* colQ = colS + 4 * NCOL;
  const int warp = __shfl_sync(0xffffffffu, vtid >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank(), NS = cluster_nctarank();
  const uint32_t cid = cluster_id_x() * NV + vc, ncl = cluster_nid_x() * NV;     // virtual cluster id / count
  const int TPG = p.TPG, S = p.stages;
  const int Kin = LCN_J * p.in_F;
  const uint32_t B_OFF = (uint32_t)TPG * ST_A_BYTES;
  const uint16_t all_mask = (uint16_t)((1u << NS) - 1u);
  const int oc0 = p.oc_start[rank], G = p.oc_start[rank + 1] - oc0;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[ST_MAX_STAGES]);
  const uint32_t tfull = smem_u32(&bars[2 * ST_MAX_STAGES]), xfull = smem_u32(&bars[2 * ST_MAX_STAGES + 1]);
  // virtual-cluster barriers: one arrival per CTA of the cluster.  xs1: every CTA has written its BN statistics into
  // every stat_all;  xs2: every CTA has finished the layer (A_l stored, TMEM and pipeline stages free)
  const uint32_t xs1 = smem_u32(&bars[2 * ST_MAX_STAGES + 2]), xs2 = smem_u32(&bars[2 * ST_MAX_STAGES + 3]);
  const uint32_t xres = smem_u32(&bars[2 * ST_MAX_STAGES + 4]);      // residual tiles of the current layer have landed

  if (vtid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, p.mc ? NS : 1u);
    }
    mbar_init(tfull, 1);
    mbar_init(xfull, ST_EPI_THREADS);
    mbar_init(xs1, NS);
    mbar_init(xs2, NS);
    mbar_init(xres, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1 && vc == 0) tmem_alloc(smem_u32(&tmem_base_s), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                       // barrier inits are visible to every CTA of the cluster
  const uint32_t tmem_base = tmem_base_s + (uint32_t)vc * (uint32_t)(p.tmem_cols / NV);

  __nv_bfloat16* buf0 = p.scratch + (size_t)cid * 2 * TPG * LCN_J * 8192;
  __nv_bfloat16* buf1 = buf0 + (size_t)TPG * LCN_J * 8192;
  const float inv_n = 1.f / (float)p.bn_group;
  // pipeline position: stage index and phase parity of the next stage use (same sequence in every role and CTA);
  // kept incrementally -- a runtime modulo / division per iteration costs ~200 cycles of the issuing warps
  int st_s = 0;
  uint32_t st_ph = 0;
  uint32_t layer_cnt = 0, grp_cnt = 0, bn_cnt = 0, res_cnt = 0;

  for (int g = (int)cid; g < p.n_groups; g += (int)ncl, ++grp_cnt) {
    for (int l = 0; l < p.n_lin; ++l, ++layer_cnt) {
      const bool first = l == 0, head = l == p.n_lin - 1;
      const __nv_bfloat16* ain = ((l - 1) & 1) ? buf1 : buf0;
      __nv_bfloat16* aout = (l & 1) ? buf1 : buf0;
      int t_lo = 0, t_hi = TPG, Gl = G;
      if (head) {
        Gl = 1;
        if ((int)rank < TPG) { t_lo = (int)rank; t_hi = t_lo + 1; } else { t_hi = 0; }
      }
      const int n_it = first ? 1 : LCN_J;

      if (warp == 0) {
        // ===================== TMA producer =====================
        // warp-uniform like the MMA issuer: addresses come from kernel parameters / loop counters, one elected lane issues
        // every CTA of the virtual cluster has finished the previous layer: A_{l-1} is in L2, all stages are free
        if (layer_cnt > 0) mbar_wait_cluster(xs2, (layer_cnt - 1) & 1u);
        if (first) {
          const int s = st_s;
          const uint32_t ph = st_ph;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(full0 + 8 * s, (uint32_t)G * ST_B_BYTES);
            bulk_g2s(sbase + s * ST_STAGE_BYTES + B_OFF, p.wf16 + (size_t)oc0 * 4096, (uint32_t)G * ST_B_BYTES, full0 + 8 * s);
          }
          __syncwarp();
        } else {
          fence_proxy_async_all();          // the cluster's bulk stores of A_{l-1} precede these bulk loads
          const __nv_bfloat16* wl = head ? p.wl16f : p.wp16f + (size_t)(l - 1) * p.nnz * 4096;
          int s = st_s;
          uint32_t ph = st_ph;
          uint32_t issuer = 0;              // multicast A loads rotate over the CTAs of the cluster
          for (int kc = 0; kc < LCN_J; ++kc) {
            int cnt, slot;
            if (head) {
              cnt = t_hi > t_lo ? 1 : 0;
              slot = kc;
            } else {
              cnt = p.sch_cnt[rank][kc];
              slot = p.sch_slot[rank][kc];
            }
            if (p.dbg & 4) cnt = 0;
            const uint32_t sa = sbase + s * ST_STAGE_BYTES;
            const uint32_t fb = full0 + 8 * s;
            mbar_wait(empty0 + 8 * s, ph ^ 1u);
            if (elect_one()) {
              if (p.dbg & 2) {
                mbar_expect_tx(fb, (uint32_t)cnt * ST_B_BYTES);
              } else if (p.mc) {
                mbar_expect_tx(fb, (uint32_t)TPG * ST_A_BYTES + (uint32_t)cnt * ST_B_BYTES);
                if (issuer == rank)
                  for (int t = 0; t < TPG; ++t)
                    bulk_g2s_mc(sa + t * ST_A_BYTES, ain + ((size_t)t * LCN_J + kc) * 8192, ST_A_BYTES, fb, all_mask);
              } else {
                mbar_expect_tx(fb, (uint32_t)(t_hi - t_lo) * ST_A_BYTES + (uint32_t)cnt * ST_B_BYTES);
                for (int t = t_lo; t < t_hi; ++t)
                  bulk_g2s(sa + t * ST_A_BYTES, ain + ((size_t)t * LCN_J + kc) * 8192, ST_A_BYTES, fb);
              }
              if (cnt) bulk_g2s(sa + B_OFF, wl + (size_t)slot * 4096, (uint32_t)cnt * ST_B_BYTES, fb);
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
            if (++issuer == NS) issuer = 0;
          }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs this code (warp-uniform control flow and operands, so the descriptors live in uniform
        // registers); one elected lane issues.  The per-K-chunk MMA program comes from the kernel parameters
        // (constant bank -> uniform loads).  Per (run, tile) the four K16 steps go back to back into one accumulator.
        const uint64_t desc_hi = (uint64_t)((1024u >> 4) & 0x3FFF) << 32 | (1ull << 46) | (2ull << 61) | (1ull << 16);
        if (first) {
          const int s = st_s;
          const uint32_t ph = st_ph;
          mbar_wait(xfull, grp_cnt & 1u);
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t sb = sbase + s * ST_STAGE_BYTES + B_OFF;
          const int s1 = s + 1 >= S ? s + 1 - S : s + 1, s2 = s + 2 >= S ? s + 2 - S : s + 2;
          const uint32_t xh = sbase + s1 * ST_STAGE_BYTES, xl = sbase + s2 * ST_STAGE_BYTES;
          const int nk16 = (Kin + 15) >> 4;
          if (elect_one()) {
            for (int t = 0; t < TPG; ++t)
              for (int q0 = 0; q0 < G; q0 += ST_MAX_RUN) {
                const int len = min(ST_MAX_RUN, G - q0);
                const uint32_t idesc = umma_idesc(64 * len, 0, 0);
                const uint32_t d = tmem_base + (uint32_t)(t * G + q0) * 64;
                const uint64_t bd = desc_hi | (uint64_t)(((sb + q0 * ST_B_BYTES) >> 4) & 0x3FFF);
                const uint64_t ah = desc_hi | (uint64_t)(((xh + t * ST_A_BYTES) >> 4) & 0x3FFF);
                const uint64_t al = desc_hi | (uint64_t)(((xl + t * ST_A_BYTES) >> 4) & 0x3FFF);
                for (int k = 0; k < nk16; ++k) umma_f16(d, ah + 2 * k, bd + 2 * k, idesc, (uint32_t)(k > 0));
                for (int k = 0; k < nk16; ++k) umma_f16(d, al + 2 * k, bd + 2 * k, idesc, 1u);
              }
            if (p.mc) umma_commit_mc(empty0 + 8 * s, all_mask); else umma_commit(empty0 + 8 * s);
          }
          __syncwarp();
        } else {
#ifdef LCN_TC_PROFILE
          long long acc_wait = 0, acc_issue = 0, acc_commit = 0;
#endif
          int s = st_s;
          uint32_t ph = st_ph;
          for (int kc = 0; kc < LCN_J; ++kc) {
#ifdef LCN_TC_PROFILE
            long long tw0 = clock64();
#endif
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
#ifdef LCN_TC_PROFILE
            acc_wait += clock64() - tw0;
#endif
            if (lane == 0) {
              if (kc == 0) ST_STAMP(16 * l + 8);
              if (kc == 8) ST_STAMP(16 * l + 9);
              if (kc == 16) ST_STAMP(16 * l + 10);
            }
            const uint32_t sa = sbase + s * ST_STAGE_BYTES;
            const uint64_t ad0 = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);
            const uint64_t bd0 = desc_hi | (uint64_t)(((sa + B_OFF) >> 4) & 0x3FFF);
#ifdef LCN_TC_PROFILE
            long long tm0 = clock64(), tm1 = 0;
#endif
            if (elect_one()) {
              if (!(p.dbg & 1)) {
                if (head) {
                  if (t_hi > t_lo) {
                    const uint32_t idesc = umma_idesc(64, 0, 0);
                    const uint64_t ad = ad0 + (uint64_t)(t_lo * (ST_A_BYTES >> 4));
                    // two partial accumulators (K16 steps 0-1 / 2-3 of every K chunk), summed in the epilogue
                    umma_f16(tmem_base, ad, bd0, idesc, (uint32_t)(kc > 0));
                    umma_f16(tmem_base, ad + 2, bd0 + 2, idesc, 1u);
                    umma_f16(tmem_base + 64, ad + 4, bd0 + 4, idesc, (uint32_t)(kc > 0));
                    umma_f16(tmem_base + 64, ad + 6, bd0 + 6, idesc, 1u);
                  }
                } else {
                  const int n = p.prog_cnt[rank][kc];
                  for (int r = 0; r < n; ++r) {
                    const uint4 e = p.prog[rank][kc][r];
                    const uint64_t bd = bd0 + e.y;
                    for (int t = 0; t < TPG; ++t) {
                      const uint64_t ad = ad0 + (uint64_t)(t * (ST_A_BYTES >> 4));
                      const uint32_t d = tmem_base + (uint32_t)(t * G) * 64 + e.x;
                      umma_f16(d, ad, bd, e.z, e.w);
                      umma_f16(d, ad + 2, bd + 2, e.z, 1u);
                      umma_f16(d, ad + 4, bd + 4, e.z, 1u);
                      umma_f16(d, ad + 6, bd + 6, e.z, 1u);
                    }
                  }
                }
              }
#ifdef LCN_TC_PROFILE
              tm1 = clock64();
#endif
              if (p.mc) umma_commit_mc(empty0 + 8 * s, all_mask); else umma_commit(empty0 + 8 * s);
#ifdef LCN_TC_PROFILE
              acc_issue += tm1 - tm0;
              acc_commit += clock64() - tm1;
#endif
            }
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
          }
#ifdef LCN_TC_PROFILE
          if (blockIdx.x == 0 && vc == 0 && grp_cnt == 1) {
            if (lane == 0) g_st_prof[16 * l + 11] = acc_wait;
            if (acc_issue) { g_st_prof[16 * l + 13] = acc_issue; g_st_prof[16 * l + 14] = acc_commit; }
          }
#endif
        }
        if (elect_one()) umma_commit(tfull);
        __syncwarp();
      } else {
        // ===================== epilogue: 8 warps =====================
        const int e = vtid - 64;
        // TMEM lane quarter of a warp = its hardware warp index % 4 (not the index inside the virtual CTA)
        const int lq = (int)(threadIdx.x >> 5) & 3, hset0 = (warp - 2) >> 2;
        const int row = lq * 32 + lane;
        const uint32_t tlane = (uint32_t)(lq * 32) << 16;
        if (first) {
          // bf16 hi / lo tiles of the group's 2D input in the A areas of the two idle pipeline stages
          const int s1 = st_s + 1 >= S ? st_s + 1 - S : st_s + 1, s2 = st_s + 2 >= S ? st_s + 2 - S : st_s + 2;
          uint8_t* xh = sgen + s1 * ST_STAGE_BYTES;
          uint8_t* xl = sgen + s2 * ST_STAGE_BYTES;
          for (int t = e >> 7; t < TPG; t += ST_EPI_THREADS / 128) {
            const int r = e & 127;
            const int rin = t * LCN_TILE + r;
            const int64_t src = (int64_t)g * p.bn_group + rin;
            const bool ok = rin < p.bn_group && src < p.n_rows;
            const float* xr = p.x + (ok ? src : 0) * Kin;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = c * 8 + 2 * j;
                float v0 = (ok && k < Kin) ? __ldg(xr + k) : 0.f;
                float v1 = (ok && k + 1 < Kin) ? __ldg(xr + k + 1) : 0.f;
                __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                float2 hf = __bfloat1622float2(h);
                __nv_bfloat162 lo = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
                hw[j] = *reinterpret_cast<uint32_t*>(&h);
                lw[j] = *reinterpret_cast<uint32_t*>(&lo);
              }
              const uint32_t o = (uint32_t)t * ST_A_BYTES + r * 128 + ((c ^ (r & 7)) << 4);
              *reinterpret_cast<uint4*>(xh + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(xl + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
          fence_proxy_async();
          mbar_arrive(xfull);
        }
        if (e == 0) ST_STAMP(16 * l + 0);
        // one warp polls the accumulator-ready barrier, the other seven block at a hardware barrier: 256 threads
        // spinning on mbarrier.try_wait slow down every barrier operation of the main loop (measured: 4x)
        if (warp == 2) mbar_wait(tfull, layer_cnt & 1u);
        epi_bar<EW>(vc);
        tc_fence_after();
        if (e == 0) ST_STAMP(16 * l + 1);
        if (!head) {
          const bool has_res = p.res[l] != 0;
          constexpr int NH = 2 / HSTEP;          // work items of a warp: (unit u, 32-column half)
          const int U = TPG * G;
          if (has_res && warp == 2) {
            // residual A_{l-2} (the buffer this layer overwrites, written by this CTA's own bulk stores): its tiles are
            // bulk-copied straight into the places where the output tiles will be staged -- every thread later reads
            // the 64 B it overwrites.  The copies land while pass 1 and the statistics exchange run.
            if (elect_one()) {
              mbar_expect_tx(xres, (uint32_t)U * ST_A_BYTES);
              for (int u = 0; u < U; ++u) {
                const int t = u / G, q = u - t * G;
                bulk_g2s(sbase + (uint32_t)u * ST_A_BYTES, aout + ((size_t)t * LCN_J + oc0 + q) * 8192, ST_A_BYTES, xres);
              }
            }
            __syncwarp();
          }
          // ---- pass 1: per-column sum / sum of squares of the accumulators over the group's valid rows ----
          // (32x32 transpose through a per-warp scratch in the idle pipeline stages: lane = row writes, lane = column sums)
          float* scr = scr0 + (warp - 2) * (32 * 33);
          for (int q = 0; q < G; ++q)
            for (int hset = hset0; hset < 2; hset += HSTEP) {
              float cs = 0.f, cq = 0.f;
              for (int t = 0; t < TPG; ++t) {
                uint32_t v[32];
                tmem_ld32(tmem_base + tlane + (uint32_t)(t * G + q) * 64 + hset * 32, v);
                const bool valid = t * LCN_TILE + row < p.bn_group;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; ++i) scr[lane * 33 + i] = valid ? __uint_as_float(v[i]) : 0.f;
                __syncwarp();
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                  const float a = scr[r * 33 + lane];
                  cs += a;
                  cq = fmaf(a, a, cq);
                }
              }
              colS[lq * NCOL + q * 64 + hset * 32 + lane] = cs;
              colQ[lq * NCOL + q * 64 + hset * 32 + lane] = cq;
            }
          epi_bar<EW>(vc);
          if (e == 0) ST_STAMP(16 * l + 2);
          const float* bias = p.params + p.b_off[l];
          for (int c = e; c < G * 64; c += ST_EPI_THREADS) {
            // Z = acc + bias: mean = bias + S/n, M2 = Q - S^2/n  (shift by the bias keeps the cancellation small)
            const float Ssum = (colS[c] + colS[NCOL + c]) + (colS[2 * NCOL + c] + colS[3 * NCOL + c]);
            const float Qsum = (colQ[c] + colQ[NCOL + c]) + (colQ[2 * NCOL + c] + colQ[3 * NCOL + c]);
            const float ma = Ssum * inv_n;
            const float mean_c = __ldg(bias + oc0 * 64 + c) + ma;
            const float m2 = fmaxf(Qsum - Ssum * ma, 0.f);
            const uint32_t la = smem_u32(&stat_all[oc0 * 64 + c]);
            for (uint32_t r = 0; r < NS; ++r) st_cluster_f32x2(mapa_shared(la, r), mean_c, m2);
          }
          epi_bar<EW>(vc);
          if (warp == 2) {
            if (lane < (int)NS) mbar_arrive_cluster(xs1, (uint32_t)lane);
            mbar_wait_cluster(xs1, bn_cnt & 1u);                 // #1: every CTA holds the 17 x 64 column statistics
          }
          ++bn_cnt;
          epi_bar<EW>(vc);
          if (e == 0) ST_STAMP(16 * l + 3);
          for (int c = e; c < G * 64; c += ST_EPI_THREADS) {
            // BatchNormalization over batch x joints, biased variance, eps 1e-3 (models_att.py:599-607)
            const int f = c & 63;
            float msum = 0.f;
            for (int j = 0; j < LCN_J; ++j) msum += stat_all[j * 64 + f].x;
            const float mean = msum * (1.f / LCN_J);
            float m2 = 0.f;
            for (int j = 0; j < LCN_J; ++j) {
              const float2 sj = stat_all[j * 64 + f];
              const float d = sj.x - mean;
              m2 += sj.y + (float)p.bn_group * d * d;
            }
            const float var = m2 * inv_n * (1.f / LCN_J);
            const float sc = __ldg(p.params + p.gamma_off[l] + f) * rsqrtf(var + LCN_BN_EPS);
            colA[c] = sc;
            colB[c] = __ldg(bias + oc0 * 64 + c) * sc + __ldg(p.params + p.beta_off[l] + f) - mean * sc;
          }
          if (has_res) {
            if (warp == 2) mbar_wait(xres, res_cnt & 1u);          // the residual tiles have landed
            ++res_cnt;
          }
          epi_bar<EW>(vc);
          // ---- pass 2: BN + LeakyReLU (+ residual) out of TMEM -> bf16 swizzled tiles in smem -> bulk store per tile ----
          for (int t = 0; t < TPG; ++t) {
            for (int it = t * G * NH; it < (t + 1) * G * NH; ++it) {
              const int u = it / NH, hset = hset0 + (it - u * NH) * HSTEP;
              const int q = u - t * G;
              uint32_t v[32];
              tmem_ld32_nowait(tmem_base + tlane + (uint32_t)u * 64 + hset * 32, v);
              tmem_ld_wait();
              float f[32];
              const float4* a4 = reinterpret_cast<const float4*>(colA + q * 64 + hset * 32);
              const float4* b4 = reinterpret_cast<const float4*>(colB + q * 64 + hset * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 av = a4[i], bv = b4[i];
                float y0 = fmaf(__uint_as_float(v[4 * i]), av.x, bv.x);
                float y1 = fmaf(__uint_as_float(v[4 * i + 1]), av.y, bv.y);
                float y2 = fmaf(__uint_as_float(v[4 * i + 2]), av.z, bv.z);
                float y3 = fmaf(__uint_as_float(v[4 * i + 3]), av.w, bv.w);
                f[4 * i] = fmaxf(y0, LCN_LRELU * y0);
                f[4 * i + 1] = fmaxf(y1, LCN_LRELU * y1);
                f[4 * i + 2] = fmaxf(y2, LCN_LRELU * y2);
                f[4 * i + 3] = fmaxf(y3, LCN_LRELU * y3);
              }
              uint8_t* tile_s = sgen + (size_t)u * ST_A_BYTES;
              if (has_res) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const uint4 rv = *reinterpret_cast<const uint4*>(tile_s + row * 128 + (((hset * 4 + c) ^ (row & 7)) << 4));
                  const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 r2 = __bfloat1622float2(hp[k]);
                    f[c * 8 + 2 * k] += r2.x;
                    f[c * 8 + 2 * k + 1] += r2.y;
                  }
                }
              }
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                __nv_bfloat162 b0 = __floats2bfloat162_rn(f[c * 8 + 0], f[c * 8 + 1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(f[c * 8 + 2], f[c * 8 + 3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[c * 8 + 4], f[c * 8 + 5]);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(f[c * 8 + 6], f[c * 8 + 7]);
                uint4 w;
                w.x = *reinterpret_cast<uint32_t*>(&b0);
                w.y = *reinterpret_cast<uint32_t*>(&b1);
                w.z = *reinterpret_cast<uint32_t*>(&b2);
                w.w = *reinterpret_cast<uint32_t*>(&b3);
                *reinterpret_cast<uint4*>(tile_s + row * 128 + (((hset * 4 + c) ^ (row & 7)) << 4)) = w;
              }
            }
            if (t == TPG - 1) tc_fence_before();
            fence_proxy_async();
            epi_bar<EW>(vc);
            if (warp == 2) {
              if (t == TPG - 1 && lane == 0) ST_STAMP(16 * l + 4);
              if (elect_one()) {
                for (int q = 0; q < G; ++q) {
                  bulk_s2g(aout + ((size_t)t * LCN_J + oc0 + q) * 8192, sbase + (uint32_t)(t * G + q) * ST_A_BYTES, ST_A_BYTES);
                  if (p.taps != nullptr)
                    bulk_s2g(p.taps + (size_t)l * p.taps_stride + (((size_t)g * TPG + t) * LCN_J + oc0 + q) * 8192,
                             sbase + (uint32_t)(t * G + q) * ST_A_BYTES, ST_A_BYTES);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (t == TPG - 1) {
                  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                  fence_proxy_async_all();
                  ST_STAMP(16 * l + 5);
                }
              }
              __syncwarp();
              if (t == TPG - 1 && lane < (int)NS) mbar_arrive_cluster(xs2, (uint32_t)lane);     // #2 (this CTA's part)
            }
          }
        } else if (t_hi > t_lo) {
          // ---- head: 51 valid columns + xy skip connection -> fp32 prediction rows (models_att.py:765-773) ----
          const int t = t_lo;
          const int rin = t * LCN_TILE + row;
          const int64_t src = (int64_t)g * p.bn_group + rin;
          const bool ok = rin < p.bn_group && src < p.n_rows;
          for (int hset = hset0; hset < 2; hset += HSTEP) {
            uint32_t v[32];
            tmem_ld32(tmem_base + tlane + hset * 32, v);
            {
              uint32_t w[32];
              tmem_ld32(tmem_base + tlane + 64u + hset * 32, w);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
            }
            if (ok) {
              const float* bias = p.params + p.b_off[l];
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int c = hset * 32 + i;
                if (c < 51) {
                  float val = __uint_as_float(v[i]) + __ldg(bias + c);
                  const int j = c / 3, cc = c - j * 3;
                  if (cc < 2) val += __ldg(p.x + src * Kin + j * p.in_F + cc);
                  p.out[src * 51 + c] = val;
                }
              }
            }
          }
          tc_fence_before();
        }
        if (head) {
          epi_bar<EW>(vc);
          if (warp == 2 && lane < (int)NS) mbar_arrive_cluster(xs2, (uint32_t)lane);           // #2 (head layer)
        }
        if (e == 0) ST_STAMP(16 * l + 6);
      }
      for (int a = 0; a < n_it; ++a)
        if (++st_s == S) { st_s = 0; st_ph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // no CTA exits while another one may still write into its shared memory / arrive on its barriers
  if (warp == 1 && vc == 0) tmem_dealloc(tmem_base_s, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool stack_verbose() {
  static const bool v = getenv("LCN_STACK_VERBOSE") != nullptr;      // diagnostics of the configuration search only
  return v;
}

// contiguous split of the 17 output joints over ns CTAs, at most gmax each, minimising the largest block count
static void stack_partition(const lcn_model* m, int ns, int gmax, int8_t* oc_start) {
  int col[LCN_J];
  for (int j = 0; j < LCN_J; ++j) col[j] = __builtin_popcount(m->sup.col[j]);
  int best = 1 << 30, sizes[10], cur[10];
  // depth-first over compositions (ns <= 8, gmax <= 6: a few thousand candidates)
  struct Rec {
    static void go(int k, int used, int ns, int gmax, const int* col, int* cur, int mx, int* best, int* sizes) {
      if (k == ns) {
        if (used == LCN_J && mx < *best) { *best = mx; for (int i = 0; i < ns; ++i) sizes[i] = cur[i]; }
        return;
      }
      for (int a = 1; a <= gmax && used + a <= LCN_J; ++a) {
        if ((ns - k - 1) * gmax < LCN_J - used - a) continue;
        int load = 0;
        for (int j = used; j < used + a; ++j) load += col[j];
        cur[k] = a;
        go(k + 1, used + a, ns, gmax, col, cur, load > mx ? load : mx, best, sizes);
      }
    }
  };
  Rec::go(0, 0, ns, gmax, col, cur, 0, &best, sizes);
  oc_start[0] = 0;
  for (int i = 0; i < ns; ++i) oc_start[i + 1] = (int8_t)(oc_start[i] + sizes[i]);
}

struct StackCfg {
  int ns;        // CTAs per cluster (= per BatchNorm group)
  int gmax;      // output chunks per CTA (upper bound)
  int stages;    // pipeline stages
  int mc;        // multicast the activation tiles over the cluster
  int nv;        // virtual CTAs (BN groups in flight) per CTA: 2 -> <EW=4, NV=2> kernel, 1 -> <EW=8, NV=1>
  size_t vc_bytes;
  int tmem_cols;
  int stage_bytes;
  size_t smem;   // dynamic shared memory per CTA
};

static size_t stack_vc_bytes(const StackCfg& c) {
  size_t b = (size_t)c.stages * c.stage_bytes + LCN_J * 64 * sizeof(float2) + 2 * (size_t)c.gmax * 64 * sizeof(float);
  return (b + 1023) & ~(size_t)1023;
}

static void stack_fill(StackCfg* c, int tpg) {
  c->gmax = (LCN_J + c->ns - 1) / c->ns;
  const int cols = tpg * c->gmax * 64;      // per virtual CTA (the head layer needs 128)
  const int per_vc = cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  c->tmem_cols = per_vc * c->nv;             // > 512: rejected by the caller
  c->stage_bytes = tpg * ST_A_BYTES + c->gmax * ST_B_BYTES;
  c->vc_bytes = stack_vc_bytes(*c);
  c->smem = c->nv * c->vc_bytes + 1024;
}

template <int EW, int NV>
static int stack_active_clusters(const StackCfg& c, int sm_count) {
  static std::once_flag once;
  static bool attr_ok = false;
  std::call_once(once, [] {
    attr_ok = cudaFuncSetAttribute(k_lcn_stack<EW, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 231424) == cudaSuccess &&
              cudaFuncSetAttribute(k_lcn_stack<EW, NV>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (!attr_ok) {
      if (stack_verbose()) fprintf(stderr, "lcn_stack: cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      (void)cudaGetLastError();
    }
  });
  if (!attr_ok) return 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)c.ns;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(NV * (64 + 32 * EW));
  cfg.dynamicSmemBytes = c.smem;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)(sm_count / c.ns * c.ns));
  int active = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&active, k_lcn_stack<EW, NV>, &cfg);
  if (stack_verbose())
    fprintf(stderr, "lcn_stack: occupancy query EW=%d NV=%d ns=%d smem=%zu -> %s, %d active clusters\n", EW, NV, c.ns, c.smem,
            cudaGetErrorString(e), active);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return active;
}

// Configuration for `tpg` row tiles per BatchNorm group; `clusters` = how many clusters the device keeps resident.
// Decided once per (tpg) and cached: the occupancy query needs the device, the workspace layout needs the answer.
static StackCfg stack_config(int tpg, int sm_count, int* clusters) {
  static std::mutex mu;
  static StackCfg cache[3];
  static int cache_cl[3] = {0, 0, 0};
  std::lock_guard<std::mutex> lock(mu);
  if (cache_cl[tpg]) {
    *clusters = cache_cl[tpg];
    return cache[tpg];
  }
  StackCfg c;
  memset(&c, 0, sizeof(c));
  c.mc = 1;
  int active = 0;
  {
    // two groups per SM: <= 256 TMEM columns each
    c.nv = 2;
    c.ns = tpg == 2 ? 9 : 5;
    c.stages = 2;
    stack_fill(&c, tpg);
    if (c.ns >= 1 && c.ns <= ST_MAX_NS && c.tmem_cols <= 512 && c.smem <= 231424) active = stack_active_clusters<4, 2>(c, sm_count);
  }
  if (active == 0) {
    c.nv = 1;
    c.ns = tpg == 2 ? 6 : 3;
    c.stages = 3;
    stack_fill(&c, tpg);
    active = stack_active_clusters<8, 1>(c, sm_count);
    if (active <= 0) active = sm_count / c.ns;
  }
  const int want = sm_count / c.ns;
  if (active > want) active = want;
  if (stack_verbose())
    fprintf(stderr, "lcn_stack: tpg=%d ns=%d gmax=%d stages=%d mc=%d nv=%d tmem=%d stage=%d smem=%zu clusters=%d\n", tpg, c.ns,
            c.gmax, c.stages, c.mc, c.nv, c.tmem_cols, c.stage_bytes, c.smem, active);
  cache[tpg] = c;
  cache_cl[tpg] = active;
  *clusters = active;
  return c;
}

bool lcn_stack_eligible(const lcn_model* m, int bn_group, int training) {
  return !training && m->d.path == LCN_PATH_BF16 && m->FC == 1 &&
         bn_group <= 2 * LCN_TILE && LCN_J * m->d.in_F <= 64 && m->n_lin >= 3;
}

size_t lcn_stack_scratch_bytes(const lcn_model* m, int bn_group) {
  int tpg = (bn_group + LCN_TILE - 1) / LCN_TILE, clusters = 0;
  const StackCfg c = stack_config(tpg, m->sm_count, &clusters);
  return (size_t)clusters * c.nv * 2 * tpg * LCN_J * ST_A_BYTES;
}

int lcn_stack_forward(const lcn_model* m, const WsLayout& lay, const float* params, char* ws, const float* x,
                      float* out, void* taps, cudaStream_t st) {
  StackParams p;
  memset(&p, 0, sizeof(p));
  const int tpg = lay.tiles_per_group;
  int max_clusters = 0;
  const StackCfg c = stack_config(tpg, m->sm_count, &max_clusters);
  const int ns = c.ns;
  p.x = x;
  p.out = out;
  p.params = params;
  p.wf16 = reinterpret_cast<const __nv_bfloat16*>(ws + lay.off_wf16);
  p.wp16f = reinterpret_cast<const __nv_bfloat16*>(ws + lay.off_wp16f);
  p.wl16f = reinterpret_cast<const __nv_bfloat16*>(ws + lay.off_wl16f);
  p.scratch = reinterpret_cast<__nv_bfloat16*>(ws + lay.off_stack);
  p.taps = reinterpret_cast<__nv_bfloat16*>(taps);
  p.taps_stride = (int64_t)lay.tiles * LCN_J * 8192;
  for (int l = 0; l < m->n_lin; ++l) {
    p.b_off[l] = m->L[l].b_off;
    p.gamma_off[l] = m->L[l].gamma_off;
    p.beta_off[l] = m->L[l].beta_off;
    p.res[l] = m->L[l].res_from >= 0 ? 1 : 0;
  }
  for (int a = 0; a < LCN_J; ++a) p.kmask[a] = m->sup.row[a];
  stack_partition(m, ns, c.gmax, p.oc_start);
  for (int r = 0; r < ns; ++r) {
    const int oc0 = p.oc_start[r], G = p.oc_start[r + 1] - oc0;
    uint32_t written = 0;
    for (int kc = 0; kc < LCN_J; ++kc) {
      uint32_t bits = 0;
      for (int q = 0; q < G; ++q)
        if ((p.kmask[kc] >> (oc0 + q)) & 1u) bits |= 1u << q;
      int n = 0, rnk = 0, q = 0;
      while (q < G) {
        if (!((bits >> q) & 1u)) { ++q; continue; }
        const uint32_t acc = (written >> q) & 1u;
        int len = 1;
        while (len < ST_MAX_RUN && q + len < G && ((bits >> (q + len)) & 1u) && (((written >> (q + len)) & 1u) == acc)) ++len;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((64 * len) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        p.prog[r][kc][n++] = make_uint4((uint32_t)q * 64u, (uint32_t)(rnk * ST_B_BYTES) >> 4, idesc, acc);
        written |= ((1u << len) - 1u) << q;
        rnk += len;
        q += len;
      }
      p.prog_cnt[r][kc] = (uint8_t)n;
      p.sch_cnt[r][kc] = (uint8_t)__builtin_popcount(bits);
      int slot = 0;
      if (bits) {
        for (int a = 0; a < kc; ++a) slot += __builtin_popcount(p.kmask[a]);
        slot += __builtin_popcount(p.kmask[kc] & ((1u << (oc0 + __builtin_ctz(bits))) - 1u));
      }
      p.sch_slot[r][kc] = (uint16_t)slot;
    }
  }
  p.n_lin = m->n_lin;
  p.in_F = m->d.in_F;
  p.TPG = tpg;
  p.stages = c.stages;
  p.mc = c.mc;
  p.nnz = m->nnz;
  p.tmem_cols = c.tmem_cols;
  p.stage_bytes = c.stage_bytes;
  p.vc_bytes = (int)c.vc_bytes;
  p.cols = c.gmax * 64;
#ifdef LCN_TC_PROFILE
  { const char* e = getenv("LCN_STACK_DBG"); p.dbg = e ? atoi(e) : 0; }    // profiling build only (timing experiments)
#else
  p.dbg = 0;
#endif
  p.n_rows = lay.n_rows;
  p.bn_group = lay.bn_group;
  p.n_groups = lay.n_groups;

  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)ns;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(c.nv == 2 ? 384 : 320);
  cfg.dynamicSmemBytes = c.smem;
  cfg.stream = st;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const int need = (lay.n_groups + c.nv - 1) / c.nv;
  const int clusters = need < max_clusters ? need : max_clusters;
  cfg.gridDim = dim3((unsigned)(clusters * ns));
  if (c.nv == 2) {
    LCN_CHECK_CUDA((cudaLaunchKernelEx(&cfg, k_lcn_stack<4, 2>, p)));
  } else {
    LCN_CHECK_CUDA((cudaLaunchKernelEx(&cfg, k_lcn_stack<8, 1>, p)));
  }
  return LCN_OK;
}
