// Data-parallel gradient exchange inside the library: a two-shot all-reduce written for this bucket, over NVLink peer
// memory.  The reference has no distributed code at all (network/models_att.py:155-158: one process, one
// tf.Session); this is the exchange step a data-parallel host needs between optimizer.compute_gradients and
// apply_gradients (:408-409).
//
// One process per GPU.  Every rank owns ONE device allocation [flags | gradient bucket] that all peers map through CUDA
// IPC (lcn_dp_export -> the host all-gathers the 64-byte handles -> lcn_dp_connect); lcn_dp_bucket() is the pointer the
// caller passes as d_grads_raw, so the backward pass writes its gradients straight into peer-visible memory.  At the
// end of lcn_model_backward:
//
//   k_dp_publish   epoch += 1; ready[me] = epoch in every peer's flag block                (st.release.sys over NVLink)
//   k_dp_reduce    (lcn_kernels.cu) waits until ready[p] >= epoch for every peer p; the units of the bucket -- nonzero
//                  joint-pair blocks of the weight gradients, small tensors -- are dealt round-robin to the ranks; for
//                  its units a rank loads the block from ALL buckets (16-byte loads, the peers' straight over NVLink),
//                  takes the mean and stores it into ALL buckets: reduce-scatter by peer LOADS, all-gather by peer
//                  STORES, in place, one pass, no staging or pack / unpack copy; the masked-out 40 % of the bucket never
//                  travels; the last CTA publishes done[me] = epoch to every peer
//   k_dp_wait      waits until done[p] >= epoch for every p: every unit of this rank's bucket holds the mean, and every
//                  peer has finished reading it -- the chain rule / Adam kernels that follow see averaged gradients
//
// Each gradient byte crosses the NVLink fabric once in each direction.  Measured against ncclAllReduce of the packed
// bucket (with its pack / unpack passes, two graph replays) and against NCCL all-reduces per layer overlapped with the
// backward pass (SLOWER than no overlap: the NCCL CTAs take SMs from single-wave GEMMs and two-blocks-per-SM elementwise
// kernels): DESIGN.md section 5.
//
// Why this is safe without a cluster-wide barrier: flags only ever grow (epoch numbers), every rank executes the same
// sequence of exchanges, unit sets of different ranks are disjoint (in-place is race free), and the two waits order the
// reuse of the bucket -- a rank starts the next backward pass (which overwrites its bucket) only after k_dp_wait(e), i.e.
// after every peer published done(e), which a peer does after its last read of that bucket; a peer touches this rank's
// bucket for epoch e+1 only after this rank published ready(e+1).  All three kernels take only device pointers and read
// the epoch from device memory, so the exchange is captured into the train-step CUDA graph like any other kernel of the
// step.  No kernel needs a peer's kernel to be co-scheduled in order to FINISH its own loads and stores; the waits are on
// flags that the peers' stream-ordered work sets unconditionally, so a late peer delays, never deadlocks.
#include <stdlib.h>

#include <algorithm>

#include "lcn_internal.cuh"

#define LCN_DP_MAX_WORLD 8
#define LCN_DP_FLAG_BYTES 4096

struct LcnDpFlags {                       // lives at the start of every rank's exchange allocation
  unsigned long long epoch;               // local: exchanges started by this rank
  unsigned long long pad0[15];
  unsigned long long ready[LCN_DP_MAX_WORLD];   // ready[p]: written by rank p -- its xbuf holds epoch's bucket
  unsigned long long pad1[8];
  unsigned long long done[LCN_DP_MAX_WORLD];    // done[p]: written by rank p -- slice p of our rbuf holds epoch's mean
  unsigned long long pad2[8];
  unsigned int ticket;                    // local: CTAs of k_dp_reduce that have finished
};

struct LcnDp {
  int rank = 0, world = 1;
  bool enabled = true, connected = false;
  int64_t count = 0;                      // floats in the bucket (= lcn_model_param_count)
  size_t bytes = 0;
  char* base = nullptr;                   // own allocation
  char* peer[LCN_DP_MAX_WORLD] = {};      // mapped allocations (peer[rank] == base)
};

struct DpFlagPtrs {                       // by-value kernel argument
  LcnDpFlags* flags[LCN_DP_MAX_WORLD];
  int rank, world;
};

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void k_dp_publish(DpFlagPtrs d) {
  lcn_pdl_prologue();
  __shared__ unsigned long long e_s;
  if (threadIdx.x == 0) {
    e_s = d.flags[d.rank]->epoch + 1;
    d.flags[d.rank]->epoch = e_s;
    __threadfence_system();                // the bucket (previous kernels of this stream) is visible system-wide before the flag
  }
  __syncthreads();
  if ((int)threadIdx.x < d.world) st_release_sys(&d.flags[threadIdx.x]->ready[d.rank], e_s);
}

__global__ void k_dp_wait(DpFlagPtrs d) {
  lcn_pdl_prologue();
  LcnDpFlags* mine = d.flags[d.rank];
  if ((int)threadIdx.x < d.world) {
    const unsigned long long e = mine->epoch;
    unsigned long long spins = 0;
    while (ld_acquire_sys(&mine->done[threadIdx.x]) < e) {
      if (++spins > (1ull << 28)) {        // seconds: a peer died or the ranks disagree on the call sequence
        printf("lcn_dp: wait for a peer's done flag timed out\n");
        __trap();
      }
    }
  }
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------------------------
static size_t dp_bytes(int64_t count) { return LCN_DP_FLAG_BYTES + (((size_t)count * 4 + 255) & ~(size_t)255); }

extern "C" int lcn_dp_export(lcn_model* m, void* h_handle64) {
  LCN_REQUIRE(m != nullptr && h_handle64 != nullptr, "null argument");
  LCN_REQUIRE(m->dp == nullptr, "the model already has an exchange buffer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  static_assert(sizeof(LcnDpFlags) <= LCN_DP_FLAG_BYTES, "flag block");
  LcnDp* dp = new LcnDp();
  dp->count = m->n_params;
  dp->bytes = dp_bytes(dp->count);
  // the one device allocation the library makes: peers have to map it, so it cannot come from the caller's allocator
  cudaError_t e = cudaMalloc(&dp->base, dp->bytes);
  if (e == cudaSuccess) e = cudaMemset(dp->base, 0, dp->bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, dp->base);
  if (e != cudaSuccess) {
    lcn_set_error("lcn_dp_export: %s", cudaGetErrorString(e));
    if (dp->base) cudaFree(dp->base);
    delete dp;
    (void)cudaGetLastError();
    return LCN_ECUDA;
  }
  memcpy(h_handle64, &h, sizeof(h));
  m->dp = dp;
  return LCN_OK;
}

extern "C" float* lcn_dp_bucket(const lcn_model* m) {
  return (m && m->dp) ? reinterpret_cast<float*>(m->dp->base + LCN_DP_FLAG_BYTES) : nullptr;
}

extern "C" int lcn_dp_connect(lcn_model* m, const void* h_handles, int rank, int world) {
  LCN_REQUIRE(m != nullptr && h_handles != nullptr, "null argument");
  LCN_REQUIRE(m->dp != nullptr && !m->dp->connected, "call lcn_dp_export first (once)");
  LCN_REQUIRE(world >= 1 && world <= LCN_DP_MAX_WORLD && rank >= 0 && rank < world, "rank %d / world %d (max %d)", rank, world,
              LCN_DP_MAX_WORLD);
  LcnDp* dp = m->dp;
  dp->rank = rank;
  dp->world = world;
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      dp->peer[p] = dp->base;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(h_handles) + (size_t)p * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      lcn_set_error("lcn_dp_connect: cudaIpcOpenMemHandle(rank %d) -> %s (are the GPUs NVLink / P2P peers?)", p, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return LCN_ECUDA;
    }
    dp->peer[p] = static_cast<char*>(ptr);
  }
  dp->connected = true;
  return LCN_OK;
}

extern "C" int lcn_dp_world(const lcn_model* m) { return (m && m->dp && m->dp->connected) ? m->dp->world : 1; }

extern "C" int lcn_dp_enable(lcn_model* m, int on) {
  LCN_REQUIRE(m != nullptr, "null model");
  LCN_REQUIRE((m->dp != nullptr && m->dp->connected) || !on, "lcn_dp_enable: the model is not connected (lcn_dp_export / lcn_dp_connect)");
  if (m->dp) m->dp->enabled = on != 0;
  return LCN_OK;
}

void lcn_dp_destroy(lcn_model* m) {
  if (m == nullptr || m->dp == nullptr) return;
  LcnDp* dp = m->dp;
  cudaDeviceSynchronize();
  for (int p = 0; p < dp->world; ++p)
    if (p != dp->rank && dp->peer[p] != nullptr) cudaIpcCloseMemHandle(dp->peer[p]);
  if (dp->base) cudaFree(dp->base);
  (void)cudaGetLastError();
  delete dp;
  m->dp = nullptr;
}

bool lcn_dp_active(const lcn_model* m) { return m->dp != nullptr && m->dp->connected && m->dp->enabled && m->dp->world > 1; }

// Average the raw-gradient bucket over the ranks, in place, on `st` (called at the end of lcn_model_backward).
int lcn_dp_exchange(const lcn_model* m, float* graw, cudaStream_t st) {
  LcnDp* dp = m->dp;
  LCN_REQUIRE(graw == reinterpret_cast<float*>(dp->base + LCN_DP_FLAG_BYTES),
              "data-parallel model: d_grads_raw must be the peer-mapped bucket returned by lcn_dp_bucket()");
  DpFlagPtrs f;
  memset(&f, 0, sizeof(f));
  f.rank = dp->rank;
  f.world = dp->world;
  float* buckets[LCN_DP_MAX_WORLD];
  unsigned long long* done_at[LCN_DP_MAX_WORLD];
  for (int p = 0; p < dp->world; ++p) {
    f.flags[p] = reinterpret_cast<LcnDpFlags*>(dp->peer[p]);
    buckets[p] = reinterpret_cast<float*>(dp->peer[p] + LCN_DP_FLAG_BYTES);
    done_at[p] = &f.flags[p]->done[dp->rank];
  }
  LcnDpFlags* mine = f.flags[dp->rank];
  lcn_launch(k_dp_publish, dim3(1), dim3(32), 0, st, f);
  LCN_CHECK_LAUNCH();
  int rc = lcn_launch_dp_reduce(m, buckets, mine->ready, done_at, &mine->epoch, &mine->ticket, dp->rank, dp->world, st);
  if (rc) return rc;
  lcn_launch(k_dp_wait, dim3(1), dim3(32), 0, st, f);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

LCN_KTRACE_EXPORT(dp)
