// Data-parallel gradient exchange inside the library ("NCCL allreduce of the gradient ... overlapped with
// wgrad of earlier layers").  The reference has no distributed code at all (network/models_att.py:155-158: one process,
// one tf.Session); this is the exchange step a data-parallel host needs between optimizer.compute_gradients and
// apply_gradients (:408-409).
//
// One process per GPU.  lcn_dp_init() gives the model its own NCCL communicator, a communication stream and events.
// From then on lcn_model_backward() averages the gradient bucket over the ranks ITSELF, layer by layer, while the
// backward pass is still running:
//   * the weight gradient of mid layer l is complete as soon as its weight-gradient GEMM has run (side stream): an
//     in-place ncclAllReduce(avg) of that layer's [17F x 17F] region is enqueued on the communication stream right
//     behind it and overlaps the BatchNorm backward / dgrad / wgrad kernels of the layers below;
//   * what only completes at the end of the pass (edge-layer weights, every bias, BatchNorm gamma / beta) goes out as
//     ONE grouped launch (ncclGroupStart / End) after the last kernel;
//   * the caller's stream joins the communication stream before lcn_model_backward returns control of the bucket, so
//     lcn_model_adam_step sees averaged gradients.  Everything is event-ordered and capturable: a data-parallel train
//     step is ONE CUDA graph (forward, backward + exchange, Adam), replayed every step.
// The tensor-core GEMMs run single waves of <= 128 CTAs on 148 SMs, so the communicator is created with at
// most 16 CTAs: the collectives fit on the SMs the GEMMs leave free instead of competing for theirs.  No kernel of the
// backward pass needs all of its blocks resident at once (the grid-barrier BatchNorm kernel of round 1 is gone), so
// a resident NCCL kernel waiting for its peers cannot starve the compute kernels it shares the GPU with.
//
// libnccl is resolved at run time (dlopen "libnccl.so.2": in a torch process this is the NCCL 2.28 torch already
// loaded), so liblcn_b200.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include "lcn_internal.cuh"

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static const NcclApi api = [] {
    NcclApi a;
    a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (a.h == nullptr) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (a.h == nullptr) return a;
#define LCN_SYM(field, name) *reinterpret_cast<void**>(&a.field) = dlsym(a.h, name)
    LCN_SYM(GetUniqueId, "ncclGetUniqueId");
    LCN_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
    LCN_SYM(CommInitRank, "ncclCommInitRank");
    LCN_SYM(CommDestroy, "ncclCommDestroy");
    LCN_SYM(AllReduce, "ncclAllReduce");
    LCN_SYM(GroupStart, "ncclGroupStart");
    LCN_SYM(GroupEnd, "ncclGroupEnd");
    LCN_SYM(GetErrorString, "ncclGetErrorString");
#undef LCN_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

#define LCN_CHECK_NCCL(expr)                                                                             \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess) {                                                                             \
      lcn_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, nccl().GetErrorString(_r));             \
      return LCN_ECUDA;                                                                                  \
    }                                                                                                    \
  } while (0)

}  // namespace

struct LcnDp {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  bool enabled = true;
  cudaStream_t st = nullptr;                 // communication stream
  cudaEvent_t ev_ready = nullptr;            // producer -> communication stream
  cudaEvent_t ev_done = nullptr;             // communication stream -> consumer
};

extern "C" int lcn_dp_unique_id(void* h_id128) {
  LCN_REQUIRE(h_id128 != nullptr, "null argument");
  LCN_REQUIRE(nccl().ok, "libnccl.so.2 could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  LCN_CHECK_NCCL(nccl().GetUniqueId(&id));
  memcpy(h_id128, &id, sizeof(id));
  return LCN_OK;
}

extern "C" int lcn_dp_init(lcn_model* m, const void* h_id128, int rank, int world) {
  LCN_REQUIRE(m != nullptr && h_id128 != nullptr, "null argument");
  LCN_REQUIRE(world >= 1 && rank >= 0 && rank < world, "rank %d / world %d", rank, world);
  LCN_REQUIRE(m->dp == nullptr, "the model already has a communicator");
  LCN_REQUIRE(nccl().ok, "libnccl.so.2 could not be loaded");
  LcnDp* dp = new LcnDp();
  dp->rank = rank;
  dp->world = world;
  ncclUniqueId id;
  memcpy(&id, h_id128, sizeof(id));
  ncclResult_t r;
  if (nccl().CommInitRankConfig != nullptr) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.maxCTAs = 16;                        // stay on the SMs the single-wave GEMMs leave free
    r = nccl().CommInitRankConfig(&dp->comm, world, id, rank, &cfg);
  } else {
    r = nccl().CommInitRank(&dp->comm, world, id, rank);
  }
  if (r != ncclSuccess) {
    lcn_set_error("ncclCommInitRank failed: %s", nccl().GetErrorString(r));
    delete dp;
    return LCN_ECUDA;
  }
  bool ok = cudaStreamCreateWithFlags(&dp->st, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&dp->ev_ready, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&dp->ev_done, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    lcn_set_error("lcn_dp_init: stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    nccl().CommDestroy(dp->comm);
    delete dp;
    return LCN_ECUDA;
  }
  m->dp = dp;
  return LCN_OK;
}

extern "C" int lcn_dp_world(const lcn_model* m) { return (m && m->dp) ? m->dp->world : 1; }

extern "C" int lcn_dp_enable(lcn_model* m, int on) {
  LCN_REQUIRE(m != nullptr, "null model");
  LCN_REQUIRE(m->dp != nullptr || !on, "lcn_dp_enable: the model has no communicator (lcn_dp_init)");
  if (m->dp) m->dp->enabled = on != 0;
  return LCN_OK;
}

void lcn_dp_destroy(lcn_model* m) {
  if (m == nullptr || m->dp == nullptr) return;
  LcnDp* dp = m->dp;
  if (dp->st) cudaStreamSynchronize(dp->st);
  if (dp->comm) nccl().CommDestroy(dp->comm);
  if (dp->ev_ready) cudaEventDestroy(dp->ev_ready);
  if (dp->ev_done) cudaEventDestroy(dp->ev_done);
  if (dp->st) cudaStreamDestroy(dp->st);
  delete dp;
  m->dp = nullptr;
}

bool lcn_dp_active(const lcn_model* m) { return m->dp != nullptr && m->dp->enabled && m->dp->world > 1; }

// Average buf[0, count) over the ranks, in place, on the communication stream, once everything enqueued on `producer`
// so far has run.
int lcn_dp_allreduce_after(const lcn_model* m, cudaStream_t producer, float* buf, size_t count) {
  LcnDp* dp = m->dp;
  LCN_CHECK_CUDA(cudaEventRecord(dp->ev_ready, producer));
  LCN_CHECK_CUDA(cudaStreamWaitEvent(dp->st, dp->ev_ready, 0));
  LCN_CHECK_NCCL(nccl().AllReduce(buf, buf, count, ncclFloat32, ncclAvg, dp->comm, dp->st));
  return LCN_OK;
}

// The same for several regions in ONE launch (ncclGroupStart / End).
int lcn_dp_allreduce_group_after(const lcn_model* m, cudaStream_t producer, float* base, const int64_t* offs,
                                 const int64_t* counts, int n) {
  LcnDp* dp = m->dp;
  LCN_CHECK_CUDA(cudaEventRecord(dp->ev_ready, producer));
  LCN_CHECK_CUDA(cudaStreamWaitEvent(dp->st, dp->ev_ready, 0));
  LCN_CHECK_NCCL(nccl().GroupStart());
  for (int i = 0; i < n; ++i) {
    ncclResult_t r = nccl().AllReduce(base + offs[i], base + offs[i], (size_t)counts[i], ncclFloat32, ncclAvg, dp->comm, dp->st);
    if (r != ncclSuccess) {
      nccl().GroupEnd();
      lcn_set_error("ncclAllReduce (grouped) failed: %s", nccl().GetErrorString(r));
      return LCN_ECUDA;
    }
  }
  LCN_CHECK_NCCL(nccl().GroupEnd());
  return LCN_OK;
}

// `consumer` continues when every collective enqueued so far has completed.
int lcn_dp_join(const lcn_model* m, cudaStream_t consumer) {
  LcnDp* dp = m->dp;
  LCN_CHECK_CUDA(cudaEventRecord(dp->ev_done, dp->st));
  LCN_CHECK_CUDA(cudaStreamWaitEvent(consumer, dp->ev_done, 0));
  return LCN_OK;
}
