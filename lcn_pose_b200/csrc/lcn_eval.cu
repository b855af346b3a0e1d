// (c) Batched evaluation: evaluate.py:53-61 per pose ->
//   tools.image_to_camera_frame (tools/tools.py:183-194)
//   [protocol 2] tools.align_to_gt -> procrustes(gt, pred) (tools/tools.py:96-181,197-202),
//                scaling=True, reflection='best' (no determinant fix: reflections are accepted)
//   err[j] = || pred_j - gt_j ||_2  (evaluate.py:61), per-joint sums, PCK@50mm (evaluate.py:78-106)
//
// HBM-bound: 444 B read (+68 B written when per-joint errors are requested) per pose.  One warp owns
// 32 consecutive poses per iteration: the 2 x 6528 B of pred/gt arrive in shared memory as two 1-D bulk TMA copies
// (cp.async.bulk, completion on a per-warp mbarrier), DOUBLE BUFFERED -- the copies of chunk i+1 (and the
// box / camera / root-depth words of its poses) are in flight while the lanes work on chunk i -- then each lane
// works on its own pose out of shared memory (row stride 51 words, odd -> conflict free) with the 3x3 SVD
// (one-sided Jacobi) entirely in registers.  2 blocks x 4 warps per SM keep ~100 KB of loads in flight per SM.
#include <algorithm>

#include "lcn_internal.cuh"
#include "lcn_tc_ptx.cuh"

#define EV_WARPS 8
#define EV_CHUNK_BYTES (32 * 51 * 4)   // 6528 B of pred (or gt) for 32 poses: a multiple of 16
#define EV_POSE 51

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// one Jacobi rotation orthogonalising columns p,q of G (3x3, column vectors) and accumulating V
// returns true when the pair was not yet orthogonal (a rotation was applied)
__device__ __forceinline__ bool jacobi_rot(float* gp, float* gq, float* vp, float* vq) {
  float alpha = gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2];
  float beta = gq[0] * gq[0] + gq[1] * gq[1] + gq[2] * gq[2];
  float gamma = gp[0] * gq[0] + gp[1] * gq[1] + gp[2] * gq[2];
  if (gamma * gamma <= 1e-14f * alpha * beta) return false;
  float zeta = (beta - alpha) / (2.f * gamma);
  float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
  float c = rsqrtf(1.f + t * t), s = c * t;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = gp[k], b = gq[k];
    gp[k] = c * a - s * b;
    gq[k] = s * a + c * b;
    a = vp[k]; b = vq[k];
    vp[k] = c * a - s * b;
    vq[k] = s * a + c * b;
  }
  return true;
}

// EV_BUFS staging buffers per warp.  2: double buffered, one block of 8 warps per SM (Protocol 1: pure streaming, the
// next chunk's copies overlap this chunk's arithmetic inside the warp).  1: 13 KB per warp, two blocks = 16 warps per
// SM (Protocol 2: the per-pose SVD chain is latency bound, more resident warps hide it).
template <int EV_BUFS>
__global__ void __launch_bounds__(EV_WARPS * 32, EV_BUFS == 1 ? 2 : 1) k_eval(const float* __restrict__ pred, const float* __restrict__ gt,
                                                       const float* __restrict__ box, const float* __restrict__ cam,
                                                       const float* __restrict__ root_depth,
                                                       const int32_t* __restrict__ action, int n_actions, int64_t n,
                                                       int flags, float* __restrict__ err_out,
                                                       float* __restrict__ pose_out, double* __restrict__ sums) {
  const int protocol2 = flags & 1, camera_frame = flags & 2;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[EV_WARPS][2];
  // per warp: two buffers of {pred[32*51], gt[32*51]}; then per block: action sums (double)
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wbuf = smem + warp * (EV_BUFS * 2 * 32 * EV_POSE);
  double* asum = reinterpret_cast<double*>(smem + EV_WARPS * EV_BUFS * 2 * 32 * EV_POSE);   // [n_actions][19]
  for (int e = threadIdx.x; e < n_actions * 19; e += blockDim.x) asum[e] = 0.0;
  const uint32_t bar0 = smem_u32(&bars[warp][0]);
  if (lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double jsum[LCN_J];
#pragma unroll
  for (int j = 0; j < LCN_J; ++j) jsum[j] = 0.0;
  unsigned long long npose = 0, npck = 0;

  int64_t n_chunks = (n + 31) / 32;
  int64_t wglobal = (int64_t)blockIdx.x * EV_WARPS + warp, wstride = (int64_t)gridDim.x * EV_WARPS;
  // full chunks travel by bulk copy into buffer `b`; the (single, last) partial chunk is loaded by the lanes
  auto issue = [&](int64_t ch, int b) {
    if (ch >= n_chunks || (ch + 1) * 32 > n) return;
    if (lane == 0) {
      const uint32_t bar = bar0 + 8 * b;
      float* dp = wbuf + b * (2 * 32 * EV_POSE);
      mbar_expect_tx(bar, 2 * EV_CHUNK_BYTES);
      bulk_g2s(smem_u32(dp), pred + ch * 32 * EV_POSE, EV_CHUNK_BYTES, bar);
      bulk_g2s(smem_u32(dp + 32 * EV_POSE), gt + ch * 32 * EV_POSE, EV_CHUNK_BYTES, bar);
    }
  };
  // per-pose scalars of a chunk (box, camera, root depth, action): fetched one chunk ahead into registers
  float4 bx_n = make_float4(0.f, 0.f, 1999.f, 0.f), cm_n = make_float4(1.f, 1.f, 0.f, 0.f);
  float rd_n = 0.f;
  int act_n = -1;
  auto fetch_scalars = [&](int64_t ch) {
    bx_n = make_float4(0.f, 0.f, 1999.f, 0.f);
    cm_n = make_float4(1.f, 1.f, 0.f, 0.f);
    rd_n = 0.f;
    act_n = -1;
    int64_t ip = ch * 32 + lane;
    if (ch < n_chunks && ip < n) {
      if (!camera_frame) {
        bx_n = ld_stream4(box + ip * 4);
        cm_n = ld_stream4(cam + ip * 4);                     // fx, fy, cx, cy
        rd_n = __ldg(root_depth + ip);
      }
      if (action != nullptr) act_n = __ldg(action + ip);
    }
  };
  issue(wglobal, 0);
  fetch_scalars(wglobal);
  int buf = 0;
  uint32_t phase[2] = {0u, 0u};
  for (int64_t ch = wglobal; ch < n_chunks; ch += wstride, buf ^= (EV_BUFS - 1)) {
    int64_t p0 = ch * 32;
    int cnt = (int)min((int64_t)32, n - p0);
    float* sp = wbuf + buf * (2 * 32 * EV_POSE);
    float* sg = sp + 32 * EV_POSE;
    const float4 bx = bx_n, cm = cm_n;
    const float rd = rd_n;
    const int act = act_n;
    // next chunk of this warp.  Double buffered: into the other buffer now (last read two iterations ago, fenced
    // below); single buffered: after this chunk's reads, at the bottom of the loop.  The per-pose scalars of the
    // next chunk are always fetched one chunk ahead.
    if (EV_BUFS == 2) issue(ch + wstride, buf ^ 1);
    fetch_scalars(ch + wstride);
    if (cnt == 32) {
      mbar_wait(bar0 + 8 * buf, phase[buf]);
      phase[buf] ^= 1u;
    } else {
      const float* gp = pred + p0 * EV_POSE;
      const float* gg = gt + p0 * EV_POSE;
      for (int v = lane; v < cnt * EV_POSE; v += 32) {
        sp[v] = gp[v];
        sg[v] = gg[v];
      }
      __syncwarp();
    }
    bool active = lane < cnt;
    float e[LCN_J];
#pragma unroll
    for (int j = 0; j < LCN_J; ++j) e[j] = 0.f;
    if (active) {
      const float inv_ratio = 2000.0f / (bx.z - bx.x + 1.0f);
      const float ifx = 1.0f / cm.x, ify = 1.0f / cm.y;
      float* mp = sp + lane * EV_POSE;
      const float* mg = sg + lane * EV_POSE;
      // prediction joint j in the camera frame (tools.py:187-193); the pose is re-read from shared memory in every
      // pass instead of living in 102 registers: the kernel then fits 128 registers -> 16 warps per SM
      auto cam_joint = [&](int j, float* q) {
        if (camera_frame) {
          q[0] = mp[j * 3 + 0]; q[1] = mp[j * 3 + 1]; q[2] = mp[j * 3 + 2];
        } else {
          float z = mp[j * 3 + 2] * inv_ratio + rd;          // tools.py:187
          q[0] = (mp[j * 3 + 0] - cm.z) * ifx * z;           // tools.py:190,192
          q[1] = (mp[j * 3 + 1] - cm.w) * ify * z;           // tools.py:191,193
          q[2] = z;
        }
      };
      if (protocol2) {
        // pass 1: moments of both poses relative to their root joints (values ~1e2 mm instead of ~1e3-1e4 mm: the
        // centred second moments keep fp32 accuracy), then the shift to the centroids in closed form
        float p0[3], g0r[3];
        cam_joint(0, p0);
        g0r[0] = mg[0]; g0r[1] = mg[1]; g0r[2] = mg[2];
        float sP[3] = {0, 0, 0}, sG[3] = {0, 0, 0}, ssA = 0.f, ssB = 0.f;
        float M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // M[a][b] = sum_k A0[k][a] * B0[k][b]
#pragma unroll
        for (int j = 1; j < LCN_J; ++j) {
          float q[3], g[3];
          cam_joint(j, q);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            q[c] -= p0[c];
            g[c] = mg[j * 3 + c] - g0r[c];
            sP[c] += q[c];
            sG[c] += g[c];
            ssA = fmaf(g[c], g[c], ssA);
            ssB = fmaf(q[c], q[c], ssB);
          }
#pragma unroll
          for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2) M[a2][b2] = fmaf(g[a2], q[b2], M[a2][b2]);
        }
        float pb[3], gb[3];                                   // centroids relative to the roots (tools.py:127-130)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          pb[c] = sP[c] * (1.0f / LCN_J);
          gb[c] = sG[c] * (1.0f / LCN_J);
        }
        ssA -= (float)LCN_J * (gb[0] * gb[0] + gb[1] * gb[1] + gb[2] * gb[2]);
        ssB -= (float)LCN_J * (pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]);
#pragma unroll
        for (int a2 = 0; a2 < 3; ++a2)
#pragma unroll
          for (int b2 = 0; b2 < 3; ++b2) M[a2][b2] -= (float)LCN_J * gb[a2] * pb[b2];
        float nA = sqrtf(fmaxf(ssA, 0.f)), nB = sqrtf(fmaxf(ssB, 0.f));
        float inv_ab = 1.0f / (nA * nB);
        // columns of M (scaled to the unit-norm problem of tools.py:133-144)
        float g0[3], g1[3], g2[3], v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
#pragma unroll
        for (int a2 = 0; a2 < 3; ++a2) {
          g0[a2] = M[a2][0] * inv_ab;
          g1[a2] = M[a2][1] * inv_ab;
          g2[a2] = M[a2][2] * inv_ab;
        }
        // one-sided Jacobi: M V = U S  (np.linalg.svd, tools.py:145).  At most 6 sweeps; the warp stops as soon as a
        // whole sweep rotated nothing in any of its lanes (cyclic Jacobi converges quadratically)
        const unsigned lanes = __activemask();
#pragma unroll 1
        for (int sweep = 0; sweep < 6; ++sweep) {
          bool r = jacobi_rot(g0, g1, v0, v1);
          r |= jacobi_rot(g0, g2, v0, v2);
          r |= jacobi_rot(g1, g2, v1, v2);
          if (!__any_sync(lanes, r)) break;
        }
        float s0 = sqrtf(g0[0] * g0[0] + g0[1] * g0[1] + g0[2] * g0[2]);
        float s1 = sqrtf(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]);
        float s2 = sqrtf(g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2]);
        float tr = s0 + s1 + s2;                            // S_trace, tools.py:159
        float i0 = s0 > 1e-20f ? 1.f / s0 : 0.f, i1 = s1 > 1e-20f ? 1.f / s1 : 0.f, i2 = s2 > 1e-20f ? 1.f / s2 : 0.f;
        // R = V U^T (tools.py:146-147): R[b][c] = sum_s V[b][s] U[c][s], U[:,s] = g_s / s_s
        float R[3][3];
#pragma unroll
        for (int b2 = 0; b2 < 3; ++b2)
#pragma unroll
          for (int c = 0; c < 3; ++c) R[b2][c] = v0[b2] * g0[c] * i0 + v1[b2] * g1[c] * i1 + v2[b2] * g2[c] * i2;
        // pass 2: Z - gt = nA*tr*(B0/nB) R - A0   (tools.py:168 minus the common centroid)
        const float sc = nA * tr / nB;
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) {
          float q[3], g[3], zc[3];
          cam_joint(j, q);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            q[c] -= p0[c] + pb[c];                          // B0 (prediction), tools.py:130
            g[c] = mg[j * 3 + c] - g0r[c] - gb[c];          // A0 (ground truth), tools.py:129
          }
          float d2 = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            zc[c] = sc * (q[0] * R[0][c] + q[1] * R[1][c] + q[2] * R[2][c]) - g[c];
            d2 = fmaf(zc[c], zc[c], d2);
          }
          e[j] = sqrtf(d2);
          if (pose_out != nullptr) {   // aligned pose Z (tools.py:168); a lane only rewrites its own 51 words
#pragma unroll
            for (int c = 0; c < 3; ++c) mp[j * 3 + c] = zc[c] + mg[j * 3 + c];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) {
          float q[3];
          cam_joint(j, q);
          float dx = q[0] - mg[j * 3 + 0], dy = q[1] - mg[j * 3 + 1], dz = q[2] - mg[j * 3 + 2];
          e[j] = sqrtf(dx * dx + dy * dy + dz * dz);
          if (pose_out != nullptr) {
            mp[j * 3 + 0] = q[0]; mp[j * 3 + 1] = q[1]; mp[j * 3 + 2] = q[2];
          }
        }
      }
      int pck = 0;
#pragma unroll
      for (int j = 0; j < LCN_J; ++j) {
        jsum[j] += (double)e[j];
        pck += e[j] < 50.0f ? 1 : 0;
      }
      npose += 1;
      npck += pck;
      if (act >= 0 && act < n_actions) {
        double* row = asum + act * 19;
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) atomicAdd(&row[j], (double)e[j]);
        atomicAdd(&row[17], 1.0);
        atomicAdd(&row[18], (double)pck);
      }
    }
    if (pose_out != nullptr) {
      __syncwarp();
      float* dstp = pose_out + p0 * EV_POSE;
      for (int v = lane; v < cnt * EV_POSE; v += 32) dstp[v] = sp[v];
    }
    if (err_out != nullptr) {
      __syncwarp();
      if (active) {
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) sp[lane * LCN_J + j] = e[j];
      }
      __syncwarp();
      float* dst = err_out + p0 * LCN_J;
      for (int v = lane; v < cnt * LCN_J; v += 32) dst[v] = sp[v];
    }
    // this buffer is overwritten by a bulk copy (async proxy) issued at the top of the NEXT iteration: order the
    // lanes' generic-proxy accesses before it
    fence_proxy_async();
    __syncwarp();
    if (EV_BUFS == 1) issue(ch + wstride, 0);
  }
  // block reduction of the "all poses" row, then one atomic per entry per block
  __syncthreads();
  double* red = reinterpret_cast<double*>(smem);   // reuse staging area: [EV_WARPS][19]
  double pn = (double)npose, pk = (double)npck;
#pragma unroll
  for (int j = 0; j < LCN_J; ++j)
    for (int o = 16; o > 0; o >>= 1) jsum[j] += __shfl_xor_sync(0xffffffffu, jsum[j], o);
  for (int o = 16; o > 0; o >>= 1) {
    pn += __shfl_xor_sync(0xffffffffu, pn, o);
    pk += __shfl_xor_sync(0xffffffffu, pk, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < LCN_J; ++j) red[warp * 19 + j] = jsum[j];
    red[warp * 19 + 17] = pn;
    red[warp * 19 + 18] = pk;
  }
  __syncthreads();
  double* all_row = sums + (size_t)n_actions * 19;
  if (threadIdx.x < 19) {
    double s = 0.0;
    for (int w = 0; w < EV_WARPS; ++w) s += red[w * 19 + threadIdx.x];
    if (s != 0.0) atomicAdd(&all_row[threadIdx.x], s);
  }
  for (int e2 = threadIdx.x; e2 < n_actions * 19; e2 += blockDim.x)
    if (asum[e2] != 0.0) atomicAdd(&sums[e2], asum[e2]);
}

extern "C" int lcn_eval_mpjpe(const float* d_pred, const float* d_gt, const float* d_box, const float* d_cam,
                              const float* d_root_depth, const int32_t* d_action, int32_t n_actions, int64_t n,
                              int flags, float* d_err, float* d_pose_out, double* d_sums, void* stream) {
  LCN_REQUIRE(d_pred && d_gt && d_sums, "null argument");
  LCN_REQUIRE((flags & 2) || (d_box && d_cam && d_root_depth), "box / cam / root_depth are required unless LCN_EVAL_CAMERA_FRAME");
  LCN_REQUIRE(n > 0, "n must be positive");
  LCN_REQUIRE(n_actions >= 0 && n_actions <= 64, "n_actions=%d outside 0..64", n_actions);
  LCN_REQUIRE((((uintptr_t)d_pred | (uintptr_t)d_gt | (uintptr_t)d_box | (uintptr_t)d_cam) & 15) == 0,
              "pred/gt/box/cam must be 16-byte aligned");
  if (d_action == nullptr) n_actions = 0;
  const int bufs = (flags & 1) ? 1 : 2;
  size_t smem = (size_t)EV_WARPS * bufs * 2 * 32 * EV_POSE * sizeof(float) + (size_t)n_actions * 19 * sizeof(double);
  static std::once_flag once;            // thread-safe one-time attribute setup (include/lcn_b200.h: re-entrancy)
  static cudaError_t once_rc = cudaSuccess;
  std::call_once(once, [] {
    once_rc = cudaFuncSetAttribute(k_eval<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (once_rc == cudaSuccess) once_rc = cudaFuncSetAttribute(k_eval<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  });
  LCN_CHECK_CUDA(once_rc);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t chunks = (n + 31) / 32;
  int64_t blocks = (chunks + EV_WARPS - 1) / EV_WARPS;
  int grid = (int)std::min<int64_t>(blocks, (int64_t)sms * (bufs == 1 ? 2 : 1));
  if (bufs == 1)
    k_eval<1><<<grid, EV_WARPS * 32, smem, (cudaStream_t)stream>>>(d_pred, d_gt, d_box, d_cam, d_root_depth, d_action,
                                                                 n_actions, n, flags, d_err, d_pose_out, d_sums);
  else
    k_eval<2><<<grid, EV_WARPS * 32, smem, (cudaStream_t)stream>>>(d_pred, d_gt, d_box, d_cam, d_root_depth, d_action,
                                                                 n_actions, n, flags, d_err, d_pose_out, d_sums);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

// tools.procrustes(A, B, scaling, reflection) in full (tools/tools.py:96-181): one thread per pose pair, every option
// of the reference's signature, and the (d, Z, tform) triple it returns.  The evaluator above inlines the
// scaling=True / reflection='best' case that evaluate.py uses; this entry serves direct callers of tools.procrustes.
// flags: bit 0 = scaling, bits 1-2 = reflection (0 'best', 1 force none, 2 force one).
__global__ void __launch_bounds__(128) k_procrustes(const float* __restrict__ A, const float* __restrict__ B, int64_t n,
                                                    int flags, float* __restrict__ Z, float* __restrict__ tform) {
  const int64_t ip = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= n) return;
  const float* a = A + ip * EV_POSE;
  const float* b = B + ip * EV_POSE;
  const int scaling = flags & 1, refl = (flags >> 1) & 3;
  float ab[3] = {0, 0, 0}, bb[3] = {0, 0, 0};
  for (int j = 0; j < LCN_J; ++j)
    for (int c = 0; c < 3; ++c) {
      ab[c] += a[j * 3 + c] - a[c];                          // relative to joint 0: keeps fp32 accuracy at ~1e3-1e4 mm
      bb[c] += b[j * 3 + c] - b[c];
    }
  for (int c = 0; c < 3; ++c) {
    ab[c] = ab[c] * (1.0f / LCN_J) + a[c];                   // A_bar, B_bar (tools.py:127-128)
    bb[c] = bb[c] * (1.0f / LCN_J) + b[c];
  }
  float ssX = 0.f, ssY = 0.f, M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < LCN_J; ++j) {
    float x[3], y[3];
    for (int c = 0; c < 3; ++c) {
      x[c] = a[j * 3 + c] - ab[c];
      y[c] = b[j * 3 + c] - bb[c];
      ssX = fmaf(x[c], x[c], ssX);
      ssY = fmaf(y[c], y[c], ssY);
    }
    for (int p = 0; p < 3; ++p)
      for (int q = 0; q < 3; ++q) M[p][q] = fmaf(x[p], y[q], M[p][q]);   // A0^T B0 (tools.py:144)
  }
  const float nA = sqrtf(ssX), nB = sqrtf(ssY), inv_ab = 1.0f / (nA * nB);
  float g[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};          // g[s] = column s of M, v[s] = column s of V
  for (int s2 = 0; s2 < 3; ++s2)
    for (int p = 0; p < 3; ++p) g[s2][p] = M[p][s2] * inv_ab;
  for (int sweep = 0; sweep < 8; ++sweep) {
    bool r = jacobi_rot(g[0], g[1], v[0], v[1]);
    r |= jacobi_rot(g[0], g[2], v[0], v[2]);
    r |= jacobi_rot(g[1], g[2], v[1], v[2]);
    if (!r) break;
  }
  float sv[3], inv[3];
  for (int s2 = 0; s2 < 3; ++s2) {
    sv[s2] = sqrtf(g[s2][0] * g[s2][0] + g[s2][1] * g[s2][1] + g[s2][2] * g[s2][2]);
    inv[s2] = sv[s2] > 1e-20f ? 1.f / sv[s2] : 0.f;
  }
  float R[3][3];
  auto make_R = [&]() {
    for (int p = 0; p < 3; ++p)
      for (int c = 0; c < 3; ++c)
        R[p][c] = v[0][p] * g[0][c] * inv[0] + v[1][p] * g[1][c] * inv[1] + v[2][p] * g[2][c] * inv[2];   // V U^T (:146-147)
  };
  make_R();
  if (refl != 0) {                                           // tools.py:149-157
    const float det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                      R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    const bool have = det < 0.f, want = refl == 2;
    if (have != want) {
      int last = 0;                                          // numpy's svd sorts descending: "[-1]" is the smallest
      if (sv[1] < sv[last]) last = 1;
      if (sv[2] < sv[last]) last = 2;
      for (int p = 0; p < 3; ++p) v[last][p] = -v[last][p];
      make_R();
      sv[last] = -sv[last];
    }
  }
  const float tr = sv[0] + sv[1] + sv[2];
  float scale, d, zs;
  if (scaling) {
    scale = tr * nA / nB;                                    // :161
    d = 1.f - tr * tr;                                       // :164
    zs = nA * tr / nB;                                       // Z = A_norm * S_trace * (B0 / B_norm) R + A_bar (:167)
  } else {
    scale = 1.f;
    d = 1.f + ssY / ssX - 2.f * tr * nB / nA;                // :170
    zs = 1.f;                                                // Z = B_norm * (B0 / B_norm) R + A_bar (:171)
  }
  if (Z != nullptr)
    for (int j = 0; j < LCN_J; ++j) {
      float y[3];
      for (int c = 0; c < 3; ++c) y[c] = b[j * 3 + c] - bb[c];
      for (int c = 0; c < 3; ++c) Z[ip * EV_POSE + j * 3 + c] = zs * (y[0] * R[0][c] + y[1] * R[1][c] + y[2] * R[2][c]) + ab[c];
    }
  if (tform != nullptr) {
    float* t = tform + ip * 14;
    for (int p = 0; p < 3; ++p)
      for (int c = 0; c < 3; ++c) t[p * 3 + c] = R[p][c];
    t[9] = scale;
    for (int c = 0; c < 3; ++c) t[10 + c] = ab[c] - scale * (bb[0] * R[0][c] + bb[1] * R[1][c] + bb[2] * R[2][c]);   // :176
    t[13] = d;
  }
}
extern "C" int lcn_procrustes(const float* d_A, const float* d_B, int64_t n, int scaling, int reflection, float* d_Z,
                              float* d_tform, void* stream) {
  LCN_REQUIRE(d_A && d_B && (d_Z || d_tform), "null argument");
  LCN_REQUIRE(n > 0, "n must be positive");
  LCN_REQUIRE(reflection >= 0 && reflection <= 2, "reflection: 0 = 'best', 1 = False, 2 = True");
  const int flags = (scaling ? 1 : 0) | (reflection << 1);
  k_procrustes<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_A, d_B, n, flags, d_Z, d_tform);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

// (f) DataReader.denormalize, tools/data.py:471-472
__global__ void k_denorm(float* __restrict__ pose, const float* __restrict__ res, int64_t n) {
  int64_t total = n * EV_POSE;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t ip = e / EV_POSE;
    int c = (int)(e % 3);
    float w = res[ip * 2], h = res[ip * 2 + 1];
    float v = pose[e];
    pose[e] = c == 0 ? (v + 1.0f) * w * 0.5f : c == 1 ? (v + h / w) * w * 0.5f : v * w * 0.5f;
  }
}
extern "C" int lcn_denormalize(float* d_pose, const float* d_res, int64_t n, void* stream) {
  LCN_REQUIRE(d_pose && d_res && n > 0, "bad argument");
  int grid = (int)std::min<int64_t>((n * EV_POSE + 255) / 256, 148 * 16);
  k_denorm<<<grid, 256, 0, (cudaStream_t)stream>>>(d_pose, d_res, n);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}
