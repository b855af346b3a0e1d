// (c) Batched evaluation: evaluate.py:53-61 per pose ->
//   tools.image_to_camera_frame (tools/tools.py:183-194)
//   [protocol 2] tools.align_to_gt -> procrustes(gt, pred) (tools/tools.py:96-181,197-202),
//                scaling=True, reflection='best' (no determinant fix: reflections are accepted)
//   err[j] = || pred_j - gt_j ||_2  (evaluate.py:61), per-joint sums, PCK@50mm (evaluate.py:78-106)
//
// HBM-bound: 444 B read (+68 B written when per-joint errors are requested) per pose.  One warp owns
// 32 consecutive poses per iteration: the 2 x 6528 B of pred/gt are staged in shared memory with
// coalesced float4 loads, then each lane works on its own pose out of shared memory (row stride 51
// words, odd -> conflict free) with the 3x3 SVD (one-sided Jacobi) entirely in registers.
#include <algorithm>

#include "lcn_internal.cuh"

#define EV_WARPS 4
#define EV_POSE 51

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// one Jacobi rotation orthogonalising columns p,q of G (3x3, column vectors) and accumulating V
__device__ __forceinline__ void jacobi_rot(float* gp, float* gq, float* vp, float* vq) {
  float alpha = gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2];
  float beta = gq[0] * gq[0] + gq[1] * gq[1] + gq[2] * gq[2];
  float gamma = gp[0] * gq[0] + gp[1] * gq[1] + gp[2] * gq[2];
  if (gamma * gamma <= 1e-14f * alpha * beta) return;
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
}

__global__ void __launch_bounds__(EV_WARPS * 32) k_eval(const float* __restrict__ pred, const float* __restrict__ gt,
                                                       const float* __restrict__ box, const float* __restrict__ cam,
                                                       const float* __restrict__ root_depth,
                                                       const int32_t* __restrict__ action, int n_actions, int64_t n,
                                                       int flags, float* __restrict__ err_out,
                                                       float* __restrict__ pose_out, double* __restrict__ sums) {
  const int protocol2 = flags & 1, camera_frame = flags & 2;
  extern __shared__ __align__(16) float smem[];
  // per warp: pred[32*51], gt[32*51]; then per block: action sums (double)
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sp = smem + warp * (2 * 32 * EV_POSE);
  float* sg = sp + 32 * EV_POSE;
  double* asum = reinterpret_cast<double*>(smem + EV_WARPS * 2 * 32 * EV_POSE);   // [n_actions][19]
  for (int e = threadIdx.x; e < n_actions * 19; e += blockDim.x) asum[e] = 0.0;
  __syncthreads();

  double jsum[LCN_J];
#pragma unroll
  for (int j = 0; j < LCN_J; ++j) jsum[j] = 0.0;
  unsigned long long npose = 0, npck = 0;

  int64_t n_chunks = (n + 31) / 32;
  int64_t wglobal = (int64_t)blockIdx.x * EV_WARPS + warp, wstride = (int64_t)gridDim.x * EV_WARPS;
  for (int64_t ch = wglobal; ch < n_chunks; ch += wstride) {
    int64_t p0 = ch * 32;
    int cnt = (int)min((int64_t)32, n - p0);
    const float* gp = pred + p0 * EV_POSE;
    const float* gg = gt + p0 * EV_POSE;
    __syncwarp();
    if (cnt == 32) {
#pragma unroll 4
      for (int v = lane; v < 32 * EV_POSE / 4; v += 32) {
        *reinterpret_cast<float4*>(sp + v * 4) = ld_stream4(gp + v * 4);
        *reinterpret_cast<float4*>(sg + v * 4) = ld_stream4(gg + v * 4);
      }
    } else {
      for (int v = lane; v < cnt * EV_POSE; v += 32) {
        sp[v] = gp[v];
        sg[v] = gg[v];
      }
    }
    __syncwarp();
    bool active = lane < cnt;
    float e[LCN_J];
#pragma unroll
    for (int j = 0; j < LCN_J; ++j) e[j] = 0.f;
    int act = -1;
    if (active) {
      int64_t ip = p0 + lane;
      float4 bx = make_float4(0.f, 0.f, 1999.f, 0.f), cm = make_float4(1.f, 1.f, 0.f, 0.f);
      float rd = 0.f;
      if (!camera_frame) {
        bx = *reinterpret_cast<const float4*>(box + ip * 4);
        cm = *reinterpret_cast<const float4*>(cam + ip * 4);   // fx, fy, cx, cy
        rd = root_depth[ip];
      }
      if (action != nullptr) act = action[ip];
      float inv_ratio = 2000.0f / (bx.z - bx.x + 1.0f);
      float ifx = 1.0f / cm.x, ify = 1.0f / cm.y;
      const float* mp = sp + lane * EV_POSE;
      const float* mg = sg + lane * EV_POSE;
      float P[LCN_J][3], G[LCN_J][3];
      float pb[3] = {0, 0, 0}, gb[3] = {0, 0, 0};
#pragma unroll
      for (int j = 0; j < LCN_J; ++j) {
        if (camera_frame) {
          P[j][0] = mp[j * 3 + 0]; P[j][1] = mp[j * 3 + 1]; P[j][2] = mp[j * 3 + 2];
        } else {
          float z = mp[j * 3 + 2] * inv_ratio + rd;          // tools.py:187
          P[j][0] = (mp[j * 3 + 0] - cm.z) * ifx * z;        // tools.py:190,192
          P[j][1] = (mp[j * 3 + 1] - cm.w) * ify * z;        // tools.py:191,193
          P[j][2] = z;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          G[j][c] = mg[j * 3 + c];
          pb[c] += P[j][c];
          gb[c] += G[j][c];
        }
      }
      if (protocol2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          pb[c] *= (1.0f / LCN_J);
          gb[c] *= (1.0f / LCN_J);
        }
        float ssA = 0.f, ssB = 0.f;
        float M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // M[a][b] = sum_k A0[k][a] * B0[k][b]
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            P[j][c] -= pb[c];                               // B0 (prediction), tools.py:130
            G[j][c] -= gb[c];                               // A0 (ground truth), tools.py:129
            ssA = fmaf(G[j][c], G[j][c], ssA);
            ssB = fmaf(P[j][c], P[j][c], ssB);
          }
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) M[a][b] = fmaf(G[j][a], P[j][b], M[a][b]);
        }
        float nA = sqrtf(ssA), nB = sqrtf(ssB);
        float inv_ab = 1.0f / (nA * nB);
        // columns of M (scaled to the unit-norm problem of tools.py:133-144)
        float g0[3], g1[3], g2[3], v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          g0[a] = M[a][0] * inv_ab;
          g1[a] = M[a][1] * inv_ab;
          g2[a] = M[a][2] * inv_ab;
        }
        // one-sided Jacobi: M V = U S  (np.linalg.svd, tools.py:145)
#pragma unroll 1
        for (int sweep = 0; sweep < 6; ++sweep) {
          jacobi_rot(g0, g1, v0, v1);
          jacobi_rot(g0, g2, v0, v2);
          jacobi_rot(g1, g2, v1, v2);
        }
        float s0 = sqrtf(g0[0] * g0[0] + g0[1] * g0[1] + g0[2] * g0[2]);
        float s1 = sqrtf(g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2]);
        float s2 = sqrtf(g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2]);
        float tr = s0 + s1 + s2;                            // S_trace, tools.py:159
        float i0 = s0 > 1e-20f ? 1.f / s0 : 0.f, i1 = s1 > 1e-20f ? 1.f / s1 : 0.f, i2 = s2 > 1e-20f ? 1.f / s2 : 0.f;
        // R = V U^T (tools.py:146-147): R[b][c] = sum_s V[b][s] U[c][s], U[:,s] = g_s / s_s
        float R[3][3];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c) R[b][c] = v0[b] * g0[c] * i0 + v1[b] * g1[c] * i1 + v2[b] * g2[c] * i2;
        // Z - gt = nA*tr*(B0/nB) R - A0   (tools.py:168 minus the common centroid)
        float sc = nA * tr / nB;
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) {
          float d2 = 0.f, zc[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            zc[c] = sc * (P[j][0] * R[0][c] + P[j][1] * R[1][c] + P[j][2] * R[2][c]) - G[j][c];
            d2 = fmaf(zc[c], zc[c], d2);
          }
          e[j] = sqrtf(d2);
#pragma unroll
          for (int c = 0; c < 3; ++c) P[j][c] = zc[c] + G[j][c] + gb[c];   // aligned pose Z (tools.py:168)
        }
      } else {
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) {
          float dx = P[j][0] - G[j][0], dy = P[j][1] - G[j][1], dz = P[j][2] - G[j][2];
          e[j] = sqrtf(dx * dx + dy * dy + dz * dz);
        }
      }
      if (pose_out != nullptr) {   // every lane only rewrites its own 51 words of the staging area
        float* mo = sp + lane * EV_POSE;
#pragma unroll
        for (int j = 0; j < LCN_J; ++j) {
          mo[j * 3 + 0] = P[j][0]; mo[j * 3 + 1] = P[j][1]; mo[j * 3 + 2] = P[j][2];
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
  size_t smem = (size_t)EV_WARPS * 2 * 32 * EV_POSE * sizeof(float) + (size_t)n_actions * 19 * sizeof(double);
  static bool attr = false;
  if (!attr) {
    LCN_CHECK_CUDA(cudaFuncSetAttribute(k_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t chunks = (n + 31) / 32;
  int64_t blocks = (chunks + EV_WARPS - 1) / EV_WARPS;
  int grid = (int)std::min<int64_t>(blocks, (int64_t)sms * 4);
  k_eval<<<grid, EV_WARPS * 32, smem, (cudaStream_t)stream>>>(d_pred, d_gt, d_box, d_cam, d_root_depth, d_action,
                                                            n_actions, n, flags, d_err, d_pose_out, d_sums);
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
