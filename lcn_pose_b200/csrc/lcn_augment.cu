// (f2) Pose augmentations and the test-time-augmentation "undo" on the device: tools/data.py:10-25 (flip_data),
// :27-54 (translation_data, scalar branch), :289-322 (rotate_data: about z, pivot = joint 0), :205-249,269-287
// (untranslation_data / unflip_data / undo).  The reference runs them as NumPy copies and Python per-item loops on the
// host (train.py:85-104, inference.py:75-111); here each is one elementwise pass, thread = (pose, joint), HBM bound:
// 2 * 17*k*4 B per pose for an augmentation, (ops+2) * 204 B per pose for undo.
#include <algorithm>
#include <math.h>

#include "lcn_internal.cuh"

// joint that supplies output joint j under a horizontal flip: left [4,5,6,11,12,13] <-> right [1,2,3,14,15,16]
__device__ __forceinline__ int flip_src(int j) {
  const int perm[LCN_J] = {0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 14, 15, 16, 11, 12, 13};
  return perm[j];
}

template <int K>
__global__ void __launch_bounds__(256) k_augment(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int op,
                                                 float cs, float sn, float t) {
  lcn_pdl_prologue();
  const int64_t total = n * LCN_J;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ip = e / LCN_J;
    const int j = (int)(e - ip * LCN_J);
    const float* pose = src + ip * (LCN_J * K);
    float v[3] = {0.f, 0.f, 0.f};
    if (op == LCN_AUG_FLIP) {
      const int js = flip_src(j);
#pragma unroll
      for (int c = 0; c < K; ++c) v[c] = pose[js * K + c];
      v[0] = -v[0];
    } else if (op == LCN_AUG_ROTATE) {
      // (p - pivot) @ Rz^T + pivot: x' = c x - s y, y' = s x + c y; z (k = 3) is unchanged
      const float px = pose[0], py = pose[1];
      const float x = pose[j * K] - px, y = pose[j * K + 1] - py;
      v[0] = cs * x - sn * y + px;
      v[1] = sn * x + cs * y + py;
      if (K == 3) v[2] = pose[j * K + 2];
    } else {
#pragma unroll
      for (int c = 0; c < K; ++c) v[c] = pose[j * K + c] + t;
    }
#pragma unroll
    for (int c = 0; c < K; ++c) dst[ip * (LCN_J * K) + j * K + c] = v[c];
  }
}

extern "C" int lcn_augment(const float* d_src, float* d_dst, int64_t n, int k, int op, float angle_deg, float t,
                           void* stream) {
  LCN_REQUIRE(d_src && d_dst && n > 0, "bad argument");
  LCN_REQUIRE(d_src != d_dst, "lcn_augment is out of place (a flip / rotation reads other joints of the pose)");
  LCN_REQUIRE(k == 2 || k == 3, "k=%d: 2 or 3 coordinates per joint", k);
  LCN_REQUIRE(op == LCN_AUG_FLIP || op == LCN_AUG_ROTATE || op == LCN_AUG_TRANSLATE, "unknown augmentation %d", op);
  const double th = (double)angle_deg * 3.14159265358979323846 / 180.0;
  const float cs = (float)cos(th), sn = (float)sin(th);
  const int grid = (int)std::min<int64_t>((n * LCN_J + 255) / 256, 148 * 16);
  if (k == 2) lcn_launch(k_augment<2>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, d_src, d_dst, n, op, cs, sn, t);
  else lcn_launch(k_augment<3>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, d_src, d_dst, n, op, cs, sn, t);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

// undo (tools/data.py:269-287): preds [(n_ops+1), n, 17, 3]; slice_f is un-flipped, slice_r rotated by `angle` about
// joint 0 of ITS OWN pose (an inverse only for the default 180 degrees, as in the reference), slice_t un-translated;
// mean over the n_ops+1 slices.
__global__ void __launch_bounds__(256) k_tta_undo(const float* __restrict__ preds, float* __restrict__ out, int64_t n,
                                                  int n_ops, int slice_f, int slice_r, int slice_t, float cs, float sn,
                                                  float t) {
  lcn_pdl_prologue();
  const int64_t total = n * LCN_J;
  const float inv = 1.f / (float)(n_ops + 1);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ip = e / LCN_J;
    const int j = (int)(e - ip * LCN_J);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int s = 0; s <= n_ops; ++s) {
      const float* pose = preds + ((int64_t)s * n + ip) * (LCN_J * 3);
      float x, y, z;
      if (s == slice_f) {
        const int js = flip_src(j);
        x = -pose[js * 3]; y = pose[js * 3 + 1]; z = pose[js * 3 + 2];
      } else if (s == slice_r) {
        const float px = pose[0], py = pose[1];
        const float dx = pose[j * 3] - px, dy = pose[j * 3 + 1] - py;
        x = cs * dx - sn * dy + px;
        y = sn * dx + cs * dy + py;
        z = pose[j * 3 + 2];
      } else if (s == slice_t) {
        x = pose[j * 3] - t; y = pose[j * 3 + 1] - t; z = pose[j * 3 + 2] - t;
      } else {
        x = pose[j * 3]; y = pose[j * 3 + 1]; z = pose[j * 3 + 2];
      }
      a0 += x; a1 += y; a2 += z;
    }
    out[ip * 51 + j * 3] = a0 * inv;
    out[ip * 51 + j * 3 + 1] = a1 * inv;
    out[ip * 51 + j * 3 + 2] = a2 * inv;
  }
}

extern "C" int lcn_tta_undo(const float* d_preds, float* d_out, int64_t n, int n_ops, int slice_f, int slice_r,
                            int slice_t, float angle_deg, float t, void* stream) {
  LCN_REQUIRE(d_preds && d_out && n > 0, "bad argument");
  LCN_REQUIRE(n_ops >= 0 && n_ops <= 8, "n_ops=%d outside 0..8", n_ops);
  LCN_REQUIRE(slice_f <= n_ops && slice_r <= n_ops && slice_t <= n_ops, "slice index beyond n_ops");
  const double th = (double)angle_deg * 3.14159265358979323846 / 180.0;
  const int grid = (int)std::min<int64_t>((n * LCN_J + 255) / 256, 148 * 16);
  lcn_launch(k_tta_undo, dim3(grid), dim3(256), 0, (cudaStream_t)stream, d_preds, d_out, n, n_ops, slice_f, slice_r, slice_t,
             (float)cos(th), (float)sin(th), t);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}
