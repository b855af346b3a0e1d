// Stand-alone ops of the reference's layer API and of the data path either side of the LCN stack (include/lcn_b200.h):
//   lcn_mask_weights     cgcnn.mask_weights              network/models_att.py:576-586
//   lcn_batch_norm       cgcnn.batch_normalization_warp  network/models_att.py:588-612 (Keras BN, batch statistics)
//   lcn_mse_loss         base_model.loss                 network/models_att.py:352-366 (mse + regularization * sum l2_loss)
//   lcn_l2_regularizer   base_model._variable            network/models_att.py:465-472 (tf.nn.l2_loss of every w*, b*)
//   lcn_loss_ema         ExponentialMovingAverage(0.9)   network/models_att.py:370-379 (zero-debiased, every step)
//   lcn_gather_rows      train_data[idx] of fit          network/models_att.py:200
//   lcn_normalize        DataReader.read_2d / read_3d    tools/data.py:338-445
// The hot path itself (the fused layer stack, its backward pass and Adam) lives in lcn_kernels.cu / lcn_gemm_tc.cu /
// lcn_stack_tc.cu; these are the per-op entry points the reference's method-level API maps to, plus the HBM-bound
// row movers that keep fit() / predict() / evaluate() free of host round trips.
#include "lcn_internal.cuh"

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the block; valid in every thread.  sh: >= 33 doubles
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = threadIdx.x < nw ? sh[threadIdx.x] : 0.0;
    t = warp_sum_d(t);
    if (threadIdx.x == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// masked = reshape(w, [J, Fi, J, Fo]) * mask[J, 1, J, 1]
__global__ void __launch_bounds__(256) k_mask_weights(const float* __restrict__ w, int Fi, int Fo,
                                                      const float* __restrict__ mask, float* __restrict__ out) {
  lcn_pdl_prologue();
  const int Kout = LCN_J * Fo;
  const int64_t n = (int64_t)LCN_J * Fi * Kout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / Kout), c = (int)(e - (int64_t)r * Kout);
    out[e] = w[e] * mask[(r / Fi) * LCN_J + c / Fo];
  }
}

// Keras BatchNormalization(axis=-1) on reshape(y, [-1, 17, F]) with batch statistics: per channel f the mean and the
// biased variance over rows x 17 joints (two passes, like tf.nn.moments), y' = gamma (y - mean) / sqrt(var + eps) + beta.
// One block per channel.
__global__ void __launch_bounds__(256) k_batch_norm(const float* __restrict__ y, int64_t rows, int F,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    float eps, float* __restrict__ out, float* __restrict__ stats) {
  lcn_pdl_prologue();
  __shared__ double sh[33];
  const int f = blockIdx.x, P = LCN_J * F;
  const int64_t n = rows * LCN_J;
  double s = 0.0;
  for (int64_t e = threadIdx.x; e < n; e += blockDim.x) s += (double)y[(e / LCN_J) * P + (e % LCN_J) * F + f];
  const double mean = block_sum_d(s, sh) / (double)n;
  double q = 0.0;
  for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
    const double d = (double)y[(e / LCN_J) * P + (e % LCN_J) * F + f] - mean;
    q += d * d;
  }
  const double var = block_sum_d(q, sh) / (double)n;
  const float mu = (float)mean, rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma[f], b = beta[f];
  for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
    const size_t o = (size_t)(e / LCN_J) * P + (e % LCN_J) * F + f;
    out[o] = fmaf((y[o] - mu) * rstd, g, b);
  }
  if (stats != nullptr && threadIdx.x == 0) {
    stats[2 * f] = mu;
    stats[2 * f + 1] = (float)var;
  }
}

// out[0] = mean((a - b)^2) over n elements (+ reg_scale * reg_in[0] when reg_in != NULL).  One block.
__global__ void __launch_bounds__(1024) k_mse_loss(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                   const float* __restrict__ reg_in, float reg_scale,
                                                   float* __restrict__ out) {
  lcn_pdl_prologue();
  __shared__ double sh[33];
  double s = 0.0;
  for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
    const double d = (double)a[e] - (double)b[e];
    s += d * d;
  }
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) {
    double v = s / (double)n;
    if (reg_in != nullptr) v += (double)reg_scale * (double)reg_in[0];
    out[0] = (float)v;
  }
}

struct RegTable {
  int n;
  int64_t off[2 * LCN_MAX_LIN], size[2 * LCN_MAX_LIN];
};
// out[0] = sum over the regularised tensors (every w*, b*) of sum(v^2) / 2.  One block per tensor + a last-block finish.
__global__ void __launch_bounds__(1024) k_l2_regularizer(const float* __restrict__ params, RegTable rt,
                                                         double* __restrict__ acc /* [2]: sum, ticket */,
                                                         float* __restrict__ out) {
  lcn_pdl_prologue();
  __shared__ double sh[33];
  const int t = blockIdx.x;
  const float* v = params + rt.off[t];
  double s = 0.0;
  for (int64_t e = threadIdx.x; e < rt.size[t]; e += blockDim.x) s += (double)v[e] * v[e];
  s = block_sum_d(s, sh);
  if (threadIdx.x == 0) {
    atomicAdd(acc, 0.5 * s);
    __threadfence();
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(acc + 1), 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      out[0] = (float)*reinterpret_cast<volatile double*>(acc);
      *reinterpret_cast<volatile double*>(acc) = 0.0;          // self-resetting: ready for the next launch
      *reinterpret_cast<volatile unsigned*>(acc + 1) = 0u;
    }
  }
}

// ema[0] = decay * ema[0] + (1 - decay) * (loss + reg_scale * reg); ema[1] += 1  (the zero-debias factor
// 1 - decay^ema[1] is applied by the reader)
__global__ void k_loss_ema(const float* __restrict__ loss, const float* __restrict__ reg, float reg_scale, float decay,
                           float* __restrict__ ema) {
  lcn_pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float v = loss[0];
    if (reg != nullptr) v = fmaf(reg_scale, reg[0], v);
    ema[0] = decay * ema[0] + (1.f - decay) * v;
    ema[1] += 1.f;
  }
}

// dst_k[b, :] = src_k[idx[b], :] for up to two row sets that share the index vector (inputs [N, 34] and labels [N, 51]).
// 16-byte vectors when both row pitches allow it is not worth the divergence here: rows are 136 / 204 bytes, a warp
// reads whole rows with consecutive lanes -> fully coalesced 4-byte accesses; HBM/L2 bound, 2 x 340 B per pose.
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ src_a, int cols_a, float* __restrict__ dst_a,
                                                     const float* __restrict__ src_b, int cols_b, float* __restrict__ dst_b,
                                                     const int64_t* __restrict__ idx, int64_t n_idx, int64_t n_src) {
  lcn_pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < n_idx; b += nwarps) {
    int64_t r = idx[b];
    r = r < 0 ? 0 : (r >= n_src ? n_src - 1 : r);             // out-of-range indices are clamped, never read outside
    for (int c = lane; c < cols_a; c += 32) dst_a[b * cols_a + c] = src_a[r * cols_a + c];
    if (src_b != nullptr)
      for (int c = lane; c < cols_b; c += 32) dst_b[b * cols_b + c] = src_b[r * cols_b + c];
  }
}

// read_2d: xy / res_w * 2 - [1, res_h / res_w]; read_3d: the same for xy, z / res_w * 2.  One thread per joint.
__global__ void __launch_bounds__(256) k_normalize(const float* __restrict__ j3d, const float* __restrict__ res, int64_t n,
                                                   float* __restrict__ x2d, float* __restrict__ y3d) {
  lcn_pdl_prologue();
  const int64_t total = n * LCN_J;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = e / LCN_J;
    const float w = res[2 * p], h = res[2 * p + 1];
    const float u = j3d[3 * e] / w * 2.f - 1.f;
    const float v = j3d[3 * e + 1] / w * 2.f - h / w;
    if (x2d != nullptr) {
      x2d[2 * e] = u;
      x2d[2 * e + 1] = v;
    }
    if (y3d != nullptr) {
      y3d[3 * e] = u;
      y3d[3 * e + 1] = v;
      y3d[3 * e + 2] = j3d[3 * e + 2] / w * 2.f;
    }
  }
}

int grid_for(int64_t work_items, int per_block) {
  int64_t g = (work_items + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

extern "C" int lcn_mask_weights(const float* d_w, int32_t rows, int32_t cols, const float* d_mask, float* d_out,
                                void* stream) {
  LCN_REQUIRE(d_w && d_mask && d_out, "null argument");
  LCN_REQUIRE(rows > 0 && cols > 0 && rows % LCN_J == 0 && cols % LCN_J == 0,
              "mask_weights: shape [%d, %d] is not a multiple of 17 joints on both sides", rows, cols);
  lcn_launch(k_mask_weights, dim3(grid_for((int64_t)rows * cols, 1024)), dim3(256), 0, (cudaStream_t)stream, d_w,
             rows / LCN_J, cols / LCN_J, d_mask, d_out);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

extern "C" int lcn_batch_norm(const float* d_y, int64_t rows, int32_t F, const float* d_gamma, const float* d_beta,
                              float eps, float* d_out, float* d_stats, void* stream) {
  LCN_REQUIRE(d_y && d_gamma && d_beta && d_out, "null argument");
  LCN_REQUIRE(rows > 0 && F > 0, "batch_norm: rows=%lld F=%d must be positive", (long long)rows, F);
  lcn_launch(k_batch_norm, dim3(F), dim3(256), 0, (cudaStream_t)stream, d_y, rows, F, d_gamma, d_beta, eps, d_out, d_stats);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

extern "C" int lcn_mse_loss(const float* d_pred, const float* d_labels, int64_t n_elems, const float* d_reg,
                            float reg_scale, float* d_loss, void* stream) {
  LCN_REQUIRE(d_pred && d_labels && d_loss && n_elems > 0, "bad argument");
  lcn_launch(k_mse_loss, dim3(1), dim3(1024), 0, (cudaStream_t)stream, d_pred, d_labels, n_elems, d_reg, reg_scale, d_loss);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

extern "C" int lcn_l2_regularizer(const lcn_model* m, const float* d_params, void* d_scratch16, float* d_out, void* stream) {
  LCN_REQUIRE(m && d_params && d_scratch16 && d_out, "null argument");
  RegTable rt;
  rt.n = 0;
  for (int s = 0; s < m->segs.n; ++s)
    if (m->segs.s[s].kind == SEG_W || m->segs.s[s].kind == SEG_B) {
      rt.off[rt.n] = m->segs.s[s].off;
      rt.size[rt.n] = m->segs.s[s].size;
      ++rt.n;
    }
  lcn_launch(k_l2_regularizer, dim3(rt.n), dim3(1024), 0, (cudaStream_t)stream, d_params, rt,
             reinterpret_cast<double*>(d_scratch16), d_out);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

extern "C" int lcn_loss_ema(const float* d_loss, const float* d_reg, float reg_scale, float decay, float* d_ema,
                            void* stream) {
  LCN_REQUIRE(d_loss && d_ema, "null argument");
  lcn_launch(k_loss_ema, dim3(1), dim3(32), 0, (cudaStream_t)stream, d_loss, d_reg, reg_scale, decay, d_ema);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

extern "C" int lcn_gather_rows(const float* d_src_a, int32_t cols_a, float* d_dst_a, const float* d_src_b, int32_t cols_b,
                               float* d_dst_b, const int64_t* d_idx, int64_t n_idx, int64_t n_src, void* stream) {
  LCN_REQUIRE(d_src_a && d_dst_a && d_idx && cols_a > 0, "null argument");
  LCN_REQUIRE((d_src_b == nullptr) == (d_dst_b == nullptr), "second row set: give both pointers or neither");
  LCN_REQUIRE(n_idx >= 0 && n_src > 0, "bad sizes");
  if (n_idx == 0) return LCN_OK;
  lcn_launch(k_gather_rows, dim3(grid_for(n_idx, 8)), dim3(256), 0, (cudaStream_t)stream, d_src_a, cols_a, d_dst_a, d_src_b,
             cols_b, d_dst_b, d_idx, n_idx, n_src);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

extern "C" int lcn_normalize(const float* d_joint_3d_image, const float* d_res, int64_t n, float* d_x2d, float* d_y3d,
                             void* stream) {
  LCN_REQUIRE(d_joint_3d_image && d_res && (d_x2d || d_y3d), "null argument");
  if (n <= 0) return LCN_OK;
  lcn_launch(k_normalize, dim3(grid_for(n * LCN_J, 256)), dim3(256), 0, (cudaStream_t)stream, d_joint_3d_image, d_res, n,
             d_x2d, d_y3d);
  LCN_CHECK_LAUNCH();
  return LCN_OK;
}

LCN_KTRACE_EXPORT(layers)
