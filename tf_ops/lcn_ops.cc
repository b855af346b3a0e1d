// TensorFlow custom ops over the C ABI of liblcn_b200.so (include/lcn_b200.h) -- the second host binding of the same
// kernels (the first is the ctypes binding lcn_pose_b200/_lib.py that every test and bench.py uses).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: TensorFlow 2.13 (requirements.txt:1 of the reference) has no wheel for the
// image's Python 3.12 and there is no network, so there are no TensorFlow headers to build against.  `make -C tf_ops`
// builds it wherever `python -c "import tensorflow"` succeeds and is a no-op with a message otherwise.  The C ABI calls
// below are exactly the ones the ctypes path makes, in the same order (lcn_pose_b200/engine.py).
//
// What it replaces in the reference's graph (network/models_att.py):
//   LcnForward     cgcnn._inference_lcn (:707-775)             logits = lcn_forward(params, data, dropout, step)
//   LcnLoss        base_model.loss (:352-366) + the gradient   loss, grads = lcn_loss(params, data, labels, dropout, step)
//                  that optimizer.compute_gradients builds (:408); tf.RegisterGradient("LcnLoss") in lcn_tf.py hands
//                  `grads` to any stock optimizer
//   LcnAdam        AdamOptimizer.apply_gradients (:404-409)    fused TF1 Adam on the flat parameter variable
//   LcnWeightPrep  clip_by_norm + mask_weights (:576-586,659)  explicit re-preparation (LcnForward / LcnLoss / LcnAdam do it themselves)
//   LcnEval        evaluate.py:53-61                            per-joint errors of a pose batch
// All variables of the reference (mask, w*, b*, BN gamma / beta) live in ONE flat float32 resource variable `params`
// laid out by lcn_model_tensor_info (names = the reference's variable names); Adam's m / v have the same layout.
//
// Every op: device pointers only, stream from the OpKernelContext, no synchronisation, no allocation inside the
// library (workspaces come from the TF allocator), errors surface as Status (no exceptions cross the C ABI).
#include <cstring>
#include <string>
#include <vector>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/resource_mgr.h"
#include "tensorflow/core/framework/resource_var.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/platform/stream_executor.h"

#define EIGEN_USE_GPU
#include "lcn_b200.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;

namespace {

#define LCN_MODEL_ATTRS                                                                                       \
  ".Attr(\"F: int = 64\").Attr(\"in_F: int = 2\").Attr(\"num_layers: int = 3\")"                             \
  ".Attr(\"mask_type: string = 'locally_connected'\").Attr(\"residual: bool = true\")"                       \
  ".Attr(\"max_norm: bool = true\").Attr(\"path: string = 'bf16'\").Attr(\"support: list(float)\")"          \
  ".Attr(\"const_mask: list(float) = []\")"

// Model handle built from the op's attributes (host tables only; cheap, immutable after creation).
class ModelHolder {
 public:
  tf::Status Init(tf::OpKernelConstruction* c) {
    lcn_model_desc d;
    std::memset(&d, 0, sizeof(d));
    int F, in_F, L;
    bool residual, max_norm;
    std::string mask_type, path;
    std::vector<float> support, const_mask;
    TF_RETURN_IF_ERROR(c->GetAttr("F", &F));
    TF_RETURN_IF_ERROR(c->GetAttr("in_F", &in_F));
    TF_RETURN_IF_ERROR(c->GetAttr("num_layers", &L));
    TF_RETURN_IF_ERROR(c->GetAttr("mask_type", &mask_type));
    TF_RETURN_IF_ERROR(c->GetAttr("residual", &residual));
    TF_RETURN_IF_ERROR(c->GetAttr("max_norm", &max_norm));
    TF_RETURN_IF_ERROR(c->GetAttr("path", &path));
    TF_RETURN_IF_ERROR(c->GetAttr("support", &support));
    TF_RETURN_IF_ERROR(c->GetAttr("const_mask", &const_mask));
    if (support.size() != LCN_JOINTS * LCN_JOINTS) return tf::errors::InvalidArgument("support must have 289 entries");
    d.F = F; d.in_F = in_F; d.num_layers = L;
    d.mask_kind = mask_type.find("exponential") != std::string::npos ? LCN_MASK_CONSTANT : LCN_MASK_LOCALLY_CONNECTED;
    d.residual = residual; d.batch_norm = 1; d.max_norm = max_norm;
    d.path = path == "bf16" ? LCN_PATH_BF16 : LCN_PATH_FP32;
    for (int i = 0; i < LCN_JOINTS * LCN_JOINTS; ++i) {
      d.support[i] = support[i];
      d.const_mask[i] = i < (int)const_mask.size() ? const_mask[i] : 0.f;
    }
    if (lcn_model_create(&d, &m_) != LCN_OK) return tf::errors::InvalidArgument(lcn_last_error());
    return tf::OkStatus();
  }
  ~ModelHolder() { lcn_model_destroy(m_); }
  lcn_model* get() const { return m_; }

 private:
  lcn_model* m_ = nullptr;
};

inline void* StreamOf(tf::OpKernelContext* ctx) {
  return reinterpret_cast<void*>(ctx->eigen_gpu_device().stream());     // cudaStream_t of the op's device
}
#define LCN_OK_OR_RETURN(ctx, expr) \
  OP_REQUIRES(ctx, (expr) == LCN_OK, tf::errors::Internal("liblcn_b200: ", lcn_last_error()))

// The flat parameter vector lives in a resource variable; Adam updates it in place.
tf::Status LockedVarTensor(tf::OpKernelContext* ctx, int input, tf::core::RefCountPtr<tf::Var>* var) {
  TF_RETURN_IF_ERROR(tf::LookupResource(ctx, tf::HandleFromInput(ctx, input), var));
  return tf::OkStatus();
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
REGISTER_OP("LcnWeightPrep")
    .Input("params: float")
    .Output("workspace: uint8") LCN_MODEL_ATTRS
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->Vector(InferenceContext::kUnknownDim)); return tf::OkStatus(); })
    .Doc("clip_by_norm + mask_weights + pack for every layer (models_att.py:534-586,659-660) into a fresh workspace head.");

class LcnWeightPrepOp : public tf::OpKernel {
 public:
  explicit LcnWeightPrepOp(tf::OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, model_.Init(c)); }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& params = ctx->input(0);
    const size_t bytes = lcn_model_workspace_bytes(model_.get(), 128, 128, 0);
    tf::Tensor* ws = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({(int64_t)bytes}), &ws));
    LCN_OK_OR_RETURN(ctx, lcn_model_prepare_weights(model_.get(), params.flat<float>().data(), ws->flat<tf::uint8>().data(),
                                                    bytes, StreamOf(ctx)));
  }

 private:
  ModelHolder model_;
};
REGISTER_KERNEL_BUILDER(Name("LcnWeightPrep").Device(tf::DEVICE_GPU), LcnWeightPrepOp);

// ------------------------------------------------------------------------------------------------------------------
REGISTER_OP("LcnForward")
    .Input("params: float")
    .Input("data: float")          // [B, 17*in_F]  (ph_data, models_att.py:298-300)
    .Input("dropout: float")       // scalar, host memory (ph_dropout, :304)
    .Input("step: int64")          // scalar, host memory: dropout stream position (global_step + 1)
    .Output("logits: float")       // [B, 51]
    .Attr("bn_group: int = 0")     // 0: the whole batch is one BatchNorm group (the reference's graph)
    .Attr("seed: int = 2019") LCN_MODEL_ATTRS
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Matrix(c->Dim(c->input(1), 0), 51));
      return tf::OkStatus();
    })
    .Doc("cgcnn._inference_lcn (models_att.py:707-775): the whole LCN stack on one batch.");

class LcnForwardOp : public tf::OpKernel {
 public:
  explicit LcnForwardOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, model_.Init(c));
    OP_REQUIRES_OK(c, c->GetAttr("bn_group", &bn_group_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& params = ctx->input(0);
    const tf::Tensor& x = ctx->input(1);
    const float dropout = ctx->input(2).scalar<float>()();
    const tf::int64 step = ctx->input(3).scalar<tf::int64>()();
    const int64_t n = x.dim_size(0);
    const int bn = bn_group_ > 0 ? bn_group_ : (int)n;
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n, 51}), &out));
    const int training = dropout > 0.f;      // layout only: BatchNorm uses batch statistics either way
    const size_t bytes = lcn_model_workspace_bytes(model_.get(), n, bn, training);
    OP_REQUIRES(ctx, bytes > 0, tf::errors::InvalidArgument(lcn_last_error()));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({(int64_t)bytes}), &ws));
    void* st = StreamOf(ctx);
    LCN_OK_OR_RETURN(ctx, lcn_model_prepare_weights(model_.get(), params.flat<float>().data(), ws.flat<tf::uint8>().data(), bytes, st));
    LCN_OK_OR_RETURN(ctx, lcn_model_forward(model_.get(), params.flat<float>().data(), ws.flat<tf::uint8>().data(), bytes,
                                            x.flat<float>().data(), n, bn, training, dropout, (uint64_t)seed_, (uint64_t)step,
                                            out->flat<float>().data(), /*d_dyn=*/nullptr, st));
  }

 private:
  ModelHolder model_;
  int bn_group_;
  tf::int64 seed_;
};
REGISTER_KERNEL_BUILDER(Name("LcnForward").Device(tf::DEVICE_GPU).HostMemory("dropout").HostMemory("step"), LcnForwardOp);

// ------------------------------------------------------------------------------------------------------------------
REGISTER_OP("LcnLoss")
    .Input("params: float")
    .Input("data: float")          // [B, 17*in_F]
    .Input("labels: float")        // [B, 51]  (ph_labels, :301-303)
    .Input("dropout: float")
    .Input("step: int64")
    .Output("loss: float")         // scalar: mean((logits - labels)^2)  (:356)
    .Output("logits: float")       // [B, 51]
    .Output("grads: float")        // [n_params]: the TRUE gradient of `loss` w.r.t. `params` (chain rule through
                                   // mask_weights, clip_by_norm and the mask softmax included)
    .Attr("seed: int = 2019") LCN_MODEL_ATTRS
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Scalar());
      c->set_output(1, c->Matrix(c->Dim(c->input(1), 0), 51));
      c->set_output(2, c->input(0));
      return tf::OkStatus();
    })
    .Doc("Forward (training), loss and its gradient in one op: base_model.loss + optimizer.compute_gradients.");

class LcnLossOp : public tf::OpKernel {
 public:
  explicit LcnLossOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, model_.Init(c));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& params = ctx->input(0);
    const tf::Tensor& x = ctx->input(1);
    const tf::Tensor& y = ctx->input(2);
    const float dropout = ctx->input(3).scalar<float>()();
    const tf::int64 step = ctx->input(4).scalar<tf::int64>()();
    const int64_t n = x.dim_size(0);
    OP_REQUIRES(ctx, y.dim_size(0) == n && y.dim_size(1) == 51, tf::errors::InvalidArgument("labels must be [B, 51]"));
    tf::Tensor *loss = nullptr, *out = nullptr, *grads = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({}), &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({n, 51}), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, params.shape(), &grads));
    const size_t bytes = lcn_model_workspace_bytes(model_.get(), n, (int)n, 1);
    OP_REQUIRES(ctx, bytes > 0, tf::errors::InvalidArgument(lcn_last_error()));
    tf::Tensor ws, raw;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({(int64_t)bytes}), &ws));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, params.shape(), &raw));
    void* st = StreamOf(ctx);
    const float* p = params.flat<float>().data();
    tf::uint8* w = ws.flat<tf::uint8>().data();
    LCN_OK_OR_RETURN(ctx, lcn_model_prepare_weights(model_.get(), p, w, bytes, st));
    LCN_OK_OR_RETURN(ctx, lcn_model_forward(model_.get(), p, w, bytes, x.flat<float>().data(), n, (int)n, 1, dropout,
                                            (uint64_t)seed_, (uint64_t)step, out->flat<float>().data(), nullptr, st));
    LCN_OK_OR_RETURN(ctx, lcn_model_backward(model_.get(), p, w, bytes, x.flat<float>().data(), y.flat<float>().data(), n,
                                             dropout, (uint64_t)seed_, (uint64_t)step, loss->flat<float>().data(),
                                             raw.flat<float>().data(), st));
    LCN_OK_OR_RETURN(ctx, lcn_model_finalize_grads(model_.get(), p, w, bytes, raw.flat<float>().data(),
                                                   grads->flat<float>().data(), st));
  }

 private:
  ModelHolder model_;
  tf::int64 seed_;
};
REGISTER_KERNEL_BUILDER(Name("LcnLoss").Device(tf::DEVICE_GPU).HostMemory("dropout").HostMemory("step"), LcnLossOp);

// ------------------------------------------------------------------------------------------------------------------
REGISTER_OP("LcnAdam")
    .Input("params: resource")     // flat float32 variable, updated in place
    .Input("m: resource")
    .Input("v: resource")
    .Input("data: float")
    .Input("labels: float")
    .Input("dropout: float")
    .Input("step: int64")          // t = global_step + 1 (host): dropout stream position and Adam bias correction
    .Input("learning_rate: float") // decayed learning rate of this step (exponential_decay, :392-399), host scalar
    .Output("loss: float")
    .Attr("beta1: float = 0.9").Attr("beta2: float = 0.999").Attr("epsilon: float = 1e-8")
    .Attr("regularization: float = 0.0")
    .Attr("seed: int = 2019") LCN_MODEL_ATTRS
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->Scalar()); return tf::OkStatus(); })
    .Doc("One whole train step, op_train of the reference: forward, loss, backward, TF1 Adam (models_att.py:404-409).");

class LcnAdamOp : public tf::OpKernel {
 public:
  explicit LcnAdamOp(tf::OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, model_.Init(c));
    OP_REQUIRES_OK(c, c->GetAttr("beta1", &b1_));
    OP_REQUIRES_OK(c, c->GetAttr("beta2", &b2_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("regularization", &reg_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    tf::core::RefCountPtr<tf::Var> vp, vm, vv;
    OP_REQUIRES_OK(ctx, LockedVarTensor(ctx, 0, &vp));
    OP_REQUIRES_OK(ctx, LockedVarTensor(ctx, 1, &vm));
    OP_REQUIRES_OK(ctx, LockedVarTensor(ctx, 2, &vv));
    tf::mutex_lock lp(*vp->mu()), lm(*vm->mu()), lv(*vv->mu());
    float* p = vp->tensor()->flat<float>().data();
    float* m = vm->tensor()->flat<float>().data();
    float* v = vv->tensor()->flat<float>().data();
    const tf::Tensor& x = ctx->input(3);
    const tf::Tensor& y = ctx->input(4);
    const float dropout = ctx->input(5).scalar<float>()();
    const tf::int64 t = ctx->input(6).scalar<tf::int64>()();
    const float lr = ctx->input(7).scalar<float>()();
    const int64_t n = x.dim_size(0);
    tf::Tensor* loss = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({}), &loss));
    const size_t bytes = lcn_model_workspace_bytes(model_.get(), n, (int)n, 1);
    OP_REQUIRES(ctx, bytes > 0, tf::errors::InvalidArgument(lcn_last_error()));
    tf::Tensor ws, raw, logits;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({(int64_t)bytes}), &ws));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, vp->tensor()->shape(), &raw));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_FLOAT, tf::TensorShape({n, 51}), &logits));
    void* st = StreamOf(ctx);
    tf::uint8* w = ws.flat<tf::uint8>().data();
    // TF1 Adam step size: lr * sqrt(1 - beta2^t) / (1 - beta1^t), epsilon outside the bias correction
    const double td = (double)t;
    const float lr_t = (float)((double)lr * std::sqrt(1.0 - std::pow((double)b2_, td)) / (1.0 - std::pow((double)b1_, td)));
    LCN_OK_OR_RETURN(ctx, lcn_model_prepare_weights(model_.get(), p, w, bytes, st));
    LCN_OK_OR_RETURN(ctx, lcn_model_forward(model_.get(), p, w, bytes, x.flat<float>().data(), n, (int)n, 1, dropout,
                                            (uint64_t)seed_, (uint64_t)t, logits.flat<float>().data(), nullptr, st));
    LCN_OK_OR_RETURN(ctx, lcn_model_backward(model_.get(), p, w, bytes, x.flat<float>().data(), y.flat<float>().data(), n,
                                             dropout, (uint64_t)seed_, (uint64_t)t, loss->flat<float>().data(),
                                             raw.flat<float>().data(), st));
    LCN_OK_OR_RETURN(ctx, lcn_model_adam_step(model_.get(), p, m, v, w, bytes, raw.flat<float>().data(), lr_t, b1_, b2_, eps_,
                                              reg_, nullptr, st));
  }

 private:
  ModelHolder model_;
  float b1_, b2_, eps_, reg_;
  tf::int64 seed_;
};
REGISTER_KERNEL_BUILDER(Name("LcnAdam").Device(tf::DEVICE_GPU).HostMemory("params").HostMemory("m").HostMemory("v")
                            .HostMemory("dropout").HostMemory("step").HostMemory("learning_rate"), LcnAdamOp);

// ------------------------------------------------------------------------------------------------------------------
REGISTER_OP("LcnEval")
    .Input("pred: float")          // [n, 17, 3] image-frame predictions (after DataReader.denormalize)
    .Input("gt: float")            // [n, 17, 3] camera-frame ground truth
    .Input("box: float")           // [n, 4]
    .Input("cam: float")           // [n, 4] = fx, fy, cx, cy
    .Input("root_depth: float")    // [n]
    .Output("err: float")          // [n, 17] per-joint error in mm
    .Output("sums: double")        // [19]: 17 per-joint sums, pose count, count(err < 50 mm)
    .Attr("protocol2: bool = false")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->Matrix(c->Dim(c->input(0), 0), 17));
      c->set_output(1, c->Vector(19));
      return tf::OkStatus();
    })
    .Doc("evaluate.py:53-61 for a pose batch: image_to_camera_frame, optional Procrustes alignment, per-joint error.");

class LcnEvalOp : public tf::OpKernel {
 public:
  explicit LcnEvalOp(tf::OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("protocol2", &p2_)); }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& pred = ctx->input(0);
    const int64_t n = pred.dim_size(0);
    tf::Tensor *err = nullptr, *sums = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n, 17}), &err));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({19}), &sums));
    auto stream = ctx->eigen_gpu_device().stream();
    OP_REQUIRES(ctx, cudaMemsetAsync(sums->flat<double>().data(), 0, 19 * sizeof(double), stream) == cudaSuccess,
                tf::errors::Internal("cudaMemsetAsync failed"));
    LCN_OK_OR_RETURN(ctx, lcn_eval_mpjpe(pred.flat<float>().data(), ctx->input(1).flat<float>().data(),
                                         ctx->input(2).flat<float>().data(), ctx->input(3).flat<float>().data(),
                                         ctx->input(4).flat<float>().data(), nullptr, 0, n, p2_ ? LCN_EVAL_PROTOCOL2 : 0,
                                         err->flat<float>().data(), nullptr, sums->flat<double>().data(), stream));
  }

 private:
  bool p2_;
};
REGISTER_KERNEL_BUILDER(Name("LcnEval").Device(tf::DEVICE_GPU), LcnEvalOp);
