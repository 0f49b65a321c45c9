// TensorFlow custom-op adapter over libctc_b200.so (include/ctc_b200.h).
//
// EXPERIMENTAL, NOT BUILT OR TESTED IN THIS IMAGE: TensorFlow (headers and runtime) is not installed here and cannot be
// fetched.  Where `import tensorflow` works, tf_adapter/build.sh compiles it (tests/test_tf_adapter.py does that and runs
// the reference's known answers through it; it is skipped without TensorFlow).
// All logic lives behind the C ABI; this file only adapts TensorFlow's buffer / stream ownership to it.  Two ops:
//   CtcB200LossGrad   loss (+ d_loss-weighted gradient w.r.t. logits)   -> ctcb200_loss_grad
//   CtcB200HvpLogits  d_loss * (d2 loss / d logits2) v, matrix-free     -> ctcb200_hvp_logits
// tf_adapter/__init__.py wires them into the reference's three nested tf.custom_gradient levels
// (tf_seq2seq_losses/base_loss.py:140-184: forward_fn -> gradient_fn -> _hessian_fn, third derivative raises).
#include <cstdint>

#include "ctc_b200.h"
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#define EIGEN_USE_GPU
#include "tensorflow/core/util/gpu_kernel_helper.h"

namespace tf = tensorflow;

static tf::Status SameAsLogits(tf::shape_inference::InferenceContext* c) {
  c->set_output(c->num_outputs() - 1, c->input(1));
  if (c->num_outputs() == 2) c->set_output(0, c->Vector(c->Dim(c->input(1), 0)));
  return tf::Status();      // OK on every TensorFlow from 2.6 to 2.16 (tf::OkStatus / tf::Status::OK come and go)
}

#define CTCB200_COMMON_INPUTS                                                                   \
  .Input("labels: int32").Input("logits: float32").Input("label_length: int32")                 \
  .Input("logit_length: int32").Input("d_loss: float32")                                        \
  .Attr("blank_index: int = 0").Attr("variant: int = 0") /* 0 classic, 1 simplified */          \
  .Attr("max_label_length_plus_one: int = 0") /* 0: labels.shape[1] + 1 */

REGISTER_OP("CtcB200LossGrad")
CTCB200_COMMON_INPUTS
    .Attr("with_gradient: bool = true")    // false: the loss alone (forward_fn); grad_logits is then an empty tensor
    .Output("loss: float32")
    .Output("grad_logits: float32")
    .SetShapeFn(SameAsLogits);

REGISTER_OP("CtcB200HvpLogits")
CTCB200_COMMON_INPUTS
    .Input("v: float32")
    .Output("out: float32")
    .SetShapeFn(SameAsLogits);

class CtcB200Base : public tf::OpKernel {
 public:
  explicit CtcB200Base(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("blank_index", &blank_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("variant", &variant_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("max_label_length_plus_one", &u_));
  }

 protected:
  // tf_seq2seq_losses/base_loss.py:129-138, then the descriptor and a 256-byte aligned temp workspace of class `what`
  bool Prepare(tf::OpKernelContext* ctx, int what, ctcb200_desc* d, tf::Tensor* ws, void** ws_ptr, size_t* ws_bytes) {
    const tf::Tensor &labels = ctx->input(0), &logits = ctx->input(1), &ll = ctx->input(2), &tl = ctx->input(3);
    const bool ranks_ok = logits.dims() == 3 && labels.dims() == 2 && ll.dims() == 1 && tl.dims() == 1;
    if (!ranks_ok || logits.dim_size(0) != labels.dim_size(0) || logits.dim_size(0) != ll.dim_size(0) ||
        logits.dim_size(0) != tl.dim_size(0) || ctx->input(4).NumElements() != logits.dim_size(0)) {
      ctx->SetStatus(tf::errors::InvalidArgument("ctc_b200: rank or batch mismatch"));
      return false;
    }
    *d = ctcb200_desc{};
    d->B = static_cast<int32_t>(logits.dim_size(0));
    d->T = static_cast<int32_t>(logits.dim_size(1));
    d->V = static_cast<int32_t>(logits.dim_size(2));
    d->Lw = static_cast<int32_t>(labels.dim_size(1));
    d->blank = blank_; d->variant = variant_; d->U = u_; d->flags = 0;
    *ws_bytes = ctcb200_workspace_bytes(d, what);
    if (*ws_bytes == 0 && d->B > 0) {
      ctx->SetStatus(tf::errors::InvalidArgument("ctc_b200: unsupported shape (pass max_label_length for wide label tensors)"));
      return false;
    }
    const tf::Status st = ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<int64_t>(*ws_bytes + 256)}), ws);
    if (!st.ok()) { ctx->SetStatus(st); return false; }
    const auto base = reinterpret_cast<uintptr_t>(ws->flat<tf::uint8>().data());
    *ws_ptr = reinterpret_cast<void*>((base + 255) & ~uintptr_t(255));
    return true;
  }
  int blank_ = 0, variant_ = 0, u_ = 0;
};

class CtcB200LossGradOp : public CtcB200Base {
 public:
  explicit CtcB200LossGradOp(tf::OpKernelConstruction* ctx) : CtcB200Base(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("with_gradient", &with_gradient_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    ctcb200_desc d; tf::Tensor ws; void* ws_ptr = nullptr; size_t ws_bytes = 0;
    if (!Prepare(ctx, CTCB200_WS_LOSS_GRAD_LOGITS, &d, &ws, &ws_ptr, &ws_bytes)) return;
    tf::Tensor *loss = nullptr, *grad = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({d.B}), &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, with_gradient_ ? ctx->input(1).shape() : tf::TensorShape({0, 0, 0}), &grad));
    const int rc = ctcb200_loss_grad(&d, ctx->input(1).flat<float>().data(), ctx->input(0).flat<tf::int32>().data(),
                                     ctx->input(2).flat<tf::int32>().data(), ctx->input(3).flat<tf::int32>().data(),
                                     ctx->input(4).flat<float>().data(), loss->flat<float>().data(),
                                     with_gradient_ ? grad->flat<float>().data() : nullptr, /*grad_logprobas=*/nullptr,
                                     ws_ptr, ws_bytes, ctx->eigen_device<Eigen::GpuDevice>().stream());
    OP_REQUIRES(ctx, rc == CTCB200_OK, tf::errors::Internal("ctc_b200: ", ctcb200_strerror(rc)));
  }

 private:
  bool with_gradient_ = true;
};

class CtcB200HvpLogitsOp : public CtcB200Base {
 public:
  using CtcB200Base::CtcB200Base;
  void Compute(tf::OpKernelContext* ctx) override {
    ctcb200_desc d; tf::Tensor ws; void* ws_ptr = nullptr; size_t ws_bytes = 0;
    if (!Prepare(ctx, CTCB200_WS_HVP_LOGITS, &d, &ws, &ws_ptr, &ws_bytes)) return;
    OP_REQUIRES(ctx, ctx->input(5).shape() == ctx->input(1).shape(), tf::errors::InvalidArgument("ctc_b200: v must match logits"));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, ctx->input(1).shape(), &out));
    const int rc = ctcb200_hvp_logits(&d, ctx->input(1).flat<float>().data(), ctx->input(0).flat<tf::int32>().data(),
                                      ctx->input(2).flat<tf::int32>().data(), ctx->input(3).flat<tf::int32>().data(),
                                      ctx->input(4).flat<float>().data(), ctx->input(5).flat<float>().data(),
                                      out->flat<float>().data(), ws_ptr, ws_bytes,
                                      ctx->eigen_device<Eigen::GpuDevice>().stream());
    OP_REQUIRES(ctx, rc == CTCB200_OK, tf::errors::Internal("ctc_b200: ", ctcb200_strerror(rc)));
  }
};

REGISTER_KERNEL_BUILDER(Name("CtcB200LossGrad").Device(tf::DEVICE_GPU), CtcB200LossGradOp);
REGISTER_KERNEL_BUILDER(Name("CtcB200HvpLogits").Device(tf::DEVICE_GPU), CtcB200HvpLogitsOp);
