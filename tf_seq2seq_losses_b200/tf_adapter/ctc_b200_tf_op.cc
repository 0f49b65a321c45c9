// TensorFlow custom-op adapter over libctc_b200.so (include/ctc_b200.h).
//
// NOT BUILT IN THIS IMAGE: TensorFlow (headers and runtime) is not installed here and cannot be fetched, so this file
// is compiled and tested only where `import tensorflow` works:
//   g++ -std=c++17 -shared -fPIC ctc_b200_tf_op.cc -o ctc_b200_tf_op.so \
//       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()+tf.sysconfig.get_link_flags()))') \
//       -I../../include -L.. -lctc_b200 -Wl,-rpath,'$ORIGIN/..'
// All logic lives behind the C ABI; this file only adapts TensorFlow's buffer/stream ownership to it.  The Python side
// (tf_adapter/__init__.py) wraps the op in the same three-level tf.custom_gradient the reference uses
// (tf_seq2seq_losses/base_loss.py:140-184).
#include "ctc_b200.h"
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#define EIGEN_USE_GPU
#include "tensorflow/core/util/gpu_kernel_helper.h"

namespace tf = tensorflow;

REGISTER_OP("CtcB200LossGrad")
    .Input("labels: int32")
    .Input("logits: float32")
    .Input("label_length: int32")
    .Input("logit_length: int32")
    .Attr("blank_index: int = 0")
    .Attr("variant: int = 0")             // 0 classic, 1 simplified
    .Attr("max_label_length_plus_one: int = 0")
    .Output("loss: float32")
    .Output("grad_logits: float32")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->Vector(c->Dim(c->input(1), 0)));
      c->set_output(1, c->input(1));
      return tf::OkStatus();
    });

class CtcB200LossGradOp : public tf::OpKernel {
 public:
  explicit CtcB200LossGradOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("blank_index", &blank_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("variant", &variant_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("max_label_length_plus_one", &u_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& labels = ctx->input(0);
    const tf::Tensor& logits = ctx->input(1);
    const tf::Tensor& label_length = ctx->input(2);
    const tf::Tensor& logit_length = ctx->input(3);
    // tf_seq2seq_losses/base_loss.py:129-138
    OP_REQUIRES(ctx, logits.dims() == 3 && labels.dims() == 2 && label_length.dims() == 1 && logit_length.dims() == 1,
                tf::errors::InvalidArgument("rank mismatch"));
    OP_REQUIRES(ctx, logits.dim_size(0) == labels.dim_size(0) && logits.dim_size(0) == label_length.dim_size(0) &&
                         logits.dim_size(0) == logit_length.dim_size(0),
                tf::errors::InvalidArgument("batch mismatch"));
    ctcb200_desc d{};
    d.B = static_cast<int32_t>(logits.dim_size(0));
    d.T = static_cast<int32_t>(logits.dim_size(1));
    d.V = static_cast<int32_t>(logits.dim_size(2));
    d.Lw = static_cast<int32_t>(labels.dim_size(1));
    d.blank = blank_; d.variant = variant_; d.U = u_; d.flags = 0;
    tf::Tensor* loss = nullptr;
    tf::Tensor* grad = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({d.B}), &loss));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, logits.shape(), &grad));
    const size_t ws_bytes = ctcb200_workspace_bytes(&d, CTCB200_WS_LOSS_GRAD_LOGITS);
    OP_REQUIRES(ctx, ws_bytes > 0 || d.B == 0, tf::errors::InvalidArgument("ctc_b200: unsupported shape"));
    tf::Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<tf::int64>(ws_bytes + 256)}), &ws));
    auto base = reinterpret_cast<uintptr_t>(ws.flat<tf::uint8>().data());
    void* ws_ptr = reinterpret_cast<void*>((base + 255) & ~uintptr_t(255));
    auto stream = ctx->eigen_device<Eigen::GpuDevice>().stream();
    const int rc = ctcb200_loss_grad(&d, logits.flat<float>().data(), labels.flat<tf::int32>().data(),
                                     label_length.flat<tf::int32>().data(), logit_length.flat<tf::int32>().data(),
                                     /*d_loss=*/nullptr, loss->flat<float>().data(), grad->flat<float>().data(),
                                     /*grad_logprobas=*/nullptr, ws_ptr, ws_bytes, stream);
    OP_REQUIRES(ctx, rc == CTCB200_OK, tf::errors::Internal("ctc_b200: ", ctcb200_strerror(rc)));
  }

 private:
  int blank_ = 0, variant_ = 0, u_ = 0;
};

REGISTER_KERNEL_BUILDER(Name("CtcB200LossGrad").Device(tf::DEVICE_GPU), CtcB200LossGradOp);
