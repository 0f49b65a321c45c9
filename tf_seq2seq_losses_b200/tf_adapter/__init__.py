"""TensorFlow face of libctc_b200.so -- EXPERIMENTAL: importable only where TensorFlow is installed.  It is NOT in the
image this repository was developed in, so this module has never been executed there (tests/test_tf_adapter.py runs it
where TensorFlow exists); the torch-tensor face in the parent package is the one the test-suite exercises.

``classic_ctc_loss`` / ``simplified_ctc_loss`` keep the reference signatures (tf_seq2seq_losses/__init__.py:22-28) and
the reference's three nested ``tf.custom_gradient`` levels (tf_seq2seq_losses/base_loss.py:140-184):
  forward_fn   loss alone (a loss-only library call)                                   base_loss.py:140-155
  gradient_fn  d_loss * d loss / d logits, d_loss applied inside the kernel            base_loss.py:157-175
  hessian_fn   second-order backprop: d_loss * (d2 loss / d logits2) v, matrix-free    base_loss.py:167-173,177-184
               (ctcb200_hvp_logits); differentiating it once more raises NotImplementedError like base_loss.py:179-182.
Unlike the reference the chain is written w.r.t. the *logits* (the log-softmax of tools.py:27-40 is fused into the kernels).
"""
import os

try:
    import tensorflow as tf
except ImportError as exc:  # pragma: no cover
    raise ImportError("tf_seq2seq_losses_b200.tf_adapter needs TensorFlow; use the torch-tensor API of "
                      "tf_seq2seq_losses_b200 instead") from exc

_so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ctc_b200_tf_op.so")
_ops = tf.load_op_library(_so)   # built by tf_adapter/build.sh


def _max_label_length_plus_one(labels, label_length, max_label_length):
    """base_loss.py:478-486.  Any bound >= max(label_length) + 1 gives the same loss; inside a tf.function the static
    bound labels.shape[1] + 1 is used unless ``max_label_length`` is passed (the reference's generator makes labels as wide
    as the logits, tests/common.py:89-94, and the kernels carry at most 512 label states)."""
    if max_label_length is not None:
        return int(max_label_length) + 1
    if tf.executing_eagerly() and int(tf.size(label_length)) > 0:
        return max(int(tf.reduce_max(label_length)), 0) + 1
    return int(labels.shape[1]) + 1


def _ctc_loss(labels, logits, label_length, logit_length, blank_index, variant, max_label_length=None):
    labels, label_length, logit_length = (tf.cast(t, tf.int32) for t in (labels, label_length, logit_length))
    kw = dict(blank_index=int(blank_index), variant=variant,
              max_label_length_plus_one=_max_label_length_plus_one(labels, label_length, max_label_length))
    common = lambda x, d_loss: dict(labels=labels, logits=x, label_length=label_length, logit_length=logit_length,
                                    d_loss=d_loss, **kw)

    @tf.custom_gradient
    def hessian_fn(x, d_loss, v):
        out = _ops.ctc_b200_hvp_logits(v=v, **common(x, d_loss))

        def backprop(_):
            raise NotImplementedError("Third order derivative over the ctc loss function is not implemented.")

        return out, backprop

    @tf.custom_gradient
    def gradient_fn(x, d_loss):
        _, grad = _ops.ctc_b200_loss_grad(with_gradient=True, **common(x, d_loss))

        def backprop(v):
            _, unit = _ops.ctc_b200_loss_grad(with_gradient=True, **common(x, tf.ones_like(d_loss)))
            return hessian_fn(x, d_loss, v), tf.reduce_sum(v * unit, axis=[1, 2])

        return grad, backprop

    @tf.custom_gradient
    def forward_fn(x):
        loss, _ = _ops.ctc_b200_loss_grad(with_gradient=False, **common(x, tf.ones([tf.shape(x)[0]], tf.float32)))
        return loss, lambda d_loss: gradient_fn(x, d_loss)

    return forward_fn(logits)


def classic_ctc_loss(labels, logits, label_length, logit_length, blank_index=0, max_label_length=None):
    return _ctc_loss(labels, logits, label_length, logit_length, blank_index, 0, max_label_length)


def simplified_ctc_loss(labels, logits, label_length, logit_length, blank_index=0, max_label_length=None):
    return _ctc_loss(labels, logits, label_length, logit_length, blank_index, 1, max_label_length)


simple_ctc_loss = simplified_ctc_loss
