"""TensorFlow face of libctc_b200.so -- importable only where TensorFlow is installed (it is NOT in this image, so
this module is untested here; the torch-tensor face in the parent package is the one the test-suite exercises).

``classic_ctc_loss`` / ``simplified_ctc_loss`` keep the reference signatures (tf_seq2seq_losses/__init__.py:22-28) and
return a loss whose gradient w.r.t. ``logits`` is the fused kernel's output, wired with ``tf.custom_gradient`` like the
reference's ``forward_fn`` (base_loss.py:140-155).
"""
import os

try:
    import tensorflow as tf
except ImportError as exc:  # pragma: no cover
    raise ImportError("tf_seq2seq_losses_b200.tf_adapter needs TensorFlow; use the torch-tensor API of "
                      "tf_seq2seq_losses_b200 instead") from exc

_so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ctc_b200_tf_op.so")
_ops = tf.load_op_library(_so)   # build instructions: see ctc_b200_tf_op.cc


def _ctc_loss(labels, logits, label_length, logit_length, blank_index, variant):
    @tf.custom_gradient
    def fn(logits_):
        loss, grad = _ops.ctc_b200_loss_grad(labels=labels, logits=logits_, label_length=label_length,
                                             logit_length=logit_length, blank_index=int(blank_index), variant=variant)

        def backprop(d_loss):
            return tf.reshape(d_loss, [-1, 1, 1]) * grad

        return loss, backprop

    return fn(logits)


def classic_ctc_loss(labels, logits, label_length, logit_length, blank_index=0):
    return _ctc_loss(labels, logits, label_length, logit_length, blank_index, 0)


def simplified_ctc_loss(labels, logits, label_length, logit_length, blank_index=0):
    return _ctc_loss(labels, logits, label_length, logit_length, blank_index, 1)


simple_ctc_loss = simplified_ctc_loss
