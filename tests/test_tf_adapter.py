"""The TensorFlow custom-op adapter, where TensorFlow exists (it does not in the development image: the whole module is
skipped there).  Builds tf_adapter/ctc_b200_tf_op.so against the installed TensorFlow, then checks the README example
(README.md:50-71 of the reference): loss, first derivative and the second-order tape against the oracle."""
import os
import subprocess

import numpy as np
import pytest

tf = pytest.importorskip("tensorflow")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_readme_example_through_the_tensorflow_op():
    subprocess.run(["bash", os.path.join(ROOT, "tf_seq2seq_losses_b200", "tf_adapter", "build.sh")], check=True)
    from oracle import ctc_oracle as orc
    from tests.ref_cases import README_EXAMPLE
    from tf_seq2seq_losses_b200.tf_adapter import classic_ctc_loss
    c = README_EXAMPLE
    logits, labels, ll, tl = c["logits"], c["labels"], c["label_length"], c["logit_length"]
    x = tf.constant(np.asarray(logits, np.float32))
    with tf.GradientTape() as t2:
        t2.watch(x)
        with tf.GradientTape() as t1:
            t1.watch(x)
            loss = classic_ctc_loss(tf.constant(labels), x, tf.constant(ll), tf.constant(tl), 0)
        grad = t1.gradient(tf.reduce_sum(loss), x)
        probe = tf.reduce_sum(grad * grad)
    second = t2.gradient(probe, x)
    want_loss, want_grad, data = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, orc.CLASSIC)
    assert np.allclose(loss.numpy(), want_loss, rtol=1e-5)
    assert np.max(np.abs(grad.numpy() - want_grad)) < 5e-5
    lp = orc.logit_to_logproba(np.asarray(logits, np.float64))
    hess = orc.hessian_logits(data, lp)
    want_second = 2.0 * np.einsum("btkuv,buv->btk", hess, want_grad)
    assert np.max(np.abs(second.numpy() - want_second)) < 5e-4
