"""Parity against outputs of the REFERENCE ITSELF.

tests/golden/reference/reference_outputs.npz holds what /root/reference/tf_seq2seq_losses computes -- its own Python, imported
unmodified, with TensorFlow served by the numpy shim tests/golden/tf_numpy_shim.py -- for the seeded inputs of
tests/golden/make_reference_golden.py: the data classes' loss / gradient / logarithmic_logproba_gradient / alpha / beta /
hessian / gamma (base_loss.py:186-298, classic_ctc_loss.py, simplified_ctc_loss.py) and the loss of the public functions
classic_ctc_loss / simplified_ctc_loss, in double precision.

  * CPU: the oracle (numpy and C restatements) reproduces every stored array to 1e-12; where /root/reference exists the
    generator is re-run and must reproduce the committed file bit for bit.
  * GPU: the CUDA path, through the package's reference-shaped Python face, matches the stored arrays within the fp32
    tolerances of tests/test_cuda_parity.py.
Nothing here reads /root/reference except the one CPU test that is skipped when it is absent."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ctc_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "reference", "reference_outputs.npz")
CASE_NAMES = ["small_ragged", "repeats", "blank_mid", "blank_last_empty_label", "labels_wider_than_needed", "mid", "long"]
# a real label equal to the blank -- undefined input whose arithmetic in the reference is nevertheless definite: pinned on the
# CPU here; the CUDA path's agreement with the oracle on it is tests/test_cuda_parity.py::
# test_real_label_equal_to_blank_matches_reference_semantics
CPU_ONLY_CASES = ["label_equals_blank"]
VARIANTS = [("classic", orc.CLASSIC), ("simplified", orc.SIMPLIFIED)]
FIRST_ORDER = ["loss", "gradient", "logarithmic_logproba_gradient", "alpha", "beta"]


def _load(name):
    d = np.load(FIXTURE)
    inputs = (d[f"{name}/logits"], d[f"{name}/labels"], d[f"{name}/label_length"], d[f"{name}/logit_length"],
              int(d[f"{name}/blank"]))

    def ref(tag, key):
        k = f"{name}/f64/{tag}/{key}"
        return d[k] if k in d.files else None
    return inputs, ref


def _same(got, want, atol):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    fin = np.isfinite(want)
    assert np.array_equal(got[~fin], want[~fin])            # +inf / -inf exactly where the reference has them
    if fin.any():
        assert np.max(np.abs(got[fin] - want[fin])) <= atol


def _grad_logits_from(gradient_logproba, logits):
    """d loss / d logits from the reference's d loss / d logproba: TensorFlow's autodiff of logit_to_logproba
    (tools.py:27-40), g - softmax * sum_k g."""
    x = logits.astype(np.float64)
    soft = np.exp(x - x.max(axis=2, keepdims=True))
    soft /= soft.sum(axis=2, keepdims=True)
    return gradient_logproba - soft * gradient_logproba.sum(axis=2, keepdims=True)


@pytest.mark.parametrize("tag,variant", VARIANTS)
@pytest.mark.parametrize("name", CASE_NAMES + CPU_ONLY_CASES)
def test_oracle_reproduces_the_reference_outputs(name, tag, variant):
    (logits, labels, ll, tl, blank), ref = _load(name)
    logprobas = orc.logit_to_logproba(logits.astype(np.float64))
    data = orc.CtcLossData(labels, logprobas, ll, tl, blank, variant)
    for key in FIRST_ORDER + ["hessian", "gamma"]:
        want = ref(tag, key)
        if want is not None:
            _same(getattr(data, key), want, 1e-12)
    if ref(tag, "hessian") is not None:
        _same(data.hessian_fast(), ref(tag, "hessian"), 1e-10)      # the matrix-free form the kernels use
    # the public functions (logits in): numpy oracle and its C restatement
    loss, grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, blank, variant)
    _same(loss, ref(tag, "public_loss"), 1e-12)
    want_grad = _grad_logits_from(ref(tag, "gradient"), logits)
    _same(grad, want_grad, 1e-12)
    from oracle import c_oracle
    closs, cgrad = c_oracle.loss_grad(labels, logits, ll, tl, blank, variant)
    _same(closs, ref(tag, "public_loss"), 1e-10)
    _same(cgrad, want_grad, 1e-10)


@pytest.mark.skipif(not os.path.isdir("/root/reference/tf_seq2seq_losses"), reason="the reference sources are not on this machine")
def test_fixture_is_what_the_reference_computes_here():
    """Re-runs the reference under the shim (in a subprocess: the shim registers itself as `tensorflow`) and compares
    every array with the committed fixture, bit for bit."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_reference_golden.py"), "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/tf_seq2seq_losses"), reason="the reference sources are not on this machine")
def test_oracle_equals_the_reference_on_random_problems():
    """200 random small problems (every blank position, ragged / infeasible / empty lengths): loss, gradient, logarithmic
    gradient, alpha, beta, Hessian and gamma of the oracle against the reference's own code, at 1e-12."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_reference_golden.py"), "--fuzz", "200"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="the reference sources are not on this machine")
def test_the_shim_passes_the_reference_own_unit_tests():
    """The reference's own test modules, unmodified, against its own code on the numpy shim: every test that does not need
    TensorFlow's autodiff / tf.nn.ctc_loss passes -- the literal alpha / beta tables, the exact 0 / 100.0 / 1e10 / +inf
    losses, the exact gradients, the tools.py examples.  This is what licenses the shim as a stand-in for TensorFlow here."""
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "run_reference_tests.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    summary = r.stdout.strip().splitlines()[-1]
    n_pass = int(summary.split(":")[1].split("passed")[0])
    assert n_pass >= 28 and " 0 failed" in summary, summary


@pytest.mark.gpu
@pytest.mark.parametrize("tag,variant", VARIANTS)
@pytest.mark.parametrize("name", CASE_NAMES)
def test_cuda_path_matches_the_reference_outputs(name, tag, variant):
    import torch
    import tf_seq2seq_losses_b200 as pkg
    (logits, labels, ll, tl, blank), ref = _load(name)
    cuda = lambda a: torch.as_tensor(a).cuda()      # noqa: E731
    # ---- the data-class surface, fed like the reference feeds it: float32 log-probabilities ----
    logprobas = torch.log_softmax(cuda(logits), dim=2)
    cls = pkg.ClassicCtcLossData if variant == orc.CLASSIC else pkg.SimplifiedCtcLossData
    data = cls(labels=cuda(labels), logprobas=logprobas, label_length=cuda(ll), logit_length=cuda(tl), blank_index=blank)
    want_loss = ref(tag, "loss")
    got_loss = data.loss.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.isinf(got_loss), np.isinf(want_loss))
    fin = np.isfinite(want_loss)
    assert np.all(np.abs(got_loss[fin] - want_loss[fin]) <= 1e-5 * np.maximum(1.0, np.abs(want_loss[fin])))
    # gradient / Hessian w.r.t. log-probabilities: absolute tolerances of tests/test_cuda_parity.py (5e-4 beyond 64 frames)
    grad_atol = 5e-5 if logits.shape[1] <= 64 else 5e-4
    for key, atol in (("gradient", grad_atol), ("hessian", 5e-5)):
        want = ref(tag, key)
        if want is not None:
            got = getattr(data, key).cpu().numpy().astype(np.float64)
            assert got.shape == want.shape
            assert np.max(np.abs(got - want)) <= atol, key
    # alpha / beta: -inf exactly where the reference has it, finite values to 1e-5 relative
    for key in ("alpha", "beta"):
        want = ref(tag, key)
        if want is None:
            continue
        got = getattr(data, key).cpu().numpy().astype(np.float64)
        assert got.shape == want.shape
        assert np.array_equal(np.isinf(got), np.isinf(want)), key
        both = ~np.isinf(want)
        assert np.max(np.abs(got[both] - want[both]) / np.maximum(1.0, np.abs(want[both]))) < 1e-5, key
    # log-domain gradient: -inf where the reference's is, finite values to the bar of the native log-domain test
    want = ref(tag, "logarithmic_logproba_gradient")
    got = data.logarithmic_logproba_gradient.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    both = np.isfinite(want)
    if both.any():
        tol = 5e-7 * float(np.max(want_loss[fin])) + 2e-5 * np.maximum(1.0, np.abs(want[both]))
        assert np.all(np.abs(got[both] - want[both]) <= tol)
    # ---- the public functions: loss and d loss / d logits through autograd ----
    fn = pkg.classic_ctc_loss if variant == orc.CLASSIC else pkg.simple_ctc_loss
    x = cuda(logits).requires_grad_(True)
    loss = fn(cuda(labels), x, cuda(ll), cuda(tl), blank)
    torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
    got_loss = loss.detach().cpu().numpy().astype(np.float64)
    want_loss = ref(tag, "public_loss")
    assert np.array_equal(np.isinf(got_loss), np.isinf(want_loss))
    assert np.all(np.abs(got_loss[fin] - want_loss[fin]) <= 1e-5 * np.maximum(1.0, np.abs(want_loss[fin])))
    want_grad = _grad_logits_from(ref(tag, "gradient"), logits)
    want_grad[np.isinf(want_loss)] = 0.0
    assert np.max(np.abs(x.grad.cpu().numpy() - want_grad)) <= grad_atol
