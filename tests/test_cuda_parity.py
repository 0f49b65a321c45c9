"""Parity of the CUDA path (through the public Python face and the C ABI) against the CPU oracle, the reference's
known-answer vectors and size-independent properties.  Needs a B200: run with ``-m gpu``.

Tolerances (fp32 kernels vs the fp64 oracle; the reference's own fp32 arithmetic sits in the same band):
  loss      |d| <= 1e-5 * max(1, |loss|)
  gradient  max-abs <= 5e-5 for T <= 64 (the reference's 4-places bar), <= 5e-4 for T >= 500 (measured 1.6e-4 at
            T=1000,V=1024; reference-like float32 arithmetic gives 4.7e-3 there)
  Hessian   max-abs <= 5e-5 (measured 2.7e-5 at T=50), symmetry <= 1e-4
  +inf, exact zeros and the 1e10 / 100.0 known answers are compared bit-exactly.
"""
import numpy as np
import pytest
import torch

from oracle import ctc_oracle as orc
from tests.ref_cases import CLASSIC, KAT_CASES, README_EXAMPLE, README_GOLDEN, SIMPLIFIED, random_inputs

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_ATOL_SHORT = 5e-5
GRAD_ATOL_LONG = 5e-4
HESS_ATOL = 5e-5
HESS_SYM_ATOL = 1e-4


def _pkg():
    import tf_seq2seq_losses_b200 as pkg
    return pkg


@pytest.fixture(params=["fused", "staged"])
def kernel_path(request):
    """The loss+gradient call has two device paths: the fused single-launch kernel (default for V >= 64, or B >= 80, when its
    shared-memory plan fits) and the three staged kernels K1/K2/K3 (narrow vocabularies, oversized rows, and the
    states / Hessian entry points).  Both are forced in turn and must agree with the oracle on every shape."""
    from tf_seq2seq_losses_b200 import _lib
    old = _lib.DEFAULT_FLAGS
    _lib.DEFAULT_FLAGS = _lib.FORCE_STAGED if request.param == "staged" else _lib.FORCE_FUSED
    yield request.param
    _lib.DEFAULT_FLAGS = old


def _cls(variant):
    pkg = _pkg()
    return pkg.ClassicCtcLossData if variant == CLASSIC else pkg.SimplifiedCtcLossData


def _fn(variant):
    pkg = _pkg()
    return pkg.classic_ctc_loss if variant == CLASSIC else pkg.simplified_ctc_loss


def _cuda(a, dtype=None):
    t = torch.as_tensor(np.asarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _data_obj(case_or_inputs, variant, blank=0):
    logits, labels, ll, tl = case_or_inputs
    logprobas = torch.log_softmax(_cuda(logits, torch.float32), dim=2)
    return _cls(variant)(labels=_cuda(labels), logprobas=logprobas, label_length=_cuda(ll), logit_length=_cuda(tl),
                         blank_index=blank), logprobas


def _loss_close(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert np.array_equal(np.isinf(got), np.isinf(want)), (got, want)
    fin = ~np.isinf(want)
    assert np.all(np.abs(got[fin] - want[fin]) <= LOSS_RTOL * np.maximum(1.0, np.abs(want[fin]))), (got, want)


# ------------------------------------------------------------------------------------------------------------------
# the reference's literal known answers, through the data-class surface
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KAT_CASES, ids=[c["name"] for c in KAT_CASES])
def test_reference_known_answers(case):
    # the reference builds logprobas with logit_to_logproba(log(one-hot)) on the host side of the data class
    logprobas = orc.logit_to_logproba(case["logits"]).astype(np.float32)
    data = _cls(case["variant"])(labels=_cuda(case["labels"]), logprobas=_cuda(logprobas),
                                 label_length=_cuda(case["label_length"]), logit_length=_cuda(case["logit_length"]),
                                 blank_index=torch.tensor(case["blank"], dtype=torch.int32))
    loss = data.loss.cpu().numpy()
    if "exp_alpha" in case:
        assert np.array_equal(np.exp(data.alpha.cpu().numpy()), np.asarray(case["exp_alpha"], dtype=np.float32))
        assert np.array_equal(np.exp(data.beta.cpu().numpy()), np.asarray(case["exp_beta"], dtype=np.float32))
    if "loss" in case:
        if "loss_places" in case:
            assert np.max(np.abs(loss - np.asarray(case["loss"]))) < 0.5 * 10.0 ** -case["loss_places"]
        else:
            assert loss.tolist() == [float(v) for v in case["loss"]]
    if "loss_below" in case:
        assert float(loss[0]) < case["loss_below"]
    if "occupancy" in case:
        occ = np.exp(data.logarithmic_logproba_gradient.cpu().numpy())
        assert np.max(np.abs(occ - np.asarray(case["occupancy"]))) < 0.5e-6
    if "gradient" in case:
        g = data.gradient.cpu().numpy()
        if case.get("gradient_exact"):
            assert np.array_equal(g, np.asarray(case["gradient"], dtype=np.float32))
        else:
            assert np.max(np.abs(g - np.asarray(case["gradient"]))) < 0.5e-6
    if case.get("hessian_zero"):
        h = data.hessian.cpu().numpy()
        B, T, V = case["logits"].shape
        assert h.shape == (B, T, V, T, V)
        if case.get("gradient_exact"):
            assert np.array_equal(h, np.zeros_like(h))       # infeasible sample: exact zeros
        else:
            assert np.max(np.abs(h)) < 0.5e-6


def test_readme_example_loss_gradient_hessian():
    """BASELINE.json configs[0]: README example, loss + gradient + Hessian (w.r.t. logits, through double backward)."""
    c = README_EXAMPLE
    pkg = _pkg()
    for variant, fn in ((CLASSIC, pkg.classic_ctc_loss), (SIMPLIFIED, pkg.simple_ctc_loss)):
        logits = _cuda(c["logits"], torch.float32).requires_grad_(True)
        loss = fn(labels=_cuda(c["labels"]), logits=logits, label_length=_cuda(c["label_length"]),
                  logit_length=_cuda(c["logit_length"]), blank_index=0)
        (grad,) = torch.autograd.grad(loss.sum(), logits, create_graph=True)
        key = "classic" if variant == CLASSIC else "simplified"
        assert np.allclose(loss.detach().cpu().numpy(), README_GOLDEN[f"{key}_loss"], atol=2e-6)
        g = grad.detach().cpu().numpy()
        if variant == CLASSIC:
            assert np.allclose(g[0], 1 / 3 - np.eye(3)[[1, 2, 0, 2, 1]], atol=2e-6)
            assert np.allclose(g[1], README_GOLDEN["classic_grad_logits_1"], atol=2e-6)
        else:
            assert np.allclose(g[0], README_GOLDEN["simplified_grad_logits_0"], atol=2e-6)
            assert np.allclose(g[1], README_GOLDEN["simplified_grad_logits_1"], atol=2e-6)
        # batch_jacobian(gradient, logits) row by row
        B, T, V = 2, 5, 3
        H = np.zeros((B, T, V, T, V), dtype=np.float64)
        for t in range(T):
            for k in range(V):
                (row,) = torch.autograd.grad(grad[:, t, k].sum(), logits, retain_graph=True)
                H[:, t, k] = row.cpu().numpy()
        data, lp = orc.ctc_loss_data(c["labels"], c["logits"], c["label_length"], c["logit_length"], 0, variant)
        want = orc.hessian_logits(data, lp)
        assert np.max(np.abs(H - want)) < HESS_ATOL
        if variant == CLASSIC:
            assert abs(np.abs(H).sum() - README_GOLDEN["classic_hessian_logits_abs_sum"]) < 1e-4
            assert np.allclose(H[1, 0, :, 3, :], np.array([[1, -1, 0], [-1, 1, 0], [0, 0, 0]]) / 49, atol=2e-6)
            assert np.array_equal(H[1, 4], np.zeros((V, T, V))) and np.array_equal(H[1, :, :, 4], np.zeros((T, V, V)))


# ------------------------------------------------------------------------------------------------------------------
# random inputs vs the oracle
# ------------------------------------------------------------------------------------------------------------------
SHAPES = [
    # B, T, V, L, ragged, blank, labels_width
    (3, 6, 5, 3, True, 0, None),
    (8, 20, 8, 9, True, 0, 20),        # labels as wide as T, like the reference's generator (tests/common.py:89-94)
    (8, 64, 10, 30, True, 0, None),    # tests/test_classic_ctc_loss.py:360-393 shape
    (4, 33, 29, 12, True, 3, None),    # V % 4 != 0 -> scalar row path; non-zero blank
    (5, 40, 64, 40, False, 63, None),  # T == L: single feasible alignment for the simplified loss
    (3, 70, 128, 70, True, 0, None),   # U = 71 -> 3 states per lane
    (6, 37, 16, 5, True, 2, None),     # odd frame counts around the meet-in-the-middle split
    (4, 300, 256, 290, True, 0, None), # U = 291 -> 10 states per lane (one CTA per SM in the fused kernel)
]


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
@pytest.mark.parametrize("shape", SHAPES, ids=[f"B{s[0]}T{s[1]}V{s[2]}L{s[3]}" for s in SHAPES])
def test_loss_and_gradient_match_oracle(shape, variant, kernel_path):
    B, T, V, L, ragged, blank, lw = shape
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=B * 1000 + T, ragged=ragged, blank=blank, labels_width=lw)
    d_loss = np.linspace(0.5, 1.5, B).astype(np.float32)
    want_loss, want_grad, data = orc.loss_and_grad_logits(labels, logits, ll, tl, blank, variant, d_loss=d_loss)
    x = _cuda(logits).requires_grad_(True)
    loss = _fn(variant)(_cuda(labels), x, _cuda(ll), _cuda(tl), blank)
    fin = torch.isfinite(loss)
    (loss * _cuda(d_loss))[fin].sum().backward()
    _loss_close(loss.detach().cpu().numpy(), want_loss)
    got = x.grad.cpu().numpy()
    want_grad = np.where(np.isinf(want_loss)[:, None, None], 0.0, want_grad)
    grad_atol = GRAD_ATOL_SHORT if T <= 64 else GRAD_ATOL_LONG
    assert np.max(np.abs(got - want_grad)) <= grad_atol
    # frames beyond logit_length: exact zeros
    for b in range(B):
        assert np.array_equal(got[b, tl[b]:], np.zeros_like(got[b, tl[b]:]))
    # data-class surface on the same inputs
    obj, _ = _data_obj((logits, labels, ll, tl), variant, blank)
    assert np.max(np.abs(obj.gradient.cpu().numpy() - data.gradient)) <= grad_atol
    a, bt = obj.alpha.cpu().numpy(), obj.beta.cpu().numpy()
    assert a.shape == data.alpha.shape and bt.shape == data.beta.shape
    for got_s, want_s in ((a, data.alpha), (bt, data.beta)):
        assert np.array_equal(np.isinf(got_s), np.isinf(want_s))
        fin_s = ~np.isinf(want_s)
        assert np.max(np.abs(got_s[fin_s] - want_s[fin_s]) / np.maximum(1.0, np.abs(want_s[fin_s]))) < 1e-5


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_narrow_vocabulary_large_batch_takes_the_fused_kernel(variant):
    """Character-sized vocabularies are served by the staged kernels in small batches and by the fused kernel from 48
    utterances on (csrc/api.cu: fused_workers); both choices, unforced, against the oracle -- V = 29 rows are not 16-byte
    aligned (cp.async row mover), V = 32 rows are (TMA)."""
    import ctypes
    from tf_seq2seq_losses_b200 import _lib
    for (B, T, V, L, seed, want_path) in [(96, 40, 29, 10, 5, "kf_fused"), (96, 40, 32, 10, 6, "kf_fused"),
                                          (24, 40, 29, 10, 7, "k1_softmax_gather,k2_recursion,k3_grad")]:
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
        x, lab = _cuda(logits, torch.float32), _cuda(labels, torch.int32)
        desc = _lib.make_desc(x, lab, 0, variant, int(ll.max()) + 1, 0)
        assert _lib.load().ctcb200_stage_names(ctypes.byref(desc)).decode() == want_path
        loss, grad, _ = _lib.loss_grad(desc, x, lab, _cuda(ll, torch.int32), _cuda(tl, torch.int32))
        want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant)
        _loss_close(loss.cpu().numpy(), want_loss)
        want_grad[np.isinf(want_loss)] = 0.0
        assert np.max(np.abs(grad.cpu().numpy() - want_grad)) <= GRAD_ATOL_SHORT


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_time_major_logits(variant, kernel_path):
    """CTCB200_TIME_MAJOR / logits_time_major=True: logits and the gradient are [T,B,V]; same numbers as the batch-major
    call on the transposed array (rows are processed identically, only their addresses change), ragged lengths, an
    infeasible sample and an upstream gradient included."""
    from tf_seq2seq_losses_b200 import _lib
    for (B, T, V, L, seed) in [(5, 37, 96, 11, 8), (3, 20, 30, 6, 9)]:
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
        ll[0], tl[0] = L, L - 1                                  # infeasible: +inf, zero gradient rows
        w = _cuda(np.linspace(0.5, 2.0, B).astype(np.float32))
        outs = []
        for time_major in (False, True):
            x = _cuda(logits)
            if time_major:
                x = x.transpose(0, 1).contiguous()
            x.requires_grad_(True)
            loss = _fn(variant)(_cuda(labels), x, _cuda(ll), _cuda(tl), 0, logits_time_major=time_major)
            (torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)) * w).sum().backward()
            g = x.grad.transpose(0, 1) if time_major else x.grad
            outs.append((loss.detach().cpu().numpy(), g.contiguous().cpu().numpy()))
        assert np.array_equal(outs[0][0], outs[1][0]) and np.isinf(outs[0][0][0])
        assert np.max(np.abs(outs[0][1] - outs[1][1])) <= 1e-6
        assert np.array_equal(outs[1][1][0], np.zeros_like(outs[1][1][0]))
        want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant, d_loss=w.cpu().numpy())
        want_grad[np.isinf(want_loss)] = 0.0
        assert np.max(np.abs(outs[1][1] - want_grad)) <= 2 * GRAD_ATOL_SHORT
    # only the loss + gradient call takes the layout flag
    x = _cuda(logits, torch.float32)
    desc = _lib.make_desc(x.transpose(0, 1).contiguous(), _cuda(labels, torch.int32), 0, variant, L + 1, _lib.TIME_MAJOR)
    with pytest.raises(_lib.CtcB200Error):
        _lib.states(desc, x, _cuda(labels, torch.int32), _cuda(ll, torch.int32), _cuda(tl, torch.int32))


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_upstream_gradient_inside_the_kernels(variant, kernel_path):
    """ctcb200_loss_grad with a d_loss vector (forward_fn.backprop, base_loss.py:150-153, done inside the kernels: the
    scaled softmax pass and scatter of the fused kernel, the scale factor of K3), incl. negative, zero and unit weights."""
    from tf_seq2seq_losses_b200 import _lib
    for (B, T, V, L, seed) in [(6, 40, 70, 12, 3), (4, 33, 130, 9, 4)]:
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
        d_loss = np.array([1.0, -0.75, 0.0, 2.5, 1.0, 0.3][:B], dtype=np.float32)
        want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant, d_loss=d_loss)
        x, lab = _cuda(logits, torch.float32), _cuda(labels, torch.int32)
        desc = _lib.make_desc(x, lab, 0, variant, int(ll.max()) + 1, _lib.DEFAULT_FLAGS)
        loss, grad, _ = _lib.loss_grad(desc, x, lab, _cuda(ll, torch.int32), _cuda(tl, torch.int32), d_loss=_cuda(d_loss))
        _loss_close(loss.cpu().numpy(), want_loss)
        want_grad[np.isinf(want_loss)] = 0.0
        assert np.max(np.abs(grad.cpu().numpy() - want_grad)) <= 3 * GRAD_ATOL_SHORT      # |d_loss| up to 2.5


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_loss_only_call_equals_the_loss_of_the_full_call(variant, kernel_path):
    """forward_fn (base_loss.py:140-155) computes the loss alone; here that is ctcb200_loss_grad with both gradient pointers
    NULL (the fused kernel stops at the middle of the sequence).  Same losses as the full call, +inf included."""
    from tf_seq2seq_losses_b200 import _lib
    B, T, V, L = 7, 41, 96, 11
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=5)
    ll[2], tl[2] = 11, 6                      # infeasible
    tl[3] = 0
    ll[4] = 0
    x, lab, llc, tlc = _cuda(logits), _cuda(labels), _cuda(ll), _cuda(tl)
    desc = _lib.make_desc(x, lab, 0, variant, L + 1)
    only = _lib.loss_only(desc, x, lab, llc, tlc).cpu().numpy()
    full, _, _ = _lib.loss_grad(desc, x, lab, llc, tlc)
    want_loss, _, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant)
    _loss_close(only, want_loss)
    _loss_close(only, full.cpu().numpy())
    with torch.no_grad():                     # the public face under no_grad is the same call
        _loss_close(_fn(variant)(lab, x, llc, tlc, 0).cpu().numpy(), want_loss)


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_logarithmic_logproba_gradient_is_native_log_domain(variant):
    """base_loss.py:270-298: log occupancies are finite wherever a (frame, token) pair is possible -- also where the
    occupancy itself underflows float32 (log below -87 / -103) -- and -inf exactly elsewhere."""
    B, T, V, L = 3, 24, 12, 6
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=77, ragged=False)
    logits *= 30.0                            # very peaky frames: most alignments have weights far below e^-100
    tl[1] = 17
    ll[2], tl[2] = 6, 4                       # infeasible sample: all -inf
    data, lp = _data_obj((logits, labels, ll, tl), variant)
    ref = orc.CtcLossData(labels, lp.cpu().numpy().astype(np.float64), ll, tl, 0, variant)
    want = ref.logarithmic_logproba_gradient
    got = data.logarithmic_logproba_gradient.cpu().numpy().astype(np.float64)
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    fin = np.isfinite(want)
    assert (want[fin] < -110.0).any(), "the case must reach below float32's exp underflow"
    # float32 log-domain arithmetic on terms as large as the loss (~900 here, one ulp = 6e-5): a few ulps of those, plus
    # the usual relative bar on the value itself
    tol = 5e-7 * float(np.max(ref.loss[np.isfinite(ref.loss)])) + 2e-5 * np.maximum(1.0, np.abs(want[fin]))
    assert np.all(np.abs(got[fin] - want[fin]) <= tol), float(np.max(np.abs(got[fin] - want[fin])))
    # and exp(.) of it is the gradient the other entry point returns
    assert np.max(np.abs(np.exp(got) + data.gradient.cpu().numpy())) <= GRAD_ATOL_SHORT


@pytest.mark.parametrize("schedule", ["eager", "deferred"])
@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_cuda_graph_capture_and_replay(variant, schedule, monkeypatch):
    """The reference runs under tf.function / autograph (tests/test_simplified_ctc_loss.py:293-320,
    tests/test_hessian.py:213-257).  The equivalent here: loss + backward captured in a torch.cuda.CUDAGraph (no host
    synchronisation, no allocation outside the graph's pool) and replayed on new data, bit-identical to the un-captured
    call -- for both schedules of the training step (gradient computed in the forward pass and scaled in backward, or
    loss-only forward and the gradient computed in backward with d_loss inside the kernel; base_loss.py)."""
    from tf_seq2seq_losses_b200 import base_loss
    monkeypatch.setattr(base_loss, "EAGER_MAX_ELEMENTS", (1 << 24) if schedule == "eager" else 0)
    B, T, V, L = 6, 50, 128, 10
    fn = _fn(variant)
    logits0, labels, ll, tl = random_inputs(B, T, V, L, seed=41)
    logits1 = random_inputs(B, T, V, L, seed=42)[0]
    lab, llc, tlc = _cuda(labels), _cuda(ll), _cuda(tl)
    w = _cuda(np.linspace(0.5, 2.0, B).astype(np.float32))

    def eager(logits):
        x = _cuda(logits).requires_grad_(True)
        loss = fn(lab, x, llc, tlc, 0, max_label_length=L)
        (torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)) * w).sum().backward()
        return loss.detach().clone(), x.grad.clone()

    want0, want1 = eager(logits0), eager(logits1)
    static_x = _cuda(logits0).requires_grad_(True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                # warm-up on a side stream, as torch's capture recipe asks
        for _ in range(2):
            static_x.grad = None
            loss = fn(lab, static_x, llc, tlc, 0, max_label_length=L)
            (torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)) * w).sum().backward()
    torch.cuda.current_stream().wait_stream(side)
    static_x.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_loss = fn(lab, static_x, llc, tlc, 0, max_label_length=L)
        (torch.where(torch.isfinite(static_loss), static_loss, torch.zeros_like(static_loss)) * w).sum().backward()
    for logits, want in ((logits1, want1), (logits0, want0), (logits1, want1)):
        with torch.no_grad():
            static_x.copy_(_cuda(logits))
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(static_loss.detach(), want[0])
        assert torch.equal(static_x.grad, want[1])
    # device-resident lengths without max_label_length: no read-back during capture (U falls back to labels.shape[1] + 1)
    graph2 = torch.cuda.CUDAGraph()
    with torch.no_grad():
        fn(lab, static_x, llc, tlc, 0)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph2):
            loss2 = fn(lab, static_x, llc, tlc, 0)
        graph2.replay()
        torch.cuda.synchronize()
    _loss_close(loss2.cpu().numpy(), want1[0].cpu().numpy())


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
@pytest.mark.parametrize("shape", [(6, 41, 64, 9), (3, 70, 136, 70), (70, 24, 1024, 6), (80, 30, 8, 5)],
                         ids=lambda s: "B%dT%dV%dL%d" % s)
def test_bfloat16_logits(shape, variant):
    """CTCB200_LOGITS_BF16 (an extension; the reference asserts float32, base_loss.py:131): the kernels read bfloat16 rows,
    widen them on the fly and compute in fp32, so the result is the oracle's on the bf16-ROUNDED inputs -- loss and
    float32 gradient to the usual fp32 tolerance, the bfloat16 gradient to its roundings on top."""
    from tf_seq2seq_losses_b200 import _lib
    B, T, V, L = shape
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=B + T)
    logits *= 3.0
    if B >= 6:
        ll[1], tl[1] = L, max(L // 2 - 1, 0)       # infeasible
        tl[2] = T // 2 + 1                          # padded frames
    xb = _cuda(logits).to(torch.bfloat16)
    rounded = xb.float().cpu().numpy()
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, rounded, ll, tl, 0, variant)
    want_grad[np.isinf(want_loss)] = 0.0
    lab, llc, tlc = _cuda(labels), _cuda(ll), _cuda(tl)
    atol = GRAD_ATOL_SHORT if T <= 64 else GRAD_ATOL_LONG
    # C ABI: bf16 in, float32 gradient out
    desc = _lib.make_desc(xb, lab, 0, variant, L + 1)
    loss, grad, _ = _lib.loss_grad(desc, xb, lab, llc, tlc)
    assert grad.dtype == torch.float32
    _loss_close(loss.cpu().numpy(), want_loss)
    assert np.max(np.abs(grad.cpu().numpy() - want_grad)) <= atol
    _loss_close(_lib.loss_only(desc, xb, lab, llc, tlc).cpu().numpy(), want_loss)
    # public face: bf16 in, bf16 gradient out (autograd wants the input's dtype), upstream gradient applied in the kernel
    x = xb.clone().requires_grad_(True)
    w = np.linspace(0.5, 2.0, B).astype(np.float32)
    out = _fn(variant)(lab, x, llc, tlc, 0)
    (torch.where(torch.isfinite(out), out, torch.zeros_like(out)) * _cuda(w)).sum().backward()
    assert x.grad.dtype == torch.bfloat16
    want_w = want_grad * w[:, None, None]
    got = x.grad.float().cpu().numpy()
    # bfloat16 keeps 8 significant bits: one rounding is a relative 2^-8; the small-shape (eager) schedule rounds the unit
    # gradient in the kernel and the d_loss-weighted product once more
    assert np.max(np.abs(got - want_w) - np.abs(want_w) * 2.0 ** -7) <= 2.5 * atol
    # shapes the fused kernel cannot take are refused, not silently converted
    bad = _cuda(logits[:, :, :V - 4]).to(torch.bfloat16).contiguous()
    with pytest.raises(_lib.CtcB200Error):
        _lib.loss_grad(_lib.make_desc(bad, lab, 0, variant, L + 1), bad, lab, llc, tlc)


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_undefined_inputs_do_not_fault(variant, kernel_path):
    """Inputs the reference leaves undefined: label_length > labels.shape[1] (the reference pads with
    the blank as a *real* label), a real label equal to the blank, labels >= V or negative, logit_length > T, negative
    lengths.  The kernels must not fault and must not produce NaN; values are don't-care."""
    B, T, V, L = 4, 20, 8, 6
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=21, ragged=False, labels_width=4)
    ll[:] = [6, 5, 2, 3]
    labels[1, 1] = 0
    labels[2, 0] = V + 3
    labels[3, 1] = -2
    tl[:] = [T + 7, T, -3, T]
    x = _cuda(logits).requires_grad_(True)
    loss = _fn(variant)(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
    torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
    torch.cuda.synchronize()
    assert not torch.isnan(loss).any() and not torch.isnan(x.grad).any()


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_real_label_equal_to_blank_matches_reference_semantics(variant, kernel_path):
    """A real label equal to the blank is undefined input, but the reference's arithmetic is still definite: the label's
    log-probability feeds the recursion like any other (base_loss.py:328-344) and its emission occupancy is dropped by the
    blank-column override (classic_ctc_loss.py:647-654).  Both device paths follow the oracle's restatement of exactly
    that, so the same utterance cannot change its loss when a batch-size heuristic picks the other kernel."""
    B, T, V, L = 4, 14, 64, 6
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=33, ragged=False)
    labels[1, 2] = 0
    labels[2, 0] = 0
    labels[2, 1] = 0
    labels[3, 5] = 0
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant)
    assert np.isfinite(want_loss).all()
    x = _cuda(logits).requires_grad_(True)
    loss = _fn(variant)(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
    loss.sum().backward()
    _loss_close(loss.detach().cpu().numpy(), want_loss)
    assert np.max(np.abs(x.grad.cpu().numpy() - want_grad)) <= GRAD_ATOL_SHORT


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_edge_samples(variant, kernel_path):
    """Empty / infeasible / zero-length samples in one batch; -inf and 1e10 logits (BASELINE.json configs[4] edge set)."""
    B, T, V, L = 6, 12, 16, 5
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=11, ragged=False)
    logits[0, :, 1:] = -np.inf                     # only the blank is possible -> infeasible for a non-empty label
    for t in range(T):                             # sample 1: 1e10 on a valid alignment (label, then blanks)
        logits[1, t, labels[1, t] if t < L else 0] = 1e10
    labels[1, 1] = labels[1, 0] + 1 if labels[1, 0] + 1 < V else 1   # avoid a repeat so the path is valid for classic
    logits[1, 1, :] = np.random.default_rng(0).standard_normal(V)
    logits[1, 1, labels[1, 1]] = 1e10
    ll[2], tl[2] = 5, 3                            # label longer than the logits -> +inf, zero gradient
    ll[3] = 0                                      # empty label: loss = -sum log p(blank)
    tl[4] = 0                                      # no frames: +inf
    logits[5, 3, 2] = -np.inf                      # a single -inf entry inside a valid row
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant)
    x = _cuda(logits).requires_grad_(True)
    loss = _fn(variant)(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
    torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
    got_loss, got = loss.detach().cpu().numpy(), x.grad.cpu().numpy()
    assert np.isinf(got_loss[[0, 2, 4]]).all() and np.isinf(want_loss[[0, 2, 4]]).all()
    _loss_close(got_loss, want_loss)
    assert not np.isnan(got).any()
    for b in (0, 2, 4):
        assert np.array_equal(got[b], np.zeros_like(got[b]))
    for b in (1, 3, 5):
        assert np.max(np.abs(got[b] - want_grad[b])) <= GRAD_ATOL_SHORT


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_degenerate_shapes(variant, kernel_path):
    """tests/test_simplified_ctc_loss.py:322-366, tests/test_classic_ctc_loss.py:309-330: T = 0 and B = 0."""
    fn = _fn(variant)
    x = torch.zeros((1, 0, 3), device="cuda", requires_grad=True)
    loss = fn(_cuda([[1, 2]]), x, _cuda([2]), _cuda([2]), 0)
    assert loss.tolist() == [float("inf")]
    (g,) = torch.autograd.grad(loss.sum(), x, allow_unused=True)
    assert g is None or list(g.shape) == [1, 0, 3]
    x = torch.zeros((0, 4, 3), device="cuda", requires_grad=True)
    loss = fn(torch.zeros((0, 2), dtype=torch.int32).cuda(), x, torch.zeros((0,), dtype=torch.int32).cuda(),
              torch.zeros((0,), dtype=torch.int32).cuda(), 0)
    assert list(loss.shape) == [0]
    loss.sum().backward()
    assert list(x.grad.shape) == [0, 4, 3]


@pytest.mark.parametrize("k4", ["registers", "generic"])
@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_hessian_matches_oracle(variant, k4, monkeypatch):
    """ClassicCtcLossData.hessian / SimplifiedCtcLossData.hessian vs the oracle (literal for tiny, matrix-free for cfg-4 shape).
    The Hessian kernel has a register-resident form (U <= 128) and a generic shared-memory form; both are checked."""
    if k4 == "generic":
        monkeypatch.setenv("CTCB200_K4_GENERIC", "1")
    shapes = [(2, 4, 3, 2, 0), (2, 6, 5, 3, 1), (3, 50, 32, 15, 2)]
    if k4 == "registers":
        shapes.append((2, 34, 12, 33, 3))    # two states per lane, every token of the vocabulary several times in the label
    for (B, T, V, L, seed) in shapes:
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
        if T == 6:
            labels[0, :2] = 2           # repeated token
        data, lp = orc.ctc_loss_data(labels, logits, ll, tl, 0, variant)
        want = data.hessian if T <= 6 else data.hessian_fast()
        obj, logprobas = _data_obj((logits, labels, ll, tl), variant)
        got = obj.hessian.cpu().numpy()
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= HESS_ATOL
        # symmetry, tests/test_hessian.py:89-108
        assert np.max(np.abs(got - np.transpose(got, (0, 3, 4, 1, 2)))) <= HESS_SYM_ATOL
        # matrix-free contraction == dense contraction (gradient_fn.backprop, base_loss.py:167-173)
        v = np.random.default_rng(seed).standard_normal((B, T, V)).astype(np.float32)
        hv = obj.hessian_vector_product(_cuda(v)).cpu().numpy()
        assert np.max(np.abs(hv - np.einsum("btkuj,buj->btk", want, v))) <= 1e-3   # sums T*V terms of |v| ~ 1


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_hvp_wrt_logits_matches_oracle(variant):
    """ctcb200_hvp_logits: d_loss * (d2 loss / d logits2) v, matrix-free with the log-softmax chain of SURVEY.md appendix B
    (what tape.gradient of the first derivative returns in README.md:58-71), vs the oracle's dense Hessian w.r.t. logits."""
    from tf_seq2seq_losses_b200 import _lib
    for (B, T, V, L, seed) in [(2, 5, 3, 2, 0), (3, 9, 6, 4, 1), (2, 30, 70, 10, 2)]:
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
        data, lp = orc.ctc_loss_data(labels, logits, ll, tl, 0, variant)
        H = orc.hessian_logits(data, lp, data.hessian if T <= 9 else data.hessian_fast())
        rng = np.random.default_rng(seed)
        v = rng.standard_normal((B, T, V)).astype(np.float32)
        dl = rng.standard_normal((B,)).astype(np.float32)
        want = dl[:, None, None] * np.einsum("btkuj,buj->btk", H, v)
        x, lab = _cuda(logits, torch.float32), _cuda(labels, torch.int32)
        desc = _lib.make_desc(x, lab, 0, variant, int(ll.max()) + 1, 0)
        got = _lib.hvp_logits(desc, x, lab, _cuda(ll, torch.int32), _cuda(tl, torch.int32), _cuda(v), _cuda(dl))
        assert np.max(np.abs(got.cpu().numpy() - want)) <= 2e-4          # sums T*V terms of |v| ~ 1
        # rows at or beyond logit_length are exactly zero
        mask = np.arange(T)[None, :] >= tl[:, None]
        assert np.array_equal(got.cpu().numpy()[mask], np.zeros_like(want[mask]))


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_gamma_matches_oracle(variant):
    """ClassicCtcLossData.gamma / SimplifiedCtcLossData.gamma vs the literal oracle unfolding; gamma[:,0,0(,0)] == alpha
    (tests/test_hessian.py:62-87); -inf below the time diagonal, identity on it (classic_ctc_loss.py:204-213,286-308)."""
    for (B, T, V, L, seed) in [(2, 4, 3, 2, 0), (2, 6, 5, 3, 1), (3, 12, 7, 8, 2), (2, 40, 9, 36, 3)]:
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
        if T == 6:
            labels[0, :2] = 2           # repeated token
        data, _ = orc.ctc_loss_data(labels, logits, ll, tl, 0, variant)
        want = data.gamma
        obj, _ = _data_obj((logits, labels, ll, tl), variant)
        got = obj.gamma.cpu().numpy()
        assert got.shape == want.shape
        assert np.array_equal(np.isneginf(got), np.isneginf(want))
        fin = np.isfinite(want)
        assert np.max(np.abs(got[fin] - want[fin])) <= 1e-4
        first = got[:, 0, 0, 0] if variant == CLASSIC else got[:, 0, 0]
        alpha = obj.alpha.cpu().numpy()
        fin = np.isfinite(alpha)
        assert np.array_equal(np.isneginf(first), np.isneginf(alpha))
        assert np.max(np.abs(first[fin] - alpha[fin])) <= 1e-4


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_second_derivative_through_logproba_chain(variant):
    """ctc_loss_from_logproba differentiated twice (tests/test_hessian.py:110-147, test_classic_ctc_loss.py:443-477)."""
    pkg = _pkg()
    B, T, V, L = 2, 4, 3, 2
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=7)
    data, lp = orc.ctc_loss_data(labels, logits, ll, tl, 0, variant)
    logprobas = _cuda(lp, torch.float32).requires_grad_(True)
    loss = pkg.ctc_loss_from_logproba(_cuda(labels), logprobas, _cuda(ll), _cuda(tl), 0, _cls(variant))
    (g,) = torch.autograd.grad(loss.sum(), logprobas, create_graph=True)
    assert np.max(np.abs(g.detach().cpu().numpy() - data.gradient)) <= GRAD_ATOL_SHORT
    H = np.zeros((B, T, V, T, V))
    for t in range(T):
        for k in range(V):
            (row,) = torch.autograd.grad(g[:, t, k].sum(), logprobas, retain_graph=True)
            H[:, t, k] = row.cpu().numpy()
    assert np.max(np.abs(H - data.hessian)) <= HESS_ATOL
    with pytest.raises(NotImplementedError):
        lp2 = _cuda(lp, torch.float32).requires_grad_(True)
        loss2 = pkg.ctc_loss_from_logproba(_cuda(labels), lp2, _cuda(ll), _cuda(tl), 0, _cls(variant))
        (g2,) = torch.autograd.grad(loss2.sum(), lp2, create_graph=True)
        (h2,) = torch.autograd.grad(g2.sum(), lp2, create_graph=True)
        torch.autograd.grad(h2.sum(), lp2)


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs at full size: oracle on a subsample + size-independent properties
# ------------------------------------------------------------------------------------------------------------------
def _full_size_check(B, T, V, L, variant, ragged, n_check=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((B, T, V), generator=g, dtype=torch.float32)
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32)
    if ragged:
        tl = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32)
        ll = torch.randint(L // 2, L + 1, (B,), generator=g, dtype=torch.int32)
    else:
        tl = torch.full((B,), T, dtype=torch.int32)
        ll = torch.full((B,), L, dtype=torch.int32)
    x = logits.cuda().requires_grad_(True)
    loss = _fn(variant)(labels.cuda(), x, ll.cuda(), tl.cuda(), 0, max_label_length=L)
    loss.sum().backward()
    grad = x.grad
    # properties: every gradient row sums to zero (softmax - occupancy), padded rows are exactly zero
    assert float(grad.sum(dim=2).abs().max()) < 2e-3
    mask = torch.arange(T, device="cuda")[None, :] >= tl.cuda()[:, None]
    assert float(grad[mask].abs().max() if mask.any() else 0.0) == 0.0
    assert torch.isfinite(loss).all()
    idx = np.linspace(0, B - 1, n_check).astype(int)
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels.numpy()[idx], logits.numpy()[idx], ll.numpy()[idx],
                                                        tl.numpy()[idx], 0, variant)
    _loss_close(loss.detach().cpu().numpy()[idx], want_loss)
    err = float(np.max(np.abs(grad.cpu().numpy()[idx] - want_grad)))
    print(f"[full-size B={B} T={T} V={V} L={L} variant={variant} ragged={ragged}] grad max-abs err {err:.3e}")
    assert err <= (GRAD_ATOL_LONG if T >= 500 else GRAD_ATOL_SHORT)


def test_config1_character_asr_shape():
    """BASELINE.json configs[1]: classic_ctc_loss B=32 T=500 V=29 L=100."""
    _full_size_check(32, 500, 29, 100, CLASSIC, ragged=False)
    _full_size_check(32, 500, 29, 100, CLASSIC, ragged=True, seed=1)


@pytest.mark.parametrize("variant", [SIMPLIFIED, CLASSIC])
def test_config2_north_star_shape(variant, kernel_path):
    """BASELINE.json configs[2]: simple_ctc_loss B=256 T=1000 V=1024 L=200 (and the classic loss on the same shape)."""
    _full_size_check(256, 1000, 1024, 200, variant, ragged=False, n_check=3)
    _full_size_check(256, 1000, 1024, 200, variant, ragged=True, n_check=3, seed=1)


def test_config3_hessian_shape():
    """BASELINE.json configs[3]: ClassicCtcLossData.hessian B=64 T=50 V=32 L=15."""
    B, T, V, L = 64, 50, 32, 15
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=4)
    obj, _ = _data_obj((logits, labels, ll, tl), CLASSIC)
    got = obj.hessian
    assert list(got.shape) == [B, T, V, T, V]
    assert float((got - got.permute(0, 3, 4, 1, 2)).abs().max()) <= HESS_SYM_ATOL
    idx = [0, 31, 63]
    data, _ = orc.ctc_loss_data(labels[idx], logits[idx], ll[idx], tl[idx], 0, CLASSIC)
    assert np.max(np.abs(got[idx].cpu().numpy() - data.hessian_fast())) <= HESS_ATOL


def test_config4_large_vocab_slice(kernel_path):
    """BASELINE.json configs[4] (B=2048 T=1600 V=5000 L=400) on a one-GPU slice of the batch, incl. the edge samples."""
    B, T, V, L = 8, 1600, 5000, 400
    g = torch.Generator().manual_seed(5)
    logits = torch.randn((B, T, V), generator=g, dtype=torch.float32)
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32)
    tl = torch.full((B,), T, dtype=torch.int32)
    ll = torch.full((B,), L, dtype=torch.int32)
    logits[0, :, 1:] = -float("inf")
    ll[2], tl[2] = 400, 300
    ll[3] = 0
    tl[4] = 0
    x = logits.cuda().requires_grad_(True)
    loss = _pkg().classic_ctc_loss(labels.cuda(), x, ll.cuda(), tl.cuda(), 0)
    torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
    got_loss, grad = loss.detach().cpu().numpy(), x.grad.cpu().numpy()
    assert np.isinf(got_loss[[0, 2, 4]]).all() and not np.isnan(grad).any()
    idx = [1, 3, 5]
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels.numpy()[idx], logits.numpy()[idx], ll.numpy()[idx],
                                                        tl.numpy()[idx], 0, CLASSIC)
    _loss_close(got_loss[idx], want_loss)
    assert np.max(np.abs(grad[idx] - want_grad)) <= GRAD_ATOL_LONG
    for b in (0, 2, 4):
        assert np.array_equal(grad[b], np.zeros_like(grad[b]))


def test_config4_full_batch_on_one_gpu():
    """BASELINE.json configs[4] at its full size on ONE GPU: classic_ctc_loss B=2048 T=1600 V=5000 L=400 (65.5 GB of
    logits + 65.5 GB of gradient + 11 GB of workspace), logits generated on the device, with the edge samples of
    SURVEY.md 8(d) injected.  Checked through size-independent properties over the whole batch and against the oracle
    on a subsample."""
    import ctypes
    from tf_seq2seq_losses_b200 import _lib
    B, T, V, L = 2048, 1600, 5000, 400
    free, _ = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip(f"needs 150 GB of free device memory, have {free / 1e9:.0f} GB")
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(11)
    logits = torch.empty((B, T, V), dtype=torch.float32, device=dev)
    for b0 in range(0, B, 64):                       # generate in slices: randn's temporaries stay small
        logits[b0:b0 + 64].normal_(generator=gen)
    labels = torch.randint(1, V, (B, L), generator=gen, dtype=torch.int32, device=dev)
    tl = torch.full((B,), T, dtype=torch.int32, device=dev)
    ll = torch.full((B,), L, dtype=torch.int32, device=dev)
    tl[8:] = torch.randint(T // 2, T + 1, (B - 8,), generator=gen, dtype=torch.int32, device=dev)
    ll[8:] = torch.randint(L // 2, L + 1, (B - 8,), generator=gen, dtype=torch.int32, device=dev)
    logits[0, :, 1:] = -float("inf")                 # only the blank is possible: infeasible, +inf and zero gradient
    tt = torch.arange(T, device=dev)                 # sample 1: 1e10 on one valid alignment (3 frames per label, then a blank)
    logits[1, tt, torch.where(tt % 4 == 3, torch.zeros_like(tt), labels[1, tt // 4].long())] = 1e10
    ll[2], tl[2] = 400, 300                          # more labels than frames: +inf
    ll[3] = 0                                        # empty label: loss = -sum_t h[t]
    tl[4] = 0                                        # no frames: +inf (label_length > 0)
    desc = _lib.make_desc(logits, labels, 0, _lib.CLASSIC, L + 1, 0)
    lib = _lib.load()
    assert lib.ctcb200_stage_names(ctypes.byref(desc)).decode() == "kf_fused"
    n = lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_LOSS_GRAD_LOGITS)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    loss = torch.empty(B, device=dev)
    grad = torch.empty_like(logits)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.ctcb200_loss_grad(ctypes.byref(desc), P(logits), P(labels), P(ll), P(tl), None, P(loss), P(grad), None,
                                     P(ws), n, st))
    torch.cuda.synchronize()
    got_loss = loss.cpu().numpy()
    assert np.isinf(got_loss[[0, 2, 4]]).all() and np.isfinite(np.delete(got_loss, [0, 2, 4])).all()
    want_inf_rows = torch.tensor([0, 2, 4], device=dev)
    assert float(grad[want_inf_rows].abs().max()) == 0.0
    # whole-batch properties, in slices to bound temporaries: rows sum to zero, padded rows are exactly zero, no NaN
    for b0 in range(0, B, 128):
        g = grad[b0:b0 + 128]
        assert not torch.isnan(g).any()
        assert float(g.sum(dim=2).abs().max()) < 5e-3
        mask = torch.arange(T, device=dev)[None, :] >= tl[b0:b0 + 128, None]
        assert float(g[mask].abs().max() if mask.any() else 0.0) == 0.0
    # sample 1: the 1e10 path has probability one -> loss 0 and a zero gradient
    assert abs(float(got_loss[1])) < 1e-3 and float(grad[1].abs().max()) < 1e-6
    idx = [3, 5, 1024, 2047]
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels[idx].cpu().numpy(), logits[idx].cpu().numpy(),
                                                        ll[idx].cpu().numpy(), tl[idx].cpu().numpy(), 0, CLASSIC)
    _loss_close(got_loss[idx], want_loss)
    err = float(np.max(np.abs(grad[idx].cpu().numpy() - want_grad)))
    print(f"[config4 full batch] grad max-abs err {err:.3e}")
    assert err <= GRAD_ATOL_LONG


@pytest.mark.parametrize("merge_repeated", [True, False])
def test_greedy_decode_matches_oracle(merge_repeated):
    """ctcb200_greedy_decode (tf.nn.ctc_greedy_decoder in dense form) is integer work: bit-exact against the oracle, with
    ties (integer-valued logits: lowest index wins), ragged and zero lengths, a non-zero blank, unaligned rows, and the
    time-major layout."""
    pkg = _pkg()
    rng = np.random.default_rng(17)
    for (B, T, V, blank, ties) in [(5, 40, 29, 0, False), (4, 33, 64, 7, True), (3, 70, 1024, 1023, False), (2, 9, 4, 2, True)]:
        logits = rng.standard_normal((B, T, V)).astype(np.float32)
        if ties:
            logits = np.round(logits * 1.5).astype(np.float32)
        tl = rng.integers(0, T + 1, size=B).astype(np.int32)
        tl[0] = T
        want_dec, want_len, want_neg = orc.greedy_decode(logits, tl, blank, merge_repeated)
        for time_major in (False, True):
            x = _cuda(logits)
            if time_major:
                x = x.transpose(0, 1).contiguous()
            dec, length, neg = pkg.ctc_greedy_decode(x, _cuda(tl), blank, merge_repeated, time_major=time_major)
            assert np.array_equal(dec.cpu().numpy(), want_dec)
            assert np.array_equal(length.cpu().numpy(), want_len)
            assert np.allclose(neg.cpu().numpy(), want_neg, rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------------------------------
# boundary behaviour
# ------------------------------------------------------------------------------------------------------------------
def test_no_cpu_fallback_and_size_limits():
    from tf_seq2seq_losses_b200 import _lib
    pkg = _pkg()
    with pytest.raises(_lib.CtcB200Error):
        pkg.classic_ctc_loss(torch.ones((1, 2), dtype=torch.int32), torch.zeros((1, 4, 3)), torch.tensor([2]),
                             torch.tensor([4]), 0)
    with pytest.raises(_lib.CtcB200Error):   # 1100 label states > 1024
        pkg.classic_ctc_loss(torch.ones((1, 1100), dtype=torch.int32).cuda(), torch.zeros((1, 1200, 3)).cuda(),
                             torch.tensor([1099]).cuda(), torch.tensor([1200]).cuda(), 0)
    with pytest.raises(AssertionError):      # base_loss.py:129-138
        pkg.classic_ctc_loss(torch.ones((2, 2), dtype=torch.int32).cuda(), torch.zeros((1, 4, 3)).cuda(),
                             torch.tensor([2]).cuda(), torch.tensor([4]).cuda(), 0)


def test_host_buffer_entry_point_matches_device_entry_point():
    """ctcb200_host_loss_grad (pinned host buffers, sliced copies) == ctcb200_loss_grad on resident tensors."""
    from tf_seq2seq_losses_b200 import _lib
    B, T, V, L = 10, 40, 64, 12
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=9)
    ctx = _lib.HostContext(B, T, V, L, 0, SIMPLIFIED, L + 1, device=0, num_slices=3)
    pin = lambda a: torch.as_tensor(a).pin_memory()
    loss_h = torch.empty((B,), dtype=torch.float32).pin_memory()
    grad_h = torch.empty((B, T, V), dtype=torch.float32).pin_memory()
    ctx.loss_grad(pin(logits), pin(labels), pin(ll), pin(tl), loss_h, grad_h)
    desc = _lib.make_desc(_cuda(logits), _cuda(labels), 0, SIMPLIFIED, L + 1)
    loss_d, grad_d, _ = _lib.loss_grad(desc, _cuda(logits), _cuda(labels), _cuda(ll), _cuda(tl))
    torch.cuda.synchronize()
    assert torch.equal(loss_h, loss_d.cpu()) and torch.equal(grad_h, grad_d.cpu())
    ctx.close()


def test_host_buffer_entry_point_uneven_tail_of_a_narrow_vocabulary():
    """The training-call workspace is not monotonic in the batch size: with V < 64 a full slice of 49 utterances takes the
    fused kernel (small scratch) while the 47-utterance tail takes the staged kernels (three times larger).  The host
    context sizes its workspace for both."""
    from tf_seq2seq_losses_b200 import _lib
    B, T, V, L = 145, 30, 32, 8
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=10)
    ctx = _lib.HostContext(B, T, V, L, 0, CLASSIC, L + 1, device=0, num_slices=3)      # slices of 49, 49, 47
    lib, d = _lib.load(), lambda b: _lib.Desc(b, T, V, L, 0, CLASSIC, L + 1, 0)
    import ctypes
    assert lib.ctcb200_workspace_bytes(ctypes.byref(d(47)), _lib.WS_LOSS_GRAD_LOGITS) > \
        lib.ctcb200_workspace_bytes(ctypes.byref(d(49)), _lib.WS_LOSS_GRAD_LOGITS)
    pin = lambda a: torch.as_tensor(a).pin_memory()
    loss_h = torch.empty((B,), dtype=torch.float32).pin_memory()
    grad_h = torch.empty((B, T, V), dtype=torch.float32).pin_memory()
    ctx.loss_grad(pin(logits), pin(labels), pin(ll), pin(tl), loss_h, grad_h)
    ctx.close()
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, CLASSIC)
    _loss_close(loss_h.numpy(), want_loss)
    want_grad[np.isinf(want_loss)] = 0.0
    assert np.max(np.abs(grad_h.numpy() - want_grad)) <= GRAD_ATOL_SHORT


# ------------------------------------------------------------------------------------------------------------------
# randomized shape sweep: every frame count / label length / worker configuration corner of the fused kernel
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_random_shape_sweep(variant, kernel_path):
    """Small random problems with frame counts 0..9 around the meet-in-the-middle split, empty labels, infeasible
    samples, V not a multiple of 4, non-zero blank -- compared with the C restatement (pinned against the numpy oracle)."""
    from oracle import c_oracle
    rng = np.random.default_rng(1234 + variant)
    for trial in range(24):
        B = int(rng.integers(1, 6))
        T = int(rng.integers(1, 14))
        V = int(rng.choice([3, 4, 7, 8, 33, 64, 100, 260]))
        Lw = int(rng.integers(1, 8))
        blank = int(rng.integers(0, V))
        logits = (rng.standard_normal((B, T, V)) * rng.choice([0.1, 1.0, 5.0])).astype(np.float32)
        labels = rng.integers(0, V - 1, size=(B, Lw)).astype(np.int32)
        labels = np.where(labels >= blank, labels + 1, labels).astype(np.int32)
        ll = rng.integers(0, Lw + 1, size=B).astype(np.int32)
        tl = rng.integers(0, T + 1, size=B).astype(np.int32)
        want_loss, want_grad = c_oracle.loss_grad(labels, logits, ll, tl, blank, variant)
        x = _cuda(logits).requires_grad_(True)
        loss = _fn(variant)(_cuda(labels), x, _cuda(ll), _cuda(tl), blank)
        torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
        _loss_close(loss.detach().cpu().numpy(), want_loss)
        got = x.grad.cpu().numpy()
        assert not np.isnan(got).any()
        assert np.max(np.abs(got - want_grad)) <= GRAD_ATOL_SHORT, (trial, B, T, V, Lw, blank, ll, tl)


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_midsize_random_sweep(variant, kernel_path):
    """Mid-sized random problems (up to 120 frames, 100 labels = 4 states per lane, character to BPE-sized vocabularies,
    many repeated labels, ragged and infeasible lengths, peaky logits, every blank position) against the C restatement;
    the fixed-seed slice of tools/fuzz.py."""
    from oracle import c_oracle
    rng = np.random.default_rng(4321 + variant)
    for trial in range(40):
        B = int(rng.integers(1, 9))
        T = int(rng.integers(1, 121))
        V = int(rng.choice([3, 5, 8, 29, 32, 33, 64, 100, 128, 260, 1024]))
        Lw = int(rng.integers(1, 101))
        blank = int(rng.integers(0, V))
        logits = (rng.standard_normal((B, T, V)) * rng.choice([0.1, 1.0, 4.0])).astype(np.float32)
        labels = rng.integers(0, V - 1, size=(B, Lw)).astype(np.int32)
        labels = np.where(labels >= blank, labels + 1, labels).astype(np.int32)
        ll = rng.integers(0, min(Lw, T) + 1, size=B).astype(np.int32)
        if rng.random() < 0.2:
            ll[rng.integers(0, B)] = Lw                          # possibly more labels than frames
        tl = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
        want_loss, want_grad = c_oracle.loss_grad(labels, logits, ll, tl, blank, variant)
        want_grad[np.isinf(want_loss)] = 0.0
        x = _cuda(logits).requires_grad_(True)
        loss = _fn(variant)(_cuda(labels), x, _cuda(ll), _cuda(tl), blank)
        torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
        _loss_close(loss.detach().cpu().numpy(), want_loss)
        got = x.grad.cpu().numpy()
        assert not np.isnan(got).any()
        # peaky logits and near-infeasible alignments reach 1.4e-4 in fp32 at T ~ 100 (800 fuzz trials): long-sequence bar
        assert np.max(np.abs(got - want_grad)) <= GRAD_ATOL_LONG, (trial, B, T, V, Lw, blank, ll, tl)


@pytest.mark.parametrize("variant", [SIMPLIFIED, CLASSIC])
@pytest.mark.parametrize("cfg", [(1, 2, 0, 2, 0), (1, 2, 0, 1, 0), (2, 2, 0, 4, 0), (2, 2, 0, 2, 0), (2, 3, 0, 3, 0), (3, 2, 1, 6, 0),
                                 (3, 3, 0, 4, 0), (4, 2, 0, 8, 0), (4, 2, 1, 8, 0), (4, 2, 0, 6, 0), (4, 2, 1, 5, 0), (4, 2, 0, 4, 0),
                                 (8, 3, 1, 16, 1), (8, 2, 0, 16, 1), (8, 2, 1, 9, 1), (6, 3, 1, 12, 1), (5, 2, 0, 5, 1), (4, 3, 1, 8, 1),
                                 (2, 2, 0, 3, 1), (1, 2, 0, 1, 1), (4, 2, 1, 8, 2), (2, 2, 0, 4, 2), (8, 3, 1, 16, 3), (2, 2, 0, 2, 1)],
                         ids=lambda c: "W%d_SL%d_XA%d_R%d_mode%d" % c if isinstance(c, tuple) else str(c))
def test_fused_worker_configurations(cfg, variant):
    """Every (workers per side, row buffers, extra phase-A buffer, ring depth, mode) plan of the fused kernel gives the same
    answer -- including the shortest rings (depth = workers per side, and a single slot), for which the exchange vectors
    of the middle may or may not alias the input rings, the split plans (mode bit 0: a two-CTA cluster per utterance) and
    both state-scratch schemes (every second row, the default wherever workers and ring depth are even and the variant is
    the simplified one; every row with mode bit 1)."""
    from tf_seq2seq_losses_b200 import _lib
    lib = _lib.load()
    lib.ctcb200_debug_fused_plan(*cfg)
    old = _lib.DEFAULT_FLAGS
    _lib.DEFAULT_FLAGS = _lib.FORCE_FUSED
    fn = _pkg().simple_ctc_loss if variant == SIMPLIFIED else _pkg().classic_ctc_loss
    try:
        # (the last shape has more utterances than the GPU has SMs: a second wave of CTAs)
        for (B, T, V, L, seed) in [(5, 61, 96, 20, 0), (3, 30, 37, 9, 1), (2, 7, 64, 70, 2), (3, 45, 64, 14, 3), (170, 14, 64, 5, 4)]:
            if (cfg[4] & 1) and V % 4:
                continue                 # the split plans need TMA-movable rows (V % 4 == 0)
            logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
            if L == 70:
                ll[:] = [2, 70]          # 71 label states (3 per lane) over 7 frames: one feasible, one infeasible sample
            if seed == 3:
                labels[:, 5:9] = labels[:, 1:5]      # repeated tokens: several states scatter into one gradient column
                labels[0, 9] = labels[0, 1]
            want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant)
            x = _cuda(logits).requires_grad_(True)
            loss = fn(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
            torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
            _loss_close(loss.detach().cpu().numpy(), want_loss)
            want_grad[np.isinf(want_loss)] = 0.0
            assert np.max(np.abs(x.grad.cpu().numpy() - want_grad)) <= GRAD_ATOL_SHORT
    finally:
        _lib.DEFAULT_FLAGS = old
        lib.ctcb200_debug_fused_plan(0, 0, 0, 0, 0)


@pytest.mark.parametrize("variant", [SIMPLIFIED, CLASSIC])
@pytest.mark.parametrize("V", [2048, 2052, 3000, 5000])
@pytest.mark.parametrize("cfg", [(0, 0, 0, 0, 0), (2, 2, 0, 4, 0), (1, 2, 0, 2, 0), (2, 2, 0, 2, 0), (0, 0, 0, 0, 8)],
                         ids=["default", "W2", "W1", "W2_R2", "no_helpers"])
def test_row_helpers_on_wide_rows(cfg, V, variant):
    """Wide fp32 rows with at most two workers per side: every worker has a helper warp that reduces / exponentiates the upper
    part of its rows (kf_fused.cuh, helper_phase).  Shapes whose helper part is exactly one chunk, one chunk and one
    element, and several chunks with a ragged tail; ragged lengths, an infeasible sample, repeated tokens and an upstream
    gradient (the scaled softmax pass); with mode bit 3 the same shapes run without helpers."""
    from oracle import c_oracle
    from tf_seq2seq_losses_b200 import _lib
    lib = _lib.load()
    lib.ctcb200_debug_fused_plan(*cfg)
    old = _lib.DEFAULT_FLAGS
    _lib.DEFAULT_FLAGS = _lib.FORCE_FUSED
    fn = _pkg().simple_ctc_loss if variant == SIMPLIFIED else _pkg().classic_ctc_loss
    try:
        for (B, T, L, seed) in [(3, 41, 12, 5), (2, 9, 40, 6), (150, 6, 3, 7)]:
            logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed + V)
            if L == 40:
                ll[:] = [4, 40]           # 40 labels over at most 9 frames: infeasible
            labels[0, 1] = labels[0, 0]   # a repeated token
            want_loss, want_grad = c_oracle.loss_grad(labels, logits, ll, tl, 0, variant)
            want_grad[np.isinf(want_loss)] = 0.0
            weights = np.linspace(0.5, 2.0, B).astype(np.float32)
            for w in (None, weights):
                x = _cuda(logits).requires_grad_(True)
                loss = fn(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
                fin = torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss))
                (fin.sum() if w is None else (fin * _cuda(w)).sum()).backward()
                _loss_close(loss.detach().cpu().numpy(), want_loss)
                want = want_grad if w is None else want_grad * w[:, None, None]
                got = x.grad.cpu().numpy()
                assert not np.isnan(got).any()
                assert np.max(np.abs(got - want)) <= GRAD_ATOL_SHORT * (1.0 if w is None else 2.0)
            # the loss-only call stops at the middle: helpers take part in phase A alone
            only = fn(_cuda(labels), _cuda(logits), _cuda(ll), _cuda(tl), 0)
            _loss_close(only.cpu().numpy(), want_loss)
    finally:
        _lib.DEFAULT_FLAGS = old
        lib.ctcb200_debug_fused_plan(0, 0, 0, 0, 0)


def test_size_limits_both_paths():
    """The largest label-state count the fused kernel carries (U = 512 -> 16 states per lane) on both device paths, the
    staged kernels beyond it (U = 640 -> 20 per lane, U = 1024 -> 32 per lane, both variants), and a vocabulary whose
    rows do not fit the fused kernel's shared-memory plan (V = 32768 -> staged kernels by construction)."""
    import ctypes
    from oracle import c_oracle
    from tf_seq2seq_losses_b200 import _lib
    B, T, V, L = 2, 560, 64, 511
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=17, ragged=False)
    ll[1], tl[1] = 300, 420
    want_loss, want_grad = c_oracle.loss_grad(labels, logits, ll, tl, 0, CLASSIC)
    old = _lib.DEFAULT_FLAGS
    try:
        for flags in (_lib.FORCE_FUSED, _lib.FORCE_STAGED):
            _lib.DEFAULT_FLAGS = flags
            x = _cuda(logits).requires_grad_(True)
            loss = _pkg().classic_ctc_loss(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
            loss.sum().backward()
            _loss_close(loss.detach().cpu().numpy(), want_loss)
            assert np.max(np.abs(x.grad.cpu().numpy() - want_grad)) <= GRAD_ATOL_LONG
    finally:
        _lib.DEFAULT_FLAGS = old
    for (T, L, variant, seed) in ((700, 639, SIMPLIFIED, 19), (1100, 1023, CLASSIC, 20), (1100, 1023, SIMPLIFIED, 21)):
        B, V = 2, 48
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed, ragged=False)
        ll[1], tl[1] = L - 300, T - 200
        want_loss, want_grad = c_oracle.loss_grad(labels, logits, ll, tl, 0, variant)
        # Labels nearly as long as the utterance: the states on the only feasible ridge sit hundreds of nats below the
        # most likely prefix, where an fp32 ulp is 3e-5 per step.  The bar is the reference's own arithmetic: the fp32
        # restatement of the reference (no renormalisation, magnitudes ~4000) is off by 2e-3 .. 5e-3 on these inputs.
        _, f32_grad = c_oracle.loss_grad(labels, logits, ll, tl, 0, variant, dtype=np.float32)
        tol = max(GRAD_ATOL_LONG, float(np.max(np.abs(f32_grad - want_grad))))
        wide = _lib.Desc(B, T, V, L, 0, variant, L + 1, 0)
        assert _lib.load().ctcb200_stage_names(ctypes.byref(wide)).decode().startswith("k1_")
        x = _cuda(logits).requires_grad_(True)
        fn = _pkg().classic_ctc_loss if variant == CLASSIC else _pkg().simple_ctc_loss
        loss = fn(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
        loss.sum().backward()
        _loss_close(loss.detach().cpu().numpy(), want_loss)
        err = float(np.max(np.abs(x.grad.cpu().numpy() - want_grad)))
        print(f"wide states T={T} L={L} variant={variant}: grad err {err:.2e} (fp32 restatement of the reference: {tol:.2e})")
        assert err <= tol
    B, T, V, L = 2, 40, 32768, 10
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=18)
    want_loss, want_grad = c_oracle.loss_grad(labels, logits, ll, tl, 0, SIMPLIFIED)
    x = _cuda(logits).requires_grad_(True)
    loss = _pkg().simple_ctc_loss(_cuda(labels), x, _cuda(ll), _cuda(tl), 0)
    loss.sum().backward()
    _loss_close(loss.detach().cpu().numpy(), want_loss)
    assert np.max(np.abs(x.grad.cpu().numpy() - want_grad)) <= GRAD_ATOL_SHORT
