"""Pins the CPU oracle (oracle/ctc_oracle.py) against the reference's own known-answer tests, against
torch.nn.functional.ctc_loss (independent native implementation of the classic loss) and against torch autograd
derivatives of a differentiable restatement.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ctc_oracle as orc
from tests.ref_cases import CLASSIC, KAT_CASES, README_EXAMPLE, README_GOLDEN, SIMPLIFIED, random_inputs


def _data(case, variant=None, dtype=np.float64):
    v = case["variant"] if variant is None else variant
    return orc.ctc_loss_data(case["labels"], case["logits"], case["label_length"], case["logit_length"],
                             case["blank"], v, dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", KAT_CASES, ids=[c["name"] for c in KAT_CASES])
def test_reference_known_answers(case, dtype):
    data, _ = _data(case, dtype=dtype)
    if "exp_alpha" in case:
        assert np.array_equal(np.exp(data.alpha), np.asarray(case["exp_alpha"], dtype=dtype))
        assert np.array_equal(np.exp(data.beta), np.asarray(case["exp_beta"], dtype=dtype))
    if "loss" in case:
        if "loss_places" in case:
            assert np.max(np.abs(data.loss - np.asarray(case["loss"]))) < 0.5 * 10.0 ** -case["loss_places"]
        else:
            assert data.loss.tolist() == list(case["loss"])
    if "loss_below" in case:
        assert float(data.loss[0]) < case["loss_below"]
    if "occupancy" in case:
        assert np.max(np.abs(np.exp(data.logarithmic_logproba_gradient) - np.asarray(case["occupancy"]))) < 0.5e-6
    if "gradient" in case:
        if case.get("gradient_exact"):
            assert np.array_equal(data.gradient, np.asarray(case["gradient"], dtype=dtype))
        else:
            assert np.max(np.abs(data.gradient - np.asarray(case["gradient"]))) < 0.5e-6
    if case.get("hessian_zero"):
        assert np.max(np.abs(data.hessian)) < 0.5e-6
        assert np.max(np.abs(data.hessian_fast())) < 0.5e-6


def test_tools_logsumexp_kat():
    """tests/test_tools.py:37-51."""
    x = np.array([-3.0753517, -np.inf, -np.inf], dtype=np.float32)
    y = np.array([-1.0e12, -4.283799e-01, -np.inf], dtype=np.float32)
    out = orc.logsumexp2(x, y)
    assert abs(out[0] - -3.0753517) < 1e-6 and abs(out[1] - -0.4283799) < 1e-6 and out[2] == -np.inf


def test_tools_unsorted_segment_logsumexp_kat():
    """tests/test_tools.py:137-148: an all -inf segment stays -inf (no NaN)."""
    data = np.array([0, -np.inf, 0, -np.inf], dtype=np.float32)
    out = orc.unsorted_segment_logsumexp(data, np.array([0, 1, 0, 1]), 2)
    assert abs(out[0] - np.log(2)) < 1e-7 and out[1] == -np.inf


def test_readme_example_goldens():
    """README.md:50-56 / tests/test_hessian.py:185-211 with the closed forms of SURVEY.md 8(c)."""
    c = README_EXAMPLE
    loss, grad, data = orc.loss_and_grad_logits(c["labels"], c["logits"], c["label_length"], c["logit_length"], 0, CLASSIC)
    assert np.allclose(loss, README_GOLDEN["classic_loss"], atol=1e-12)
    onehot = np.eye(3)[[1, 2, 0, 2, 1]]
    assert np.allclose(grad[0], 1 / 3 - onehot, atol=1e-12)
    assert np.allclose(grad[1], README_GOLDEN["classic_grad_logits_1"], atol=1e-12)
    assert np.allclose(data.gradient[1], README_GOLDEN["classic_gradient_1"], atol=1e-12)
    loss_s, grad_s, _ = orc.loss_and_grad_logits(c["labels"], c["logits"], c["label_length"], c["logit_length"], 0, SIMPLIFIED)
    assert np.allclose(loss_s, README_GOLDEN["simplified_loss"], atol=1e-12)
    assert np.allclose(grad_s[0], README_GOLDEN["simplified_grad_logits_0"], atol=1e-12)
    assert np.allclose(grad_s[1], README_GOLDEN["simplified_grad_logits_1"], atol=1e-12)
    # Hessian w.r.t. logits
    logprobas = orc.logit_to_logproba(c["logits"])
    H = orc.hessian_logits(data, logprobas, data.hessian)
    assert abs(np.abs(H).sum() - README_GOLDEN["classic_hessian_logits_abs_sum"]) < 1e-6
    for t in range(5):
        assert np.allclose(H[0, t, :, t, :], np.eye(3) / 3 - 1 / 9, atol=1e-12)
        for t2 in range(5):
            if t2 != t:
                assert np.allclose(H[0, t, :, t2, :], 0, atol=1e-12)
    assert np.allclose(H[1, 0, :, 3, :], np.array([[1, -1, 0], [-1, 1, 0], [0, 0, 0]]) / 49, atol=1e-12)
    assert np.allclose(H[1, 4], 0) and np.allclose(H[1, :, :, 4], 0)


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_alpha_beta_identity(variant):
    """tests/test_classic_ctc_loss.py:146-167, tests/test_simplified_ctc_loss.py:185-206."""
    logits, labels, ll, tl = random_inputs(3, 6, 5, 3, seed=1)
    data, _ = orc.ctc_loss_data(labels, logits, ll, tl, 0, variant)
    axes = (2, 3) if variant == CLASSIC else 2
    sums = orc.reduce_logsumexp(data.alpha + data.beta, axis=axes)
    assert np.max(np.abs(sums + data.loss[:, None])) < 1e-9


def test_classic_matches_torch_ctc_loss():
    """tests/test_classic_ctc_loss.py:332-393 with torch's native CTC standing in for tf.nn.ctc_loss."""
    for (B, T, V, L, seed) in [(8, 20, 8, 9, 0), (8, 64, 10, 30, 1)]:
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed)
        loss, grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, CLASSIC)
        x = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
        lp = torch.log_softmax(x, dim=2).transpose(0, 1)
        ref = torch.nn.functional.ctc_loss(lp, torch.tensor(labels, dtype=torch.long), torch.tensor(tl, dtype=torch.long),
                                           torch.tensor(ll, dtype=torch.long), blank=0, reduction="none", zero_infinity=False)
        ref.sum().backward()
        assert np.max(np.abs(ref.detach().numpy() - loss)) < 1e-9
        assert np.max(np.abs(x.grad.numpy() - grad)) < 1e-9


# ---- differentiable torch restatement (fp64) for first / second derivatives ------------------------------
def _lse(terms):
    """logsumexp over the reachable terms only (None = unreachable state), so autograd never sees -inf."""
    terms = [t for t in terms if t is not None]
    return torch.logsumexp(torch.stack(terms), 0) if terms else None


def _add(x, y):
    return None if (x is None or y is None) else x + y


def _torch_loss(lp, labels, label_length, logit_length, blank, variant):
    """Sum over the batch of the loss as a differentiable function of *logprobas* lp [B,T,V] (torch fp64)."""
    B, T, V = lp.shape
    total = 0.0
    for b in range(B):
        L, n_t = int(label_length[b]), int(min(logit_length[b], T))
        lab = [int(v) for v in labels[b, :L]]
        if variant == SIMPLIFIED:
            a = [torch.tensor(0.0, dtype=lp.dtype)] + [None] * L
            for t in range(n_t):
                a = [_lse([_add(lp[b, t, blank], a[l]), _add(lp[b, t, lab[l - 1]], a[l - 1]) if l > 0 else None])
                     for l in range(L + 1)]
            total = total - a[L]
        else:
            closed = [torch.tensor(0.0, dtype=lp.dtype)] + [None] * L
            opened = [None] * (L + 1)
            for t in range(n_t):
                nc, no = [], []
                for l in range(L + 1):
                    nc.append(_add(lp[b, t, blank], _lse([closed[l], opened[l]])))
                    terms = []
                    if l > 0:
                        e = lp[b, t, lab[l - 1]]
                        terms = [_add(e, opened[l]), _add(e, closed[l - 1])]
                        if l == 1 or lab[l - 1] != lab[l - 2]:
                            terms.append(_add(e, opened[l - 1]))
                    no.append(_lse(terms))
                closed, opened = nc, no
            total = total - _lse([closed[L], opened[L]])
    return total


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_gradient_and_hessian_match_autograd(variant):
    """Property pins of tests/test_classic_ctc_loss.py:395-425,479-514 and tests/test_hessian.py:149-183, with exact
    autograd derivatives in place of finite differences; also pins hessian (literal, gamma-based) == hessian_fast."""
    logits, labels, ll, tl = random_inputs(2, 5, 4, 3, seed=3)
    labels[0, :3] = [2, 2, 1]          # force a repeat
    ll[0], tl[0] = 3, 5
    logprobas = orc.logit_to_logproba(logits.astype(np.float64))
    data = orc.CtcLossData(labels, logprobas, ll, tl, 0, variant)
    lp = torch.tensor(logprobas, dtype=torch.float64, requires_grad=True)
    f = lambda z: _torch_loss(z, labels, ll, tl, 0, variant)
    (g,) = torch.autograd.grad(f(lp), lp)
    assert np.max(np.abs(g.numpy() - data.gradient)) < 1e-10
    H = torch.autograd.functional.hessian(f, lp).numpy()          # [B,T,V,B,T,V]
    for b in range(2):
        assert np.max(np.abs(H[b, :, :, b] - data.hessian[b])) < 1e-10
    assert np.max(np.abs(data.hessian - data.hessian_fast())) < 1e-10
    # symmetry, tests/test_hessian.py:89-108
    assert np.max(np.abs(data.hessian - np.transpose(data.hessian, (0, 3, 4, 1, 2)))) < 1e-12
    # second derivative w.r.t. logits through the log-softmax chain
    x = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    Hx = torch.autograd.functional.hessian(lambda z: f(torch.log_softmax(z, dim=2)), x).numpy()
    Hl = orc.hessian_logits(data, logprobas)
    for b in range(2):
        assert np.max(np.abs(Hx[b, :, :, b] - Hl[b])) < 1e-10


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_gamma_first_slice_is_alpha(variant):
    """tests/test_hessian.py:62-87: gamma[:, 0, 0(,0)] == alpha."""
    logits, labels, ll, tl = random_inputs(2, 4, 3, 2, seed=5)
    data, _ = orc.ctc_loss_data(labels, logits, ll, tl, 0, variant)
    first = data.gamma[:, 0, 0, 0] if variant == CLASSIC else data.gamma[:, 0, 0]
    assert np.array_equal(np.exp(first), np.exp(data.alpha))


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
def test_degenerate_shapes(variant):
    """tests/test_simplified_ctc_loss.py:322-366, tests/test_classic_ctc_loss.py:309-330."""
    loss, grad, _ = orc.loss_and_grad_logits(np.array([[1, 2]]), np.zeros((1, 0, 3)), np.array([2]), np.array([2]), 0, variant)
    assert loss.tolist() == [np.inf] and grad.shape == (1, 0, 3)
    loss, grad, _ = orc.loss_and_grad_logits(np.zeros((0, 2), int), np.zeros((0, 4, 3)), np.zeros((0,), int),
                                             np.zeros((0,), int), 0, variant)
    assert loss.shape == (0,) and grad.shape == (0, 4, 3)
    # label_length == 0: loss = -sum_t h[t]
    loss, grad, _ = orc.loss_and_grad_logits(np.array([[1, 2]]), np.zeros((1, 3, 3)), np.array([0]), np.array([3]), 0, variant)
    assert abs(loss[0] - 3 * np.log(3)) < 1e-12


def test_greedy_decode_known_answers():
    """Hand-checked best-path decodings (collapse repeats, then drop blanks -- "a_bb_ccc_c -> abcc", README.md:31 of the
    reference describes the same collapsing rule for the classic loss)."""
    path = [1, 0, 2, 2, 0, 3, 3, 3, 0, 3]                       # a _ b b _ c c c _ c  with blank 0
    logits = np.full((1, len(path), 4), -1.0, dtype=np.float32)
    logits[0, np.arange(len(path)), path] = 2.0
    dec, length, neg = orc.greedy_decode(logits, [len(path)], 0, True)
    assert dec[0, :4].tolist() == [1, 2, 3, 3] and length.tolist() == [4] and (dec[0, 4:] == -1).all()
    assert abs(neg[0] + 2.0 * len(path)) < 1e-12
    dec, length, _ = orc.greedy_decode(logits, [len(path)], 0, False)
    assert dec[0, :7].tolist() == [1, 2, 2, 3, 3, 3, 3] and length.tolist() == [7]
    dec, length, _ = orc.greedy_decode(logits, [3], 0, True)     # only the first three frames count
    assert dec[0, :2].tolist() == [1, 2] and length.tolist() == [2]
    dec, length, _ = orc.greedy_decode(np.zeros((1, 3, 4), dtype=np.float32), [3], 0, True)   # ties: index 0 = blank
    assert length.tolist() == [0]

