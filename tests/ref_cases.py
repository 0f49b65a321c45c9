"""The reference test-suite's literal known-answer cases, re-expressed as data.

Each case cites the reference test it restates (paths under /root/reference).  The same table drives the
oracle pinning tests (CPU) and the CUDA parity tests (GPU), so both are held to the reference's own vectors.
``one_hot_log`` rows are ``log`` of the listed probabilities (``-inf`` where 0), as the reference builds them
with ``tf.math.log(tf.constant(...))`` followed by ``logit_to_logproba``.
"""
from __future__ import annotations

import numpy as np

CLASSIC, SIMPLIFIED = 0, 1
LN = np.log


def _log(p):
    with np.errstate(divide="ignore"):
        return np.log(np.asarray(p, dtype=np.float64))


def _case(**kw):
    kw.setdefault("blank", 0)
    kw["logits"] = np.asarray(kw["logits"], dtype=np.float64)
    kw["labels"] = np.asarray(kw["labels"], dtype=np.int32)
    kw["label_length"] = np.asarray(kw["label_length"], dtype=np.int32)
    kw["logit_length"] = np.asarray(kw["logit_length"], dtype=np.int32)
    return kw


# expect keys: loss (exact unless loss_places), exp_alpha / exp_beta (exact), occupancy = exp(lg) (places 6),
# gradient (wrt logprobas; exact if gradient_exact else places 6), hessian_zero
KAT_CASES = [
    _case(name="classic_single_logit", ref="tests/test_classic_ctc_loss.py:33-65", variant=CLASSIC,
          logits=_log([[[0, 1, 0]]]), labels=[[1]], label_length=[1], logit_length=[1],
          exp_alpha=[[[[1, 0], [0, 0]], [[0, 0], [0, 1]]]],
          exp_beta=[[[[1, 1], [0, 1]], [[0, 0], [1, 1]]]],
          loss=[0.0], occupancy=[[[0.0, 1.0, 0.0]]]),
    _case(name="classic_closed_state", ref="tests/test_classic_ctc_loss.py:67-105", variant=CLASSIC,
          logits=_log([[[0, 1, 0], [1, 0, 0]]]), labels=[[1]], label_length=[1], logit_length=[2],
          exp_alpha=[[[[1, 0], [0, 0]], [[0, 0], [0, 1]], [[0, 0], [1, 0]]]],
          exp_beta=[[[[1, 1], [0, 1]], [[0, 0], [1, 1]], [[0, 0], [1, 1]]]],
          loss=[0.0], occupancy=[[[0.0, 1.0, 0.0], [1.0, 0.0, 0.0]]]),
    _case(name="classic_simple_case", ref="tests/test_classic_ctc_loss.py:107-144", variant=CLASSIC,
          logits=_log([[[0, 1, 0], [0, 0, 1], [1, 0, 0], [0, 0, 1], [0, 1, 0]]]),
          labels=[[1, 2, 2, 1]], label_length=[4], logit_length=[5],
          loss_below=1e-6,
          occupancy=[[[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0]]]),
    _case(name="classic_length_two", ref="tests/test_classic_ctc_loss.py:169-199", variant=CLASSIC,
          logits=np.zeros((2, 2, 3)), labels=[[1, 2], [1, 2]], label_length=[2, 1], logit_length=[2, 2],
          loss=[LN(9.0), LN(3.0)], loss_places=6,
          gradient=[[[0.0, -1.0, 0.0], [0.0, 0.0, -1.0]], [[-1 / 3, -2 / 3, 0.0], [-1 / 3, -2 / 3, 0.0]]]),
    _case(name="classic_too_short_logit", ref="tests/test_classic_ctc_loss.py:201-241", variant=CLASSIC,
          logits=np.zeros((1, 2, 3)), labels=[[1, 1]], label_length=[2], logit_length=[2],
          loss=[np.inf], gradient=np.zeros((1, 2, 3)), gradient_exact=True, hessian_zero=True),
    _case(name="classic_repeated_token", ref="tests/test_classic_ctc_loss.py:243-262", variant=CLASSIC,
          logits=np.zeros((1, 3, 3)), labels=[[1, 1]], label_length=[2], logit_length=[3],
          loss=[LN(27.0)], loss_places=6),
    _case(name="classic_single_token", ref="tests/test_classic_ctc_loss.py:264-283", variant=CLASSIC,
          logits=np.zeros((1, 3, 3)), labels=[[1]], label_length=[1], logit_length=[3],
          loss=[LN(27.0 / 6.0)], loss_places=6),
    _case(name="classic_wrong_prediction", ref="tests/test_classic_ctc_loss.py:285-307", variant=CLASSIC,
          logits=[[[0.0, 0.0, 100.0]]], labels=[[1]], label_length=[1], logit_length=[1],
          loss=[100.0], gradient=[[[0.0, -1.0, 0.0]]], gradient_exact=True),
    _case(name="simplified_simple_case", ref="tests/test_simplified_ctc_loss.py:35-91", variant=SIMPLIFIED,
          logits=_log([[[0, 1, 0], [1, 0, 0], [0, 0, 1], [1, 0, 0], [0, 1, 0]]]),
          labels=[[1, 2, 1]], label_length=[3], logit_length=[5],
          exp_alpha=[[[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]]],
          exp_beta=[[[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]]],
          loss_below=1e-6),
    _case(name="simplified_non_zero_blank", ref="tests/test_simplified_ctc_loss.py:93-115", variant=SIMPLIFIED,
          logits=_log([[[1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0]]]),
          labels=[[0, 2, 0]], label_length=[3], logit_length=[5], blank=1, loss_below=1e-6),
    _case(name="simplified_shorter_lengths", ref="tests/test_simplified_ctc_loss.py:117-138", variant=SIMPLIFIED,
          logits=_log([[[1, 0, 0], [0, 1, 0], [1, 0, 0], [1, 0, 0]]]),
          labels=[[1, 0]], label_length=[1], logit_length=[3], loss=[0.0]),
    _case(name="simplified_label_longer_than_logit", ref="tests/test_simplified_ctc_loss.py:140-160",
          variant=SIMPLIFIED, logits=[[[0.0, 0.0, 0.0]]], labels=[[1, 2]], label_length=[2], logit_length=[1],
          loss=[np.inf], gradient=np.zeros((1, 1, 3)), gradient_exact=True),
    _case(name="simplified_large_loss", ref="tests/test_simplified_ctc_loss.py:162-183", variant=SIMPLIFIED,
          logits=[[[1e10, 0.0, 0.0]]], labels=[[1]], label_length=[1], logit_length=[1],
          loss=[1e10], gradient=[[[0.0, -1.0, 0.0]]], gradient_exact=True),
    _case(name="simplified_length_one", ref="tests/test_simplified_ctc_loss.py:208-230", variant=SIMPLIFIED,
          logits=np.zeros((1, 1, 3)), labels=[[1]], label_length=[1], logit_length=[1],
          loss=[LN(3.0)], loss_places=6, gradient=[[[0.0, -1.0, 0.0]]]),
    _case(name="simplified_length_two", ref="tests/test_simplified_ctc_loss.py:232-258", variant=SIMPLIFIED,
          logits=np.zeros((1, 2, 3)), labels=[[1, 2]], label_length=[2], logit_length=[2],
          loss=[2 * LN(3.0)], loss_places=6, gradient=[[[0.0, -1.0, 0.0], [0.0, 0.0, -1.0]]]),
    _case(name="simplified_hessian_single_logit", ref="tests/test_hessian.py:37-60", variant=SIMPLIFIED,
          logits=_log([[[1 / 3, 1 / 3, 1 / 3]]]), labels=[[1]], label_length=[1], logit_length=[1],
          gradient=[[[0.0, -1.0, 0.0]]], hessian_zero=True),
]

# README.md:50-56 == tests/test_hessian.py:185-211; closed forms derived in SURVEY.md section 8(c)
README_EXAMPLE = _case(
    name="readme_example", ref="README.md:50-56", variant=CLASSIC,
    logits=np.zeros((2, 5, 3)), labels=[[1, 2, 2, 1], [1, 2, 1, 0]], label_length=[4, 3], logit_length=[5, 4])
README_GOLDEN = {
    "classic_loss": [LN(243.0), LN(81.0 / 7.0)],
    "simplified_loss": [LN(243.0 / 5.0), LN(81.0 / 4.0)],
    "classic_grad_logits_1": [[4 / 21, -11 / 21, 1 / 3], [4 / 21, 1 / 21, -5 / 21], [4 / 21, 1 / 21, -5 / 21],
                              [4 / 21, -11 / 21, 1 / 3], [0, 0, 0]],
    "classic_gradient_1": [[-1 / 7, -6 / 7, 0], [-1 / 7, -2 / 7, -4 / 7], [-1 / 7, -2 / 7, -4 / 7],
                           [-1 / 7, -6 / 7, 0], [0, 0, 0]],
    "simplified_grad_logits_0": [[2 / 15, -7 / 15, 1 / 3], [2 / 15, 2 / 15, -4 / 15], [2 / 15, 1 / 3, -7 / 15],
                                 [2 / 15, 2 / 15, -4 / 15], [2 / 15, -7 / 15, 1 / 3]],
    "simplified_grad_logits_1": [[1 / 12, -5 / 12, 1 / 3], [1 / 12, 1 / 12, -1 / 6], [1 / 12, 1 / 12, -1 / 6],
                                 [1 / 12, -5 / 12, 1 / 3], [0, 0, 0]],
    "classic_hessian_logits_abs_sum": 13.5782313,
}


def random_inputs(B, T, V, L, seed=0, ragged=True, blank=0, labels_width=None):
    """Synthetic inputs in the style of the reference's tests/common.py:53-104 (numpy RNG; the reference's
    tf.random streams cannot be reproduced without TensorFlow, so random cases pin properties, not vectors)."""
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((B, T, V)).astype(np.float32)
    if ragged:
        logit_length = rng.integers(max(T // 2, 1), T + 1, size=B).astype(np.int32)
        label_length = rng.integers(max(L // 2, 0), L + 1, size=B).astype(np.int32)
    else:
        logit_length = np.full((B,), T, dtype=np.int32)
        label_length = np.full((B,), L, dtype=np.int32)
    Lw = labels_width if labels_width is not None else L
    labels = rng.integers(0, V - 1, size=(B, Lw)).astype(np.int32)
    labels = np.where(labels >= blank, labels + 1, labels).astype(np.int32)   # never the blank
    return logits, labels, label_length, logit_length
