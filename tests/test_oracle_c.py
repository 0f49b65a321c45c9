"""Pins the C restatement (oracle/ctc_oracle.c) against the numpy oracle and the reference's known answers.  CPU only."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import ctc_oracle as orc
from tests.ref_cases import CLASSIC, KAT_CASES, SIMPLIFIED, random_inputs


@pytest.mark.parametrize("case", [c for c in KAT_CASES if "loss" in c], ids=lambda c: c["name"])
def test_c_oracle_known_answer_losses(case):
    loss, _ = c_oracle.loss_grad(case["labels"], case["logits"], case["label_length"], case["logit_length"],
                                 case["blank"], case["variant"])
    if "loss_places" in case:
        assert np.max(np.abs(loss - np.asarray(case["loss"]))) < 0.5 * 10.0 ** -case["loss_places"]
    else:
        assert loss.tolist() == [float(v) for v in case["loss"]]


@pytest.mark.parametrize("variant", [CLASSIC, SIMPLIFIED])
@pytest.mark.parametrize("shape", [(3, 6, 5, 3, 0, None), (8, 64, 10, 30, 0, None), (4, 33, 29, 12, 3, None),
                                   (2, 50, 32, 15, 0, 9), (4, 20, 8, 9, 0, 20)])
def test_c_oracle_matches_numpy_oracle(shape, variant):
    B, T, V, L, blank, lw = shape
    logits, labels, ll, tl = random_inputs(B, T, V, L, seed=B + T, blank=blank, labels_width=lw)
    if lw is not None and lw < L:
        ll = np.minimum(ll, lw + 2).astype(np.int32)
    ll[0] = min(ll[0], 1)
    want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, blank, variant)
    loss, grad = c_oracle.loss_grad(labels, logits, ll, tl, blank, variant)
    assert np.array_equal(np.isinf(loss), np.isinf(want_loss))
    fin = ~np.isinf(want_loss)
    assert np.max(np.abs(loss[fin] - want_loss[fin])) < 1e-9
    want_grad = np.where(np.isinf(want_loss)[:, None, None], 0.0, want_grad)
    assert np.max(np.abs(grad - want_grad)) < 1e-9
    loss32, grad32 = c_oracle.loss_grad(labels, logits, ll, tl, blank, variant, dtype=np.float32)
    assert np.max(np.abs(loss32[fin] - want_loss[fin]) / np.maximum(1, np.abs(want_loss[fin]))) < 1e-5
    assert np.max(np.abs(grad32 - want_grad)) < 3e-4      # reference-like fp32 log-domain arithmetic drifts with T
