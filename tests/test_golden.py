"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the float64 oracle and
re-computed by the REFERENCE'S OWN CODE -- every loss / gradient / alpha / beta / Hessian array within 1e-12 -- by
`tests/golden/make_reference_golden.py --check`, which tests/test_reference_golden.py runs wherever /root/reference exists).
CPU: the oracle and its C port still reproduce them.  GPU: the CUDA path matches them."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import ctc_oracle as orc

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
VARIANTS = (("classic", orc.CLASSIC), ("simplified", orc.SIMPLIFIED))


def _load(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def test_golden_files_exist():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = _load(path)
    blank = int(g["blank"])
    for vname, variant in VARIANTS:
        loss, grad, data = orc.loss_and_grad_logits(g["labels"], g["logits"], g["label_length"], g["logit_length"], blank, variant)
        assert np.array_equal(np.isinf(loss), np.isinf(g[f"{vname}_loss"]))
        fin = np.isfinite(loss)
        assert np.max(np.abs(loss[fin] - g[f"{vname}_loss"][fin])) < 1e-10
        grad = np.where(np.isinf(loss)[:, None, None], 0.0, grad)
        assert np.max(np.abs(grad - g[f"{vname}_grad_logits"])) < 1e-10
        closs, cgrad = c_oracle.loss_grad(g["labels"], g["logits"], g["label_length"], g["logit_length"], blank, variant)
        assert np.max(np.abs(closs[fin] - g[f"{vname}_loss"][fin])) < 1e-9
        assert np.max(np.abs(cgrad - g[f"{vname}_grad_logits"])) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_matches_golden(path):
    import torch
    import tf_seq2seq_losses_b200 as pkg
    g = _load(path)
    blank = int(g["blank"])
    cu = lambda a: torch.as_tensor(a).cuda()
    for vname, variant in VARIANTS:
        fn = pkg.classic_ctc_loss if variant == orc.CLASSIC else pkg.simple_ctc_loss
        cls = pkg.ClassicCtcLossData if variant == orc.CLASSIC else pkg.SimplifiedCtcLossData
        x = cu(g["logits"]).requires_grad_(True)
        loss = fn(cu(g["labels"]), x, cu(g["label_length"]), cu(g["logit_length"]), blank)
        torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
        got = loss.detach().cpu().numpy().astype(np.float64)
        want = g[f"{vname}_loss"]
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        assert np.all(np.abs(got[fin] - want[fin]) <= 1e-5 * np.maximum(1.0, np.abs(want[fin])))
        assert np.max(np.abs(x.grad.cpu().numpy() - g[f"{vname}_grad_logits"])) <= 5e-5
        data = cls(cu(g["labels"]), torch.log_softmax(cu(g["logits"]), dim=2), cu(g["label_length"]), cu(g["logit_length"]), blank)
        assert np.max(np.abs(data.gradient.cpu().numpy() - g[f"{vname}_gradient"])) <= 5e-5
        for name, arr in (("alpha", data.alpha), ("beta", data.beta)):
            a, w = arr.cpu().numpy(), g[f"{vname}_{name}"]
            assert np.array_equal(np.isinf(a), np.isinf(w))
            f2 = np.isfinite(w)
            assert np.max(np.abs(a[f2] - w[f2]) / np.maximum(1.0, np.abs(w[f2]))) < 1e-5
        if f"{vname}_hessian" in g:
            assert np.max(np.abs(data.hessian.cpu().numpy() - g[f"{vname}_hessian"])) <= 5e-5
