"""Regenerates tests/golden/*.npz.

The reference (Python over TensorFlow) cannot be imported in this image, so these fixtures are produced by the CPU
oracle (oracle/ctc_oracle.py, float64), which is itself pinned against the reference's literal known-answer tests
(tests/test_oracle_kats.py).  They freeze inputs *and* outputs, so a later change to the oracle or to the kernels
shows up as a diff against a committed file:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ctc_oracle as orc  # noqa: E402
from tests.ref_cases import random_inputs  # noqa: E402

CASES = {
    # name: (B, T, V, L, seed, blank, ragged)
    "small_ragged": (4, 12, 6, 4, 101, 0, True),
    "blank_last": (3, 20, 8, 7, 102, 7, True),
    "repeats_v4": (5, 33, 4, 12, 103, 0, True),       # tiny vocabulary: many repeated labels (classic repeat rule)
    "aligned_v64": (3, 64, 64, 30, 104, 0, False),
}


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (B, T, V, L, seed, blank, ragged) in CASES.items():
        logits, labels, ll, tl = random_inputs(B, T, V, L, seed=seed, ragged=ragged, blank=blank)
        if name == "small_ragged":
            ll[1], tl[1] = 4, 2          # infeasible sample
            ll[2] = 0                    # empty label
        arrays = dict(logits=logits, labels=labels, label_length=ll, logit_length=tl, blank=np.int32(blank))
        for vname, variant in (("classic", orc.CLASSIC), ("simplified", orc.SIMPLIFIED)):
            loss, grad, data = orc.loss_and_grad_logits(labels, logits, ll, tl, blank, variant)
            grad = np.where(np.isinf(loss)[:, None, None], 0.0, grad)
            arrays[f"{vname}_loss"] = loss
            arrays[f"{vname}_grad_logits"] = grad
            arrays[f"{vname}_gradient"] = data.gradient          # w.r.t. log-probabilities (data-class surface)
            arrays[f"{vname}_alpha"] = data.alpha
            arrays[f"{vname}_beta"] = data.beta
            if T <= 20:
                arrays[f"{vname}_hessian"] = data.hessian_fast()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
