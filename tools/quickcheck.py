"""Developer tool: a handful of small shapes through the fused kernel against the oracle, under several forced plans.
   CTCB200_LIB=... python tools/quickcheck.py ["plan name"]   (one process per plan survives a trapped kernel)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctc_oracle as orc  # noqa: E402
from tests.ref_cases import random_inputs  # noqa: E402
from tf_seq2seq_losses_b200 import _lib  # noqa: E402

lib = _lib.load()
SHAPES = [(3, 6, 5, 3), (8, 64, 10, 30), (8, 20, 8, 9), (4, 33, 29, 12), (5, 61, 96, 20), (6, 200, 64, 40), (90, 40, 64, 9), (170, 14, 64, 5),
          (4, 50, 132, 40), (4, 50, 130, 40), (3, 40, 256, 100), (150, 30, 128, 20)]
PLANS = {"default": (0, 0, 0, 0, 0), "nohalf W4": (4, 2, 1, 8, 2), "half W4": (4, 2, 1, 8, 0), "W2 R4": (2, 2, 0, 4, 0),
         "split W8": (8, 3, 1, 16, 1), "split W8 nohalf": (8, 3, 1, 16, 3), "W1 R1": (1, 2, 0, 1, 0)}
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, plan in PLANS.items():
    if only is not None and name != only:
        continue
    lib.ctcb200_debug_fused_plan(*plan)
    for (B, T, V, L) in SHAPES:
        for variant in (0, 1):
            logits, labels, ll, tl = random_inputs(B, T, V, L, seed=B * 1000 + T)
            want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant)
            want_grad[np.isinf(want_loss)] = 0.0
            x, lab, llc, tlc = (torch.as_tensor(a).cuda() for a in (logits, labels, ll, tl))
            desc = _lib.make_desc(x, lab, 0, variant, L + 1, _lib.FORCE_FUSED)
            try:
                loss, grad, _ = _lib.loss_grad(desc, x, lab, llc, tlc)
                only = _lib.loss_only(desc, x, lab, llc, tlc)
                torch.cuda.synchronize()
                fin = np.isfinite(want_loss)
                le = np.max(np.abs(loss.cpu().numpy()[fin] - want_loss[fin])) if fin.any() else 0.0
                lo = np.max(np.abs(only.cpu().numpy()[fin] - want_loss[fin])) if fin.any() else 0.0
                ge = np.max(np.abs(grad.cpu().numpy() - want_grad))
                flag = "" if (ge < 5e-4 and le < 1e-3 and lo < 1e-3) else "   <<<<<< BAD"
                print(f"{name:16s} B{B} T{T} V{V} L{L} variant {variant}: loss err {le:.2e} loss-only err {lo:.2e} grad err {ge:.2e}{flag}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"{name:16s} B{B} T{T} V{V} L{L} variant {variant}: EXCEPTION {str(e)[:150]}", flush=True)
                sys.exit(1)
lib.ctcb200_debug_fused_plan(0, 0, 0, 0, 0)
