"""Developer tool: randomised parity sweep over mid-sized shapes (up to 4 states per lane, many repeated labels, ragged and
infeasible lengths, every blank position) on both device paths against the C restatement of the oracle.
   python tools/fuzz.py [trials] [seed]            (needs a B200; the committed suite holds the fixed-seed small sweep)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_seq2seq_losses_b200 as pkg  # noqa: E402
from oracle import c_oracle  # noqa: E402
from tf_seq2seq_losses_b200 import _lib  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = 0.0
for trial in range(trials):
    variant = int(rng.integers(0, 2))
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
        ll[rng.integers(0, B)] = Lw                        # possibly more labels than frames: infeasible
    tl = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    want_loss, want_grad = c_oracle.loss_grad(labels, logits, ll, tl, blank, variant)
    want_grad[np.isinf(want_loss)] = 0.0
    fn = pkg.classic_ctc_loss if variant == _lib.CLASSIC else pkg.simplified_ctc_loss
    for flags in (_lib.FORCE_FUSED, _lib.FORCE_STAGED):
        _lib.DEFAULT_FLAGS = flags
        x = torch.tensor(logits, device="cuda", requires_grad=True)
        loss = fn(torch.tensor(labels).cuda(), x, torch.tensor(ll).cuda(), torch.tensor(tl).cuda(), blank)
        torch.where(torch.isfinite(loss), loss, torch.zeros_like(loss)).sum().backward()
        got_loss, got = loss.detach().cpu().numpy(), x.grad.cpu().numpy()
        tag = (trial, variant, "fused" if flags == _lib.FORCE_FUSED else "staged", B, T, V, Lw, blank, ll.tolist(), tl.tolist())
        assert np.array_equal(np.isinf(got_loss), np.isinf(want_loss)), tag
        fin = np.isfinite(want_loss)
        assert np.all(np.abs(got_loss[fin] - want_loss[fin]) <= 1e-5 * np.maximum(1.0, np.abs(want_loss[fin]))), tag
        assert not np.isnan(got).any(), tag
        err = float(np.max(np.abs(got - want_grad)))
        worst = max(worst, err)
        assert err <= 5e-4, (err,) + tag        # peaky logits (x4) and near-infeasible alignments reach 1.4e-4 in fp32
_lib.DEFAULT_FLAGS = 0
print(f"fuzz ok: {trials} trials, worst gradient error {worst:.2e}")
