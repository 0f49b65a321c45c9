"""Developer tool: per-utterance / per-frame error of the fused kernel on one small shape."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ctc_oracle as orc
from tests.ref_cases import random_inputs
from tf_seq2seq_losses_b200 import _lib
B, T, V, L = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3,6,5,3").split(","))
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
np.set_printoptions(precision=4, suppress=True, linewidth=200)
logits, labels, ll, tl = random_inputs(B, T, V, L, seed=B * 1000 + T)
want_loss, want_grad, _ = orc.loss_and_grad_logits(labels, logits, ll, tl, 0, variant)
x, lab, llc, tlc = (torch.as_tensor(a).cuda() for a in (logits, labels, ll, tl))
for rep in range(3):
    desc = _lib.make_desc(x, lab, 0, variant, L + 1, _lib.FORCE_FUSED)
    loss, grad, _ = _lib.loss_grad(desc, x, lab, llc, tlc)
    torch.cuda.synchronize()
    g = grad.cpu().numpy()
    print("rep", rep, "ll", ll, "tl", tl)
    print(" loss got ", loss.cpu().numpy(), "\n loss want", want_loss)
    err = np.abs(g - np.where(np.isinf(want_loss)[:, None, None], 0, want_grad)).max(axis=2)
    print(" per-frame max grad err:\n", err)
    print(" row sums got:\n", g.sum(axis=2))
