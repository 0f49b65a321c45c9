"""Developer tool: device time of the headline shapes for the library named by CTCB200_LIB (A/B of build variants).
   CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_X.so python tools/ab.py [simple|classic|big|all]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_seq2seq_losses_b200 import _lib  # noqa: E402


def case(name, B, T, V, L, variant, steps=30):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn((B, T, V), generator=g).cuda()
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).cuda()
    ll = torch.full((B,), L, dtype=torch.int32).cuda()
    tl = torch.full((B,), T, dtype=torch.int32).cuda()
    desc = _lib.make_desc(logits, labels, 0, variant, L + 1, 0)
    lib = _lib.load()
    n = lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_LOSS_GRAD)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    loss = torch.empty(B, device="cuda")
    grad = torch.empty_like(logits)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: _lib.check(lib.ctcb200_loss_grad(ctypes.byref(desc), P(logits), P(labels), P(ll), P(tl), None, P(loss),
                                                  P(grad), None, P(ws), n, st))
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / steps)
    ms = min(times)
    alg = B * (8 * T * V + 4 * L + 12)
    print(f"{os.path.basename(os.environ.get('CTCB200_LIB', 'default')):28s} {name:34s} {ms*1e3:8.1f} us  {alg/ms/1e6/6553*100:5.1f}%  "
          f"loss_sum={loss.double().sum().item():.6f} grad_abs={grad.double().abs().sum().item():.6f}", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "simple"
if which in ("simple", "all"):
    case("simplified B256 T1000 V1024 L200", 256, 1000, 1024, 200, _lib.SIMPLIFIED)
if which in ("classic", "all"):
    case("classic B256 T1000 V1024 L200", 256, 1000, 1024, 200, _lib.CLASSIC)
if which in ("big", "all"):
    case("classic B256 T1600 V5000 L400", 256, 1600, 5000, 400, _lib.CLASSIC, steps=5)
