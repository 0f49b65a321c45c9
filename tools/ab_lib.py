"""Developer tool: time ctcb200_loss_grad of several builds of the library on the same box, same inputs.
   python tools/ab_lib.py classic|simplified B,T,V,L libA.so libB.so ...   (only the symbols every round's build has)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_seq2seq_losses_b200 import _lib  # noqa: E402  (Desc layout and constants only)

variant = _lib.CLASSIC if sys.argv[1] == "classic" else _lib.SIMPLIFIED
B, T, V, L = (int(x) for x in sys.argv[2].split(","))
g = torch.Generator().manual_seed(0)
logits = torch.randn((B, T, V), generator=g).cuda()
labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).cuda()
ll = torch.full((B,), L, dtype=torch.int32).cuda()
tl = torch.full((B,), T, dtype=torch.int32).cuda()
loss = torch.empty(B, device="cuda")
grad = torch.empty_like(logits)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for rep in range(2):
    for path in sys.argv[3:]:
        lib = ctypes.CDLL(os.path.abspath(path))
        vp = ctypes.c_void_p
        lib.ctcb200_workspace_bytes.restype = ctypes.c_size_t
        lib.ctcb200_workspace_bytes.argtypes = [ctypes.POINTER(_lib.Desc), ctypes.c_int]
        lib.ctcb200_loss_grad.restype = ctypes.c_int
        lib.ctcb200_loss_grad.argtypes = [ctypes.POINTER(_lib.Desc)] + [vp] * 9 + [ctypes.c_size_t, vp]
        desc = _lib.Desc(B, T, V, L, 0, variant, L + 1, 0)
        n = lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_LOSS_GRAD)
        ws = torch.empty(n, dtype=torch.uint8, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

        def fn():
            rc = lib.ctcb200_loss_grad(ctypes.byref(desc), P(logits), P(labels), P(ll), P(tl), None, P(loss), P(grad), None,
                                       P(ws), n, st)
            assert rc == 0, rc
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print(f"{sys.argv[1]} B={B} T={T} V={V} L={L} {os.path.basename(path):28s} {best*1e3:9.1f} us  loss[0]={loss[0].item():.4f}", flush=True)
        del ws
