"""Developer tool: device time of the loss+gradient call against the batch size (the strong-scaling slices of the
headline shape), for the library named by CTCB200_LIB.   python tools/bsweep.py [simplified|classic] [B,B,...]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_seq2seq_losses_b200 import _lib  # noqa: E402

T, V, L = (int(x) for x in os.environ.get("CTCB200_TVL", "1000,1024,200").split(","))
variant = _lib.CLASSIC if (len(sys.argv) > 1 and sys.argv[1] == "classic") else _lib.SIMPLIFIED
Bs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256, 128, 64, 32, 16]
lib = _lib.load()
if os.environ.get("CTCB200_PLAN"):       # "W,SL,XA,R,mode" -> ctcb200_debug_fused_plan
    lib.ctcb200_debug_fused_plan(*(int(x) for x in os.environ["CTCB200_PLAN"].split(",")))
g = torch.Generator().manual_seed(0)
for B in Bs:
    logits = torch.randn((B, T, V), generator=g).cuda()
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).cuda()
    ll = torch.full((B,), L, dtype=torch.int32).cuda()
    tl = torch.full((B,), T, dtype=torch.int32).cuda()
    desc = _lib.make_desc(logits, labels, 0, variant, L + 1, int(os.environ.get("CTCB200_FLAGS", "0")))
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
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    alg = B * (8 * T * V + 4 * L + 12)
    print(f"B={B:4d} T={T} V={V} L={L} {'classic' if variant == _lib.CLASSIC else 'simplified':10s} {lib.ctcb200_stage_names(ctypes.byref(desc)).decode()[:9]:9s} {best*1e3:8.1f} us  "
          f"{B/best*1e3:10.0f} samples/s  {alg/best/1e6/6553*100:5.1f}% of 6553 GB/s   {best*1e6/T:6.1f} ns/frame", flush=True)
    del logits, grad, ws
