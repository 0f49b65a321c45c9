"""Developer tool: where do the fused kernel's warps wait?  Needs the timing build of the library:
   nvcc ... -DCTCB200_FUSED_TIMING kf_fused.cu  ->  tf_seq2seq_losses_b200/libctc_b200_timing.so
   CTCB200_LIB=tf_seq2seq_losses_b200/libctc_b200_timing.so python tools/fused_timing.py [variant]
Prints, per warp role, the mean cycles spent in each phase and at each wait site (see kf_fused.cu)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_seq2seq_losses_b200 import _lib  # noqa: E402

B = int(os.environ.get("CTCB200_TIMING_B", "256"))
T, V, L = (int(x) for x in os.environ.get("CTCB200_TIMING_TVL", "1000,1024,200").split(","))
variant = _lib.CLASSIC if (len(sys.argv) > 1 and sys.argv[1] == "classic") else _lib.SIMPLIFIED
g = torch.Generator().manual_seed(0)
logits = torch.randn((B, T, V), generator=g).cuda()
labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).cuda()
ll = torch.full((B,), L, dtype=torch.int32).cuda()
tl = torch.full((B,), T, dtype=torch.int32).cuda()
desc = _lib.make_desc(logits, labels, 0, variant, L + 1)
lib = _lib.load()
n = lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_LOSS_GRAD)
ws = torch.zeros(n, dtype=torch.uint8, device="cuda")
loss = torch.empty(B, device="cuda")
grad = torch.empty_like(logits)
P = lambda t: ctypes.c_void_p(t.data_ptr())
for _ in range(3):
    _lib.check(lib.ctcb200_loss_grad(ctypes.byref(desc), P(logits), P(labels), P(ll), P(tl), None, P(loss), P(grad), None,
                                     P(ws), n, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
a256 = lambda x: (x + 255) & ~255
S = 2 if variant == _lib.CLASSIC else 1
NS = (L + 1 + 31) // 32
Upad = 32 * NS
rows, srows = B * T, B * (T + 1) * S * Upad
off = a256(rows * 4) * 2 + a256(rows * Upad * 4) + a256(srows * 4)       # byte offset of the beta scratch
W = int(os.environ.get("CTCB200_FUSED_W", "4"))        # workers per side the kernel picked (1 for V = 5000)
warps = 2 * (W + 1)
dbg = ws[off: off + B * warps * 12 * 8].view(torch.int64).reshape(B, warps, 12).cpu().numpy().astype(np.float64)
names = ["phaseA", "phaseB", "tma_wait", "rec:d_wait|work:gather", "ccount_wait", "scount_wait", "rec:done_wait|work:stats", "state_cpasync_wait",
         "B:softmax", "B:occupancy", "B:scatter", "B:blank+store"]
for wi in range(warps):
    side, role = divmod(wi, W + 1)
    m = dbg[:, wi].mean(axis=0)
    print(f"side {side} {'rec   ' if role == 0 else 'work'+str(role)+' '}", "  ".join(f"{nme}={v/1e3:8.1f}k" for nme, v in zip(names, m)))
