"""Developer tool: default / forced-fused / forced-staged device time (and per-stage times of the staged path) for the
narrow-vocabulary and fallback shapes; this is what the dispatch rule in csrc/api.cu (fused_workers) is tuned on.
   python tools/exp_paths.py            (needs a B200)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tf_seq2seq_losses_b200 import _lib  # noqa: E402
import bench_configs as bc  # noqa: E402


def case(name,B,T,V,L,variant,flags,ragged=False):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn((B, T, V), generator=g).cuda()
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).cuda()
    if ragged:
        tl = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32).cuda()
        ll = torch.randint(L // 4, L//2 + 1, (B,), generator=g, dtype=torch.int32).cuda()
    else:
        ll = torch.full((B,), L, dtype=torch.int32).cuda(); tl = torch.full((B,), T, dtype=torch.int32).cuda()
    U = int(ll.max().item()) + 1
    lib = _lib.load()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    loss = torch.empty(B, device="cuda"); grad = torch.empty_like(logits)
    def run(fl):
        desc = _lib.make_desc(logits, labels, 0, variant, U, fl)
        n = lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_LOSS_GRAD)
        ws = torch.empty(n, dtype=torch.uint8, device="cuda")
        fn = lambda: _lib.check(lib.ctcb200_loss_grad(ctypes.byref(desc), P(logits), P(labels), P(ll), P(tl), None, P(loss), P(grad), None, P(ws), n, st))
        return bc.timed(fn), lib.ctcb200_stage_names(ctypes.byref(desc)).decode()
    for fl,nm in flags:
        ms, stages = run(fl)
        extra = ""
        if stages != "kf_fused":
            parts=[]
            for i,sn in enumerate(stages.split(",")):
                m,_ = run(fl | ((1<<i)<<8)); parts.append(f"{sn}={m*1e3:.1f}")
            extra = " ".join(parts)
        print(f"{name:40s} {nm:8s} {ms*1e3:9.1f} us  {stages:40s} {extra}", flush=True)
F=[(0,"default"),(_lib.FORCE_FUSED,"fused"),(_lib.FORCE_STAGED,"staged")]
case("cfg1 classic B32 T500 V29 L100",32,500,29,100,_lib.CLASSIC,F)
case("readme classic B256 T255 V32",256,255,32,255,_lib.CLASSIC,F,ragged=True)
case("readme simplified B256 T255 V32",256,255,32,255,_lib.SIMPLIFIED,F,ragged=True)
case("cfg2 simplified",256,1000,1024,200,_lib.SIMPLIFIED,[(_lib.FORCE_STAGED,"staged")])
case("cfg2 classic",256,1000,1024,200,_lib.CLASSIC,[(_lib.FORCE_STAGED,"staged")])
case("V64 classic B64 T400 V64 L80",64,400,64,80,_lib.CLASSIC,F)
case("V128 simplified B128 T400 V128 L80",128,400,128,80,_lib.SIMPLIFIED,F)
