"""Developer tool: device-time of the BASELINE.json configurations other than the headline one.
   python tools/bench_configs.py            (needs a B200)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tf_seq2seq_losses_b200 as pkg  # noqa: E402
from tf_seq2seq_losses_b200 import _lib  # noqa: E402


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def loss_grad_case(name, B, T, V, L, variant, staged, ragged=False):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn((B, T, V), generator=g).cuda()
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).cuda()
    if ragged:
        tl = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32).cuda()
        ll = torch.randint(L // 2, L + 1, (B,), generator=g, dtype=torch.int32).cuda()
    else:
        ll = torch.full((B,), L, dtype=torch.int32).cuda()
        tl = torch.full((B,), T, dtype=torch.int32).cuda()
    desc = _lib.make_desc(logits, labels, 0, variant, L + 1, _lib.FORCE_STAGED if staged else 0)
    lib = _lib.load()
    n = lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_LOSS_GRAD)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    loss = torch.empty(B, device="cuda")
    grad = torch.empty_like(logits)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: _lib.check(lib.ctcb200_loss_grad(ctypes.byref(desc), P(logits), P(labels), P(ll), P(tl), None, P(loss), P(grad),
                                                  None, P(ws), n, st))
    ms = timed(fn)
    alg = B * (8 * T * V + 4 * L + 12)
    print(f"{name:46s} {lib.ctcb200_stage_names(ctypes.byref(desc)).decode():40s} {ms*1e3:9.1f} us  {B/ms*1e3:10.0f} samples/s  "
          f"{alg/ms/1e6:7.0f} GB/s algorithmic ({alg/ms/1e6/6553*100:4.1f}% of 6553)")


def hessian_case(B, T, V, L):
    g = torch.Generator().manual_seed(0)
    lp = torch.log_softmax(torch.randn((B, T, V), generator=g), dim=2).cuda()
    labels = torch.randint(1, V, (B, L), generator=g, dtype=torch.int32).cuda()
    tl = torch.randint(T // 2, T + 1, (B,), generator=g, dtype=torch.int32).cuda()
    ll = torch.randint(L // 2, L + 1, (B,), generator=g, dtype=torch.int32).cuda()
    desc = _lib.make_desc(lp, labels, 0, _lib.CLASSIC, L + 1, _lib.INPUT_LOGPROBAS)
    lib = _lib.load()
    n = lib.ctcb200_workspace_bytes(ctypes.byref(desc), _lib.WS_HESSIAN)
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    loss = torch.empty(B, device="cuda")
    gbuf = torch.empty_like(lp)
    hess = torch.empty((B, T, V, T, V), device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: _lib.check(lib.ctcb200_hessian(ctypes.byref(desc), P(lp), P(labels), P(ll), P(tl), P(hess), P(loss), P(gbuf),
                                                P(ws), n, st))
    ms = timed(fn)
    out_bytes = hess.numel() * 4
    print(f"{'hessian classic B=%d T=%d V=%d L=%d (ragged)' % (B, T, V, L):46s} {'k1,k2,k3,k4_hessian':40s} {ms*1e3:9.1f} us  "
          f"{B/ms*1e3:10.0f} samples/s  {out_bytes/ms/1e6:7.0f} GB/s of output ({out_bytes/ms/1e6/6553*100:4.1f}% of 6553)")
    v = torch.randn_like(lp)
    out = torch.empty_like(lp)
    fn2 = lambda: _lib.check(lib.ctcb200_hvp(ctypes.byref(desc), P(lp), P(labels), P(ll), P(tl), P(v), P(out), P(ws), n, st))
    print(f"{'hvp (matrix-free) same shape':46s} {'k1,k2,k3,k4_hessian<hvp>':40s} {timed(fn2)*1e3:9.1f} us")


def readme_case(variant, name):
    """The reference's own benchmark shape (tests/benchmark.py:41-56, tests/common.py:53-104): B=256, T=255, V=32,
    labels [256,255], logit_length ~ U[127,255), label_length ~ U[63,127); forward only and loss + gradient, through the
    public Python face like the reference's benchmark (README.md:18-24 reports 0.138 / 0.28 classic and 0.0531 / 0.119
    simplified on a GTX 970, in its unit per 256-sample batch)."""
    B, T, V = 256, 255, 32
    g = torch.Generator().manual_seed(0)
    logits = torch.randn((B, T, V), generator=g).cuda()
    tl = torch.randint(T // 2, T, (B,), generator=g, dtype=torch.int32).cuda()
    ll = torch.randint(T // 4, T // 2, (B,), generator=g, dtype=torch.int32).cuda()
    labels = torch.randint(1, V, (B, T), generator=g, dtype=torch.int32).cuda()
    fn = pkg.classic_ctc_loss if variant == _lib.CLASSIC else pkg.simplified_ctc_loss
    U = int(ll.max().item())
    fwd = lambda: fn(labels, logits, ll, tl, 0, max_label_length=U)
    x = logits.clone().requires_grad_(True)

    def fwd_bwd():
        x.grad = None
        fn(labels, x, ll, tl, 0, max_label_length=U).sum().backward()

    ms_f, ms_g = timed(fwd), timed(fwd_bwd)
    print(f"{name:46s} forward {ms_f*1e3:8.1f} us   loss+gradient {ms_g*1e3:8.1f} us   ({B/ms_g*1e3:9.0f} samples/s)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "hessian":      # the Hessian configuration alone
        hessian_case(64, 50, 32, 15)
        sys.exit(0)
    readme_case(_lib.CLASSIC, "README benchmark classic B=256 T=255 V=32")
    readme_case(_lib.SIMPLIFIED, "README benchmark simplified B=256 T=255 V=32")
    loss_grad_case("cfg1 classic B=32 T=500 V=29 L=100", 32, 500, 29, 100, _lib.CLASSIC, False)
    loss_grad_case("cfg1 ragged", 32, 500, 29, 100, _lib.CLASSIC, False, ragged=True)
    loss_grad_case("cfg2 simplified B=256 T=1000 V=1024 L=200", 256, 1000, 1024, 200, _lib.SIMPLIFIED, False)
    loss_grad_case("cfg2 simplified staged", 256, 1000, 1024, 200, _lib.SIMPLIFIED, True)
    loss_grad_case("cfg2 classic", 256, 1000, 1024, 200, _lib.CLASSIC, False)
    loss_grad_case("cfg2 classic staged", 256, 1000, 1024, 200, _lib.CLASSIC, True)
    loss_grad_case("cfg2 simplified ragged", 256, 1000, 1024, 200, _lib.SIMPLIFIED, False, ragged=True)
    loss_grad_case("cfg4 slice classic B=256 T=1600 V=5000 L=400", 256, 1600, 5000, 400, _lib.CLASSIC, False)
    loss_grad_case("cfg4 slice classic staged", 256, 1600, 5000, 400, _lib.CLASSIC, True)
    loss_grad_case("small batch simplified B=32 (8-GPU strong slice)", 32, 1000, 1024, 200, _lib.SIMPLIFIED, False)
    hessian_case(64, 50, 32, 15)
