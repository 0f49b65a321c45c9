"""ctypes binding of libctc_b200.so (C ABI declared in include/ctc_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, this module raises.  The CPU
oracle under ``oracle/`` is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

CLASSIC = 0
SIMPLIFIED = 1
INPUT_LOGPROBAS = 1
FORCE_STAGED = 2
FORCE_FUSED = 4
TIME_MAJOR = 8
LOGITS_BF16 = 16
GRAD_BF16 = 32
WS_LOSS_GRAD, WS_STATES, WS_HESSIAN, WS_LOSS_GRAD_LOGITS, WS_HVP_LOGITS, WS_DECODE = 0, 1, 2, 3, 4, 5
MAX_STATES = 1024
MAX_TOKENS = 32768

_LIB_NAME = "libctc_b200.so"
_LIB_PATH = os.environ.get("CTCB200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME))

EXPORTED_SYMBOLS = (
    "ctcb200_version", "ctcb200_strerror", "ctcb200_last_cuda_error", "ctcb200_log_gradient", "ctcb200_stage_names", "ctcb200_launches_per_call",
    "ctcb200_workspace_bytes", "ctcb200_loss_grad", "ctcb200_states",
    "ctcb200_hessian", "ctcb200_hvp", "ctcb200_hvp_logits", "ctcb200_gamma", "ctcb200_greedy_decode", "ctcb200_host_create", "ctcb200_host_loss_grad",
    "ctcb200_host_grad_device_ptr", "ctcb200_host_destroy", "ctcb200_debug_fused_plan",
)


class Desc(ctypes.Structure):
    """struct ctcb200_desc."""
    _fields_ = [("B", ctypes.c_int32), ("T", ctypes.c_int32), ("V", ctypes.c_int32), ("Lw", ctypes.c_int32),
                ("blank", ctypes.c_int32), ("variant", ctypes.c_int32), ("U", ctypes.c_int32),
                ("flags", ctypes.c_uint32)]


class CtcB200Error(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Loads the shared library once; raises loudly when it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise CtcB200Error(
            f"{_LIB_PATH} not found: build it with `make -C tf_seq2seq_losses_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32p, fp = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
    dp = ctypes.POINTER(Desc)
    lib.ctcb200_version.restype = ctypes.c_int
    lib.ctcb200_strerror.restype = ctypes.c_char_p
    lib.ctcb200_strerror.argtypes = [ctypes.c_int]
    lib.ctcb200_last_cuda_error.restype = ctypes.c_char_p
    lib.ctcb200_stage_names.restype = ctypes.c_char_p
    lib.ctcb200_stage_names.argtypes = [dp]
    lib.ctcb200_log_gradient.restype = ctypes.c_int
    lib.ctcb200_log_gradient.argtypes = [dp, fp, i32p, i32p, i32p, fp, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_launches_per_call.restype = ctypes.c_int
    lib.ctcb200_launches_per_call.argtypes = [dp]
    lib.ctcb200_workspace_bytes.restype = ctypes.c_size_t
    lib.ctcb200_workspace_bytes.argtypes = [dp, ctypes.c_int]
    lib.ctcb200_loss_grad.restype = ctypes.c_int
    lib.ctcb200_loss_grad.argtypes = [dp, fp, i32p, i32p, i32p, fp, fp, fp, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_states.restype = ctypes.c_int
    lib.ctcb200_states.argtypes = [dp, fp, i32p, i32p, i32p, fp, fp, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_hessian.restype = ctypes.c_int
    lib.ctcb200_hessian.argtypes = [dp, fp, i32p, i32p, i32p, fp, fp, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_gamma.restype = ctypes.c_int
    lib.ctcb200_gamma.argtypes = [dp, fp, i32p, i32p, i32p, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_hvp.restype = ctypes.c_int
    lib.ctcb200_hvp.argtypes = [dp, fp, i32p, i32p, i32p, fp, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_hvp_logits.restype = ctypes.c_int
    lib.ctcb200_hvp_logits.argtypes = [dp, fp, i32p, i32p, i32p, fp, fp, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_greedy_decode.restype = ctypes.c_int
    lib.ctcb200_greedy_decode.argtypes = [dp, fp, i32p, ctypes.c_int, i32p, i32p, fp, vp, ctypes.c_size_t, vp]
    lib.ctcb200_host_create.restype = ctypes.c_int
    lib.ctcb200_host_create.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.ctcb200_host_loss_grad.restype = ctypes.c_int
    lib.ctcb200_host_loss_grad.argtypes = [vp, fp, i32p, i32p, i32p, fp, fp]
    lib.ctcb200_host_grad_device_ptr.restype = ctypes.c_void_p
    lib.ctcb200_host_grad_device_ptr.argtypes = [vp]
    lib.ctcb200_host_destroy.restype = None
    lib.ctcb200_host_destroy.argtypes = [vp]
    lib.ctcb200_debug_fused_plan.restype = None
    lib.ctcb200_debug_fused_plan.argtypes = [ctypes.c_int] * 5
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        detail = f": {load().ctcb200_last_cuda_error().decode()}" if code == -5 else ""
        raise CtcB200Error(f"libctc_b200: {load().ctcb200_strerror(code).decode()}{detail} (code {code})")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise CtcB200Error(f"{name} must live on a CUDA device: libctc_b200 has no CPU path")


# One workspace per (device, stream), grown on demand and reused by every later call on that stream: calls on one stream
# are ordered, so they can share scratch; no allocation happens on the steady-state path and the address is stable, which
# is what lets a call sit inside a captured CUDA graph.
_WORKSPACES: dict = {}


def _workspace(desc: Desc, what: int, device: torch.device) -> torch.Tensor:
    n = load().ctcb200_workspace_bytes(ctypes.byref(desc), what)
    if n == 0 and desc.B > 0:
        raise CtcB200Error("libctc_b200: descriptor rejected (unsupported size: more than "
                           f"{MAX_STATES} label states or {MAX_TOKENS} tokens, or an invalid field)")
    n = max(int(n), 256)
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < n:
        # torch's caching allocator returns 512-byte aligned blocks; the ABI needs 256
        ws = torch.empty(n, dtype=torch.uint8, device=device)
        if not torch.cuda.is_current_stream_capturing():     # memory from a graph's private pool stays with that graph
            _WORKSPACES[key] = ws
    return ws


def release_workspaces() -> None:
    """Drops the cached per-stream workspaces (they are re-created on the next call)."""
    _WORKSPACES.clear()


# Flags OR-ed into every descriptor built by make_desc.  Tests set this to FORCE_STAGED to exercise the three staged
# kernels on shapes the fused kernel would otherwise take (also settable with CTCB200_FORCE_STAGED=1).
DEFAULT_FLAGS = FORCE_STAGED if os.environ.get("CTCB200_FORCE_STAGED", "0") == "1" else 0


def make_desc(logits: torch.Tensor, labels: torch.Tensor, blank: int, variant: int, U: int, flags: int = 0) -> Desc:
    """Descriptor of a call on ``logits`` [B,T,V] (or [T,B,V] when ``flags`` has TIME_MAJOR)."""
    B, T, V = logits.shape
    flags = int(flags)
    if flags & TIME_MAJOR:
        B, T = T, B
    if logits.dtype == torch.bfloat16:
        flags |= LOGITS_BF16          # bf16 rows exist in the fused kernel only: never combined with FORCE_STAGED
        return Desc(B, T, V, labels.shape[1], int(blank), int(variant), int(U), flags | (DEFAULT_FLAGS & ~FORCE_STAGED))
    return Desc(B, T, V, labels.shape[1], int(blank), int(variant), int(U), flags | DEFAULT_FLAGS)


def _stream(device: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def loss_grad(desc: Desc, logits, labels, label_length, logit_length, d_loss=None, want_grad_logits=True,
              want_grad_logprobas=False, grad_logits_out=None):
    """ctcb200_loss_grad on the current stream of ``logits.device``.  Returns (loss, grad_logits, grad_logprobas)."""
    _require_cuda(logits, "logits")
    dev = logits.device
    loss = torch.empty((desc.B,), dtype=torch.float32, device=dev)
    gl = None
    if want_grad_logits:
        gdtype = torch.bfloat16 if (desc.flags & GRAD_BF16) else torch.float32
        gl = grad_logits_out if grad_logits_out is not None else torch.empty(logits.shape, dtype=gdtype, device=dev)
    gp = torch.empty_like(logits) if want_grad_logprobas else None
    ws = _workspace(desc, WS_LOSS_GRAD_LOGITS if (want_grad_logits and not want_grad_logprobas) else WS_LOSS_GRAD, dev)
    with torch.cuda.device(dev):
        check(load().ctcb200_loss_grad(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                       _ptr(logit_length), _ptr(d_loss), _ptr(loss), _ptr(gl), _ptr(gp),
                                       _ptr(ws), ws.numel(), _stream(dev)))
    return loss, gl, gp


def loss_only(desc: Desc, logits, labels, label_length, logit_length):
    """ctcb200_loss_grad with both gradient pointers NULL: the loss alone (half a call where the fused kernel applies)."""
    _require_cuda(logits, "logits")
    dev = logits.device
    loss = torch.empty((desc.B,), dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_LOSS_GRAD_LOGITS, dev)
    with torch.cuda.device(dev):
        check(load().ctcb200_loss_grad(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                       _ptr(logit_length), None, _ptr(loss), None, None, _ptr(ws), ws.numel(), _stream(dev)))
    return loss


def log_gradient(desc: Desc, logits, labels, label_length, logit_length):
    """ctcb200_log_gradient.  Returns (log_gradient [B,T,V], loss [B])."""
    _require_cuda(logits, "logits")
    dev = logits.device
    out = torch.empty((desc.B, desc.T, desc.V), dtype=torch.float32, device=dev)
    loss = torch.empty((desc.B,), dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_LOSS_GRAD, dev)
    if desc.B > 0:
        with torch.cuda.device(dev):
            check(load().ctcb200_log_gradient(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                              _ptr(logit_length), _ptr(loss), _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
    return out, loss


def states(desc: Desc, logits, labels, label_length, logit_length):
    """ctcb200_states.  Returns (alpha, beta, loss) in the reference layouts."""
    _require_cuda(logits, "logits")
    dev = logits.device
    shape = (desc.B, desc.T + 1, desc.U, 2) if desc.variant == CLASSIC else (desc.B, desc.T + 1, desc.U)
    alpha = torch.empty(shape, dtype=torch.float32, device=dev)
    beta = torch.empty(shape, dtype=torch.float32, device=dev)
    loss = torch.empty((desc.B,), dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_STATES, dev)
    with torch.cuda.device(dev):
        check(load().ctcb200_states(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                    _ptr(logit_length), _ptr(alpha), _ptr(beta), _ptr(loss), _ptr(ws), ws.numel(),
                                    _stream(dev)))
    return alpha, beta, loss


def gamma(desc: Desc, logits, labels, label_length, logit_length):
    """ctcb200_gamma.  Returns gamma in the reference layout ([B,T+1,U,2,T+1,U,2] classic, [B,T+1,U,T+1,U] simplified)."""
    _require_cuda(logits, "logits")
    dev = logits.device
    B, T, U = desc.B, desc.T, desc.U
    shape = (B, T + 1, U, 2, T + 1, U, 2) if desc.variant == CLASSIC else (B, T + 1, U, T + 1, U)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_STATES, dev)
    if B > 0:
        with torch.cuda.device(dev):
            check(load().ctcb200_gamma(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                       _ptr(logit_length), _ptr(out), _ptr(ws), ws.numel(), _stream(dev)))
    return out


def hessian(desc: Desc, logits, labels, label_length, logit_length):
    """ctcb200_hessian.  Returns (hessian [B,T,V,T,V], loss, grad_logprobas)."""
    _require_cuda(logits, "logits")
    dev = logits.device
    B, T, V = desc.B, desc.T, desc.V
    hess = torch.empty((B, T, V, T, V), dtype=torch.float32, device=dev)
    loss = torch.empty((B,), dtype=torch.float32, device=dev)
    g = torch.empty((B, T, V), dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_HESSIAN, dev)
    if B > 0 and T > 0:
        with torch.cuda.device(dev):
            check(load().ctcb200_hessian(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                         _ptr(logit_length), _ptr(hess), _ptr(loss), _ptr(g), _ptr(ws), ws.numel(),
                                         _stream(dev)))
    return hess, loss, g


def hvp(desc: Desc, logits, labels, label_length, logit_length, d_gradient):
    """ctcb200_hvp: sum_{t',k'} d_gradient[b,t',k'] * hessian[b,t,k,t',k'] -> [B,T,V]."""
    _require_cuda(logits, "logits")
    dev = logits.device
    out = torch.empty((desc.B, desc.T, desc.V), dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_HESSIAN, dev)
    if desc.B > 0 and desc.T > 0:
        d_gradient = d_gradient.contiguous()
        with torch.cuda.device(dev):
            check(load().ctcb200_hvp(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                     _ptr(logit_length), _ptr(d_gradient), _ptr(out), _ptr(ws), ws.numel(),
                                     _stream(dev)))
    return out


def hvp_logits(desc: Desc, logits, labels, label_length, logit_length, v, d_loss=None):
    """ctcb200_hvp_logits: d_loss[b] * (d2 loss[b] / d logits2) v[b] -> [B,T,V], matrix-free, log-softmax chain included."""
    _require_cuda(logits, "logits")
    dev = logits.device
    out = torch.zeros((desc.B, desc.T, desc.V), dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_HVP_LOGITS, dev)
    if desc.B > 0 and desc.T > 0:
        v = v.to(torch.float32).contiguous()
        d_loss = None if d_loss is None else d_loss.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            check(load().ctcb200_hvp_logits(ctypes.byref(desc), _ptr(logits), _ptr(labels), _ptr(label_length),
                                            _ptr(logit_length), _ptr(d_loss), _ptr(v), _ptr(out), _ptr(ws), ws.numel(),
                                            _stream(dev)))
    return out


def greedy_decode(logits, logit_length, blank: int = 0, merge_repeated: bool = True, time_major: bool = False):
    """ctcb200_greedy_decode.  Returns (decoded [B,T] int32 padded with -1, decoded_length [B], neg_sum_logits [B])."""
    _require_cuda(logits, "logits")
    dev = logits.device
    x = logits.detach().contiguous()
    B, T, V = x.shape
    if time_major:
        B, T = T, B
    desc = Desc(B, T, V, 0, int(blank), CLASSIC, 1, TIME_MAJOR if time_major else 0)
    tl = logit_length.to(device=dev, dtype=torch.int32).contiguous()
    decoded = torch.empty((B, T), dtype=torch.int32, device=dev)
    length = torch.empty((B,), dtype=torch.int32, device=dev)
    neg_sum = torch.empty((B,), dtype=torch.float32, device=dev)
    ws = _workspace(desc, WS_DECODE, dev)
    if B > 0:
        with torch.cuda.device(dev):
            check(load().ctcb200_greedy_decode(ctypes.byref(desc), _ptr(x), _ptr(tl), int(bool(merge_repeated)), _ptr(decoded),
                                               _ptr(length), _ptr(neg_sum), _ptr(ws), ws.numel(), _stream(dev)))
    return decoded, length, neg_sum


class HostContext:
    """ctcb200_host_*: loss + gradient from HOST buffers (pinned tensors), copies overlapped with the kernels."""

    def __init__(self, B, T, V, Lw, blank, variant, U, device=0, num_slices=8, flags=0):
        self.desc = Desc(B, T, V, Lw, blank, variant, U, int(flags))
        self.device = int(device)
        handle = ctypes.c_void_p()
        check(load().ctcb200_host_create(ctypes.byref(self.desc), self.device, int(num_slices), ctypes.byref(handle)))
        self._h = handle

    def loss_grad(self, logits, labels, label_length, logit_length, loss_out, grad_out=None):
        for name, t in (("logits", logits), ("labels", labels), ("label_length", label_length),
                        ("logit_length", logit_length), ("loss_out", loss_out)):
            if t.is_cuda or not t.is_contiguous():
                raise CtcB200Error(f"{name} must be a contiguous host tensor")
        check(load().ctcb200_host_loss_grad(self._h, _ptr(logits), _ptr(labels), _ptr(label_length),
                                            _ptr(logit_length), _ptr(loss_out), _ptr(grad_out)))
        return loss_out

    def grad_device_ptr(self) -> int:
        return int(load().ctcb200_host_grad_device_ptr(self._h) or 0)

    def close(self):
        if self._h:
            load().ctcb200_host_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
