"""ctypes loader for oracle/libctc_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_DIR, "libctc_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _DIR], check=True)
        _lib = ctypes.CDLL(_PATH)
    return _lib


def max_threads() -> int:
    return int(load().ctc_oracle_max_threads())


def loss_grad(labels, logits, label_length, logit_length, blank=0, variant=0, want_grad=True, dtype=np.float64,
              num_threads=0):
    """loss [B] and d loss / d logits [B,T,V] computed by the C restatement in ``dtype`` arithmetic."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    ll = np.ascontiguousarray(label_length, dtype=np.int32)
    tl = np.ascontiguousarray(logit_length, dtype=np.int32)
    B, T, V = logits.shape
    dt = np.dtype(dtype)
    fn = load().ctc_oracle_loss_grad_f64 if dt == np.float64 else load().ctc_oracle_loss_grad_f32
    loss = np.empty((B,), dtype=dt)
    grad = np.empty((B, T, V), dtype=dt) if want_grad else None
    vp = ctypes.c_void_p
    rc = fn(ctypes.c_int(variant), ctypes.c_int(B), ctypes.c_int(T), ctypes.c_int(V), ctypes.c_int(labels.shape[1]),
            ctypes.c_int(blank), vp(logits.ctypes.data), vp(labels.ctypes.data), vp(ll.ctypes.data), vp(tl.ctypes.data),
            vp(loss.ctypes.data), vp(grad.ctypes.data) if want_grad else vp(0), ctypes.c_int(num_threads))
    assert rc == 0
    return loss, grad
