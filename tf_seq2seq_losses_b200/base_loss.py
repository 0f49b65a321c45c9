"""Host-side mirror of tf_seq2seq_losses/base_loss.py over libctc_b200.so.

Same names, argument meaning and error behaviour as the reference (``ctc_loss`` base_loss.py:38-68,
``ctc_loss_from_logproba`` :71-99, ``BaseCtcLossData`` :102-543), with torch CUDA tensors standing in for TensorFlow
tensors (TensorFlow is not part of this image; torch is only the device-memory / stream / autograd plumbing).  All
arithmetic happens in the CUDA library; there is no CPU path.

Differentiation mirrors the reference's three nested ``tf.custom_gradient`` (base_loss.py:140-184):
``forward_fn`` -> ``gradient_fn`` -> ``_hessian_fn``; the third derivative raises ``NotImplementedError``.
"""
from __future__ import annotations

from functools import cached_property
from typing import Optional, Union

import torch

from . import _lib

inf = float("inf")


def _as_int32(t, device) -> torch.Tensor:
    t = torch.as_tensor(t)
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.to(device).contiguous()


def _blank_to_int(blank_index) -> int:
    # base_loss.py:122-125: blank_index may be a python int or a (scalar) tensor
    if isinstance(blank_index, torch.Tensor):
        return int(blank_index.item())
    return int(blank_index)


def _max_label_length_plus_one(label_length: torch.Tensor, max_label_length: Optional[int], labels_width: int) -> int:
    """base_loss.py:478-486 (reduce_max_with_default with default 0): the number of label states U.

    Any U >= max(label_length) + 1 gives the same loss and gradient (states beyond an utterance's label are unreachable);
    U only sets how many states the kernels carry.  So no device->host read-back is ever *needed*:
      * ``max_label_length`` (keyword extension) is trusted when given -- it must not be smaller than the true maximum,
        labels beyond it are cut off (the kernels clamp label_length to U - 1);
      * host-resident lengths are reduced on the host;
      * device-resident lengths are read back (one small synchronising copy) -- except while a CUDA graph is being
        captured, where the bound labels.shape[1] + 1 is used instead.
    """
    if max_label_length is not None:
        if not label_length.is_cuda and label_length.numel() > 0:      # free to check on the host: refuse to truncate labels
            assert int(label_length.max()) <= int(max_label_length), "max_label_length is smaller than max(label_length)"
        return int(max_label_length) + 1
    if label_length.numel() == 0:
        return 1
    if label_length.is_cuda and torch.cuda.is_current_stream_capturing():
        return int(labels_width) + 1
    return max(int(label_length.max().item()), 0) + 1


class _ForwardFn(torch.autograd.Function):
    """forward_fn, base_loss.py:140-155: returns loss; backprop = d_loss[:,None,None] * gradient_fn(logprobas)."""

    @staticmethod
    def forward(ctx, logprobas, data):
        ctx.data = data
        ctx.save_for_backward(logprobas)
        return data.loss.clone()

    @staticmethod
    def backward(ctx, d_loss):
        (logprobas,) = ctx.saved_tensors
        return d_loss[:, None, None] * _GradientFn.apply(logprobas, ctx.data), None


class _GradientFn(torch.autograd.Function):
    """gradient_fn, base_loss.py:157-175: returns gradient; backprop contracts d_gradient with the Hessian."""

    @staticmethod
    def forward(ctx, logprobas, data):
        ctx.data = data
        ctx.save_for_backward(logprobas)
        return data.gradient.clone()

    @staticmethod
    def backward(ctx, d_gradient):
        (logprobas,) = ctx.saved_tensors
        return _HessianVectorFn.apply(logprobas, d_gradient, ctx.data), None


class _HessianVectorFn(torch.autograd.Function):
    """sum_{t',k'} d_gradient * _hessian_fn(logprobas) (base_loss.py:167-184), matrix-free on the device."""

    @staticmethod
    def forward(ctx, logprobas, d_gradient, data):
        ctx.data = data
        return data.hessian_vector_product(d_gradient)

    @staticmethod
    def backward(ctx, d_out):
        # base_loss.py:179-182 raises for the derivative w.r.t. logprobas; the op is linear in d_gradient and the
        # Hessian is symmetric, so that cotangent is another Hessian-vector product.
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("Third order derivative over the ctc loss function is not implemented.")
        return None, ctx.data.hessian_vector_product(d_out), None


class BaseCtcLossData:
    """Mirror of BaseCtcLossData (base_loss.py:102-543): lazily evaluated loss / gradient / hessian / alpha / beta,
    each computed by one call into libctc_b200.so.  Takes *logprobas* (not logits), like the reference."""

    _variant: int = -1

    def __init__(self, labels, logprobas, label_length, logit_length, blank_index: Union[int, torch.Tensor],
                 swap_memory: bool = False, max_label_length: Optional[int] = None, **kwargs):
        super().__init__(**kwargs)
        self._logprobas = logprobas
        self._original_label = torch.as_tensor(labels)
        self._logit_length = torch.as_tensor(logit_length)
        self._original_label_length = torch.as_tensor(label_length)
        self._verify_inputs()
        self._blank_index = _blank_to_int(blank_index)
        self._swap_memory = swap_memory      # stored and unused, as in the reference (base_loss.py:112,127)
        dev = logprobas.device
        self._lp = logprobas.detach().contiguous()
        self._labels32 = _as_int32(self._original_label, dev)
        self._label_length32 = _as_int32(self._original_label_length, dev)
        self._logit_length32 = _as_int32(self._logit_length, dev)
        self._U = _max_label_length_plus_one(self._original_label_length, max_label_length,
                                             self._original_label.shape[1])
        self._desc = _lib.make_desc(self._lp, self._labels32, self._blank_index, self._variant, self._U,
                                    _lib.INPUT_LOGPROBAS)

    def _verify_inputs(self) -> None:
        # base_loss.py:129-138
        assert len(self._logprobas.shape) == 3
        assert self._logprobas.dtype == torch.float32
        assert len(self._original_label.shape) == 2
        assert len(self._logit_length.shape) == 1
        assert len(self._original_label_length.shape) == 1
        assert self._logprobas.shape[0] == self._original_label.shape[0]
        assert self._logprobas.shape[0] == self._logit_length.shape[0]
        assert self._logprobas.shape[0] == self._original_label_length.shape[0]

    def _args(self):
        return self._lp, self._labels32, self._label_length32, self._logit_length32

    # ---- differentiable entry points (base_loss.py:140-184) --------------------------------------------------
    def forward_fn(self, unused_logprobas: torch.Tensor) -> torch.Tensor:
        return _ForwardFn.apply(unused_logprobas, self)

    def gradient_fn(self, unused_logprobas: torch.Tensor) -> torch.Tensor:
        return _GradientFn.apply(unused_logprobas, self)

    # ---- values ---------------------------------------------------------------------------------------------
    @cached_property
    def _loss_and_gradient(self):
        loss, _, g = _lib.loss_grad(self._desc, *self._args(), want_grad_logits=False, want_grad_logprobas=True)
        return loss, g

    @property
    def loss(self) -> torch.Tensor:
        """[B]  (classic_ctc_loss.py:152-165 / simplified_ctc_loss.py:73-83)."""
        return self._loss_and_gradient[0]

    @property
    def gradient(self) -> torch.Tensor:
        """d loss / d logproba, [B,T,V]  (base_loss.py:262-268)."""
        return self._loss_and_gradient[1]

    @cached_property
    def logarithmic_logproba_gradient(self) -> torch.Tensor:
        """log(-gradient), [B,T,V]  (base_loss.py:270-298), computed in the log domain on the device: finite wherever a
        (frame, token) pair is possible at all -- also far below exp's underflow at -87 -- and -inf exactly elsewhere."""
        return _lib.log_gradient(self._desc, *self._args())[0]

    @cached_property
    def hessian(self) -> torch.Tensor:
        """d2 loss / d logproba2, [B,T,V,T,V]  (base_loss.py:186-260)."""
        return _lib.hessian(self._desc, *self._args())[0]

    def hessian_vector_product(self, d_gradient: torch.Tensor) -> torch.Tensor:
        """sum over (t',k') of d_gradient[b,t',k'] * hessian[b,t,k,t',k']  (gradient_fn.backprop, base_loss.py:167-173)."""
        return _lib.hvp(self._desc, *self._args(), d_gradient.to(torch.float32))

    @cached_property
    def _states(self):
        alpha, beta, _ = _lib.states(self._desc, *self._args())
        return alpha, beta

    @property
    def alpha(self) -> torch.Tensor:
        """[B,T+1,U,2] classic / [B,T+1,U] simplified."""
        return self._states[0]

    @property
    def beta(self) -> torch.Tensor:
        return self._states[1]

    @cached_property
    def gamma(self) -> torch.Tensor:
        """Transition log-probabilities between any two states, [B,T+1,U,2,T+1,U,2] classic / [B,T+1,U,T+1,U] simplified
        (classic_ctc_loss.py:167-308 / simplified_ctc_loss.py:85-191).  O(T^2 U^2): small shapes only (U <= 128).  The
        Hessian kernels do not use it; it completes the data-class surface."""
        return _lib.gamma(self._desc, *self._args())


def ctc_loss_from_logproba(labels, logprobas, label_length, logit_length, blank_index, ctc_loss_data_cls,
                           max_label_length: Optional[int] = None) -> torch.Tensor:
    """base_loss.py:71-99: loss as a (twice differentiable) function of the log-probabilities."""
    loss_data = ctc_loss_data_cls(labels=labels, logprobas=logprobas.detach(), label_length=label_length,
                                  logit_length=logit_length, blank_index=blank_index,
                                  max_label_length=max_label_length)
    return loss_data.forward_fn(logprobas)


# ---- fused logits path (the hot path): log-softmax + loss + d/dlogits in one library call ---------------------
# Two schedules for a training step (forward, then backward with an upstream gradient d_loss):
#   deferred  forward = loss-only call (the fused kernel stops at the middle: half the T-step chain, logits read once);
#             backward = the full loss+gradient call with d_loss applied inside the kernel.  No [B,T,V] tensor lives
#             between the passes, none is re-scaled.  Best where the call is bandwidth-bound: B=256 T=1000 V=1024
#             0.25 + 0.58 ms against 0.58 + 0.35 ms (the extra pass of the eager multiply).
#   eager     forward = the full call (d_loss = 1), the gradient is kept; backward = d_loss[:,None,None] * gradient like
#             the reference (base_loss.py:150-153).  Best where the T-step chain is the runtime and a [B,T,V] pass is
#             cheap: the reference's benchmark shape B=256 T=255 V=32 takes 170 us this way, 240 us deferred.
# Below EAGER_MAX_ELEMENTS logits elements the eager schedule is used.
EAGER_MAX_ELEMENTS = 1 << 24


class _FusedLossFn(torch.autograd.Function):
    """ctc_loss (base_loss.py:38-68) with the log-softmax (tools.py:27-40) and its backward fused into the kernels;
    forward_fn / gradient_fn of the reference (base_loss.py:140-175) in one of the two schedules above."""

    @staticmethod
    def forward(ctx, logits, labels, label_length, logit_length, desc):
        x = logits.detach().contiguous()
        ctx.desc, ctx.aux = desc, (labels, label_length, logit_length)
        ctx.eager = bool(logits.requires_grad and x.numel() < EAGER_MAX_ELEMENTS)
        if ctx.eager:
            loss, grad, _ = _lib.loss_grad(desc, x, labels, label_length, logit_length)
            ctx.save_for_backward(logits, grad)
        else:
            loss = _lib.loss_only(desc, x, labels, label_length, logit_length)
            ctx.save_for_backward(logits)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        logits = ctx.saved_tensors[0]
        unit_grad = ctx.saved_tensors[1] if ctx.eager else None
        return _FusedGradFn.apply(logits, d_loss, unit_grad, ctx.desc, ctx.aux), None, None, None, None


class _FusedGradFn(torch.autograd.Function):
    """d_loss * d loss / d logits, differentiable once more (Hessian w.r.t. logits).  `unit_grad` is the
    gradient for d_loss = 1 when the forward pass already produced it (eager schedule), else None: the fused
    loss+gradient call runs here with d_loss applied in the kernel."""

    @staticmethod
    def forward(ctx, logits, d_loss, unit_grad, desc, aux):
        ctx.desc, ctx.aux = desc, aux
        ctx.save_for_backward(logits, d_loss)
        labels, label_length, logit_length = aux
        if unit_grad is not None:
            dl = d_loss.to(torch.float32)
            dl = dl[None, :, None] if desc.flags & _lib.TIME_MAJOR else dl[:, None, None]
            if unit_grad.dtype == torch.float32:
                return dl * unit_grad
            return (dl * unit_grad.to(torch.float32)).to(unit_grad.dtype)     # bf16 gradient: multiply in fp32, round once
        dl = d_loss.detach().to(torch.float32).contiguous()
        _, grad, _ = _lib.loss_grad(desc, logits.detach().contiguous(), labels, label_length, logit_length, d_loss=dl)
        return grad

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, v):
        logits, d_loss = ctx.saved_tensors
        labels, label_length, logit_length = ctx.aux
        time_major = bool(ctx.desc.flags & _lib.TIME_MAJOR)
        if time_major and ctx.needs_input_grad[0]:
            raise NotImplementedError("second derivative w.r.t. time-major logits: pass batch-major logits")
        if (ctx.desc.flags & _lib.LOGITS_BF16) and ctx.needs_input_grad[0]:
            raise NotImplementedError("second derivative w.r.t. bfloat16 logits: pass float32 logits")
        x = logits.detach().contiguous()
        d_logits = None
        if ctx.needs_input_grad[0]:
            # (d2 loss / d logits2) v = J^T H J v - s (p.v - p p^T v), J = I - 1 p^T per frame: one library call
            # (ctcb200_hvp_logits: K1, K2, K3, hvp_pre, K4<HVP>, hvp_post)
            d_logits = _lib.hvp_logits(ctx.desc, x, labels, label_length, logit_length, v, d_loss)
        d_d_loss = None
        if ctx.needs_input_grad[1]:      # the cotangent of d_loss: <v, d loss / d logits> per utterance
            _, grad, _ = _lib.loss_grad(ctx.desc, x, labels, label_length, logit_length)
            d_d_loss = (v.float() * grad.float()).sum(dim=(0, 2) if time_major else (1, 2))
        return d_logits, d_d_loss, None, None, None


def ctc_loss(labels, logits, label_length, logit_length, blank_index, ctc_loss_data_cls,
             max_label_length: Optional[int] = None, logits_time_major: bool = False) -> torch.Tensor:
    """base_loss.py:38-68.  Returns the per-sample loss [B]; differentiable twice w.r.t. ``logits``.
    ``logits_time_major`` (keyword extension, the reference is batch-major only): ``logits`` is [T,B,V] and so is its
    gradient; only the first derivative is available in that layout."""
    assert len(logits.shape) == 3
    # float32 like the reference (base_loss.py:131); bfloat16 logits are an extension (the kernels widen them on the fly,
    # all arithmetic stays fp32, the gradient comes back in bfloat16; first derivative only)
    assert logits.dtype in (torch.float32, torch.bfloat16)
    labels_t, ll_t, tl_t = torch.as_tensor(labels), torch.as_tensor(label_length), torch.as_tensor(logit_length)
    assert len(labels_t.shape) == 2 and len(ll_t.shape) == 1 and len(tl_t.shape) == 1
    assert logits.shape[1 if logits_time_major else 0] == labels_t.shape[0] == ll_t.shape[0] == tl_t.shape[0]
    dev = logits.device
    labels32, ll32, tl32 = _as_int32(labels_t, dev), _as_int32(ll_t, dev), _as_int32(tl_t, dev)
    U = _max_label_length_plus_one(ll_t, max_label_length, labels_t.shape[1])
    desc = _lib.make_desc(logits, labels32, _blank_to_int(blank_index), ctc_loss_data_cls._variant, U,
                          (_lib.TIME_MAJOR if logits_time_major else 0) |
                          (_lib.GRAD_BF16 if logits.dtype == torch.bfloat16 else 0))
    return _FusedLossFn.apply(logits, labels32, ll32, tl32, desc)
