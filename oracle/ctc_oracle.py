"""CPU restatement (numpy) of the CTC hot path of alexeytochin/tf_seq2seq_losses.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; it is the
checker used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.

Parity status: PINNED.  The reference is pure Python over TensorFlow and TensorFlow is not installable in
this image; its own code is nevertheless executed here with ``tensorflow`` served by a numpy implementation
of the ops it calls (``tests/golden/tf_numpy_shim.py``), and this restatement reproduces the outputs of that
run -- loss, gradient, alpha, beta, Hessian, gamma -- bit for bit (``tests/golden/reference/reference_outputs.npz``,
``tests/test_reference_golden.py``).  It is also pinned against
every literal known-answer test the reference's own test-suite holds for the path (see
``tests/test_oracle_kats.py``; the vectors are re-expressed there with the reference file:line), against
``torch.nn.functional.ctc_loss`` (an independent native implementation of the classic loss) and against
``torch.autograd`` first/second derivatives of a differentiable restatement.

Every function cites the reference lines it follows (paths relative to ``/root/reference``).
All functions take/return numpy arrays; ``dtype`` selects float64 (checker) or float32 (reference-like
arithmetic, used for the CPU baseline timing).

Conventions: B batch, T = logits.shape[1], V tokens, U = max(label_length)+1, blank = blank index.
Classic state tensors carry a trailing axis s (0 = closed, 1 = open).
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf
CLASSIC = 0
SIMPLIFIED = 1


# --------------------------------------------------------------------------------------------------
# tools.py numerics
# --------------------------------------------------------------------------------------------------
def reduce_logsumexp(x: np.ndarray, axis, keepdims: bool = False) -> np.ndarray:
    """tf.reduce_logsumexp: m = max(x) with non-finite m replaced by 0; log(sum(exp(x - m))) + m.

    Used by tf_seq2seq_losses/tools.py:37 and throughout classic_ctc_loss.py / simplified_ctc_loss.py.
    """
    x = np.asarray(x)
    if x.size == 0:
        return np.full(np.max(x, axis=axis, keepdims=keepdims, initial=NEG_INF).shape, NEG_INF, x.dtype)
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0).astype(x.dtype)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
    if not keepdims:
        out = np.squeeze(out, axis=axis)
    return out


def logit_to_logproba(logit: np.ndarray, axis: int = 2) -> np.ndarray:
    """tf_seq2seq_losses/tools.py:27-40."""
    return logit - reduce_logsumexp(logit, axis=axis, keepdims=True)


def logsumexp2(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Elementwise log(e^x + e^y), tf_seq2seq_losses/tools.py:57-71 (x == y, incl. both -inf, gives x + log 2)."""
    with np.errstate(invalid="ignore"):
        return np.logaddexp(x, y)


def apply_logarithmic_mask(x: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """x + log(float(mask)), tf_seq2seq_losses/tools.py:43-54."""
    with np.errstate(divide="ignore"):
        return x + np.log(mask.astype(x.dtype))


def unsorted_segment_logsumexp(data: np.ndarray, segment_ids: np.ndarray, num_segments: int) -> np.ndarray:
    """tf_seq2seq_losses/tools.py:95-119.  data [N, ...], segment_ids [N] -> [num_segments, ...].

    unsorted_segment_max initialises with the lowest *finite* float, so an empty or all -inf segment
    ends as lowest + log(0) = -inf rather than NaN (pinned by tests/test_tools.py:137-148).
    """
    lowest = np.finfo(data.dtype).min
    seg_max = np.full((num_segments,) + data.shape[1:], lowest, dtype=data.dtype)
    np.maximum.at(seg_max, segment_ids, data)
    normed = data - seg_max[segment_ids]
    seg_sum = np.zeros_like(seg_max)
    np.add.at(seg_sum, segment_ids, np.exp(normed))
    with np.errstate(divide="ignore"):
        return seg_max + np.log(seg_sum)


# --------------------------------------------------------------------------------------------------
# BaseCtcLossData (base_loss.py:102-543) with both concrete variants folded in by ``variant``
# --------------------------------------------------------------------------------------------------
class CtcLossData:
    """Restatement of BaseCtcLossData + ClassicCtcLossData / SimplifiedCtcLossData.

    Constructor arguments mirror base_loss.py:105-114: it takes *logprobas*, not logits.
    """

    def __init__(self, labels, logprobas, label_length, logit_length, blank_index=0, variant=CLASSIC,
                 dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.variant = int(variant)
        self._logprobas = np.asarray(logprobas, dtype=self.dtype)
        self._original_label = np.asarray(labels, dtype=np.int64)
        self._logit_length = np.asarray(logit_length, dtype=np.int64)
        self._label_length = np.asarray(label_length, dtype=np.int64)
        self.blank = int(blank_index)
        # base_loss.py:129-138
        assert self._logprobas.ndim == 3
        assert self._original_label.ndim == 2
        assert self._logit_length.ndim == 1 and self._label_length.ndim == 1
        assert self._logprobas.shape[0] == self._original_label.shape[0]
        assert self._logprobas.shape[0] == self._logit_length.shape[0]
        assert self._logprobas.shape[0] == self._label_length.shape[0]
        self.B, self.T, self.V = self._logprobas.shape
        # base_loss.py:482-486 (reduce_max_with_default, default 0)
        self.Lmax = int(self._label_length.max()) if self.B > 0 else 0
        self.U = self.Lmax + 1
        self._cache = {}

    # ---- small cached-property helper -----------------------------------------------------------
    def _c(self, name, fn):
        if name not in self._cache:
            self._cache[name] = fn()
        return self._cache[name]

    # ---- masks / labels -------------------------------------------------------------------------
    @property
    def logit_length_mask(self):
        """base_loss.py:500-506 -> [B,T] bool."""
        return np.arange(self.T)[None, :] < self._logit_length[:, None]

    @property
    def label_length_mask(self):
        """base_loss.py:508-513 -> [B,U] bool."""
        return np.arange(self.U)[None, :] < self._label_length[:, None]

    @property
    def label(self):
        """_cleaned_label, base_loss.py:395-418: truncate / pad to U columns, blank beyond label_length."""
        def f():
            lab = self._original_label
            if lab.shape[1] > self.Lmax:
                lab = lab[:, : self.U]
            if lab.shape[1] < self.U:
                pad = np.full((self.B, self.U - lab.shape[1]), self.blank, dtype=lab.dtype)
                lab = np.concatenate([lab, pad], axis=1)
            return np.where(self.label_length_mask, lab, self.blank)
        return self._c("label", f)

    @property
    def preceded_label(self):
        """base_loss.py:519-525: roll(label, +1)."""
        return np.roll(self.label, 1, axis=1)

    # ---- log-probabilities ----------------------------------------------------------------------
    @property
    def logproba(self):
        """_logproba, base_loss.py:378-393: rows t >= logit_length become log one_hot(blank)."""
        def f():
            blank_row = np.full((self.V,), NEG_INF, dtype=self.dtype)
            if self.V > 0:
                blank_row[self.blank] = 0.0
            return np.where(self.logit_length_mask[:, :, None], self._logprobas, blank_row[None, None, :])
        return self._c("logproba", f)

    @property
    def blank_logproba(self):
        """base_loss.py:365-371 -> h [B,T]."""
        return self.logproba[:, :, self.blank]

    def _gather_tokens(self, lp, idx):
        """tf.gather(params=lp [B,T,V], indices=idx [B,U], axis=2, batch_dims=1) -> [B,T,U]."""
        if self.B == 0:
            return np.zeros((0, self.T, idx.shape[1]), dtype=self.dtype)
        return np.take_along_axis(lp, np.broadcast_to(idx[:, None, :], (self.B, self.T, idx.shape[1])), axis=2)

    @property
    def expected_token_logproba(self):
        """_expected_token_logproba, base_loss.py:328-344 -> d [B,T,U] (-inf for l >= label_length)."""
        def f():
            return apply_logarithmic_mask(self._gather_tokens(self.logproba, self.label),
                                          self.label_length_mask[:, None, :])
        return self._c("d", f)

    # classic-only tables --------------------------------------------------------------------------
    @property
    def open_to_open_diagonal(self):
        """classic_ctc_loss.py:478-492 -> d1 [B,T,U]: d where label[l] != label[l-1], else -inf."""
        def f():
            rep_mask = self.label != np.roll(self.label, 1, axis=1)
            return apply_logarithmic_mask(self.expected_token_logproba, rep_mask[:, None, :])
        return self._c("d1", f)

    @property
    def any_to_open_diagonal(self):
        """classic_ctc_loss.py:464-476 -> [B,T,U,2] (s = state before the step)."""
        return np.stack([self.expected_token_logproba, self.open_to_open_diagonal], axis=3)

    @property
    def not_blank_horizontal(self):
        """classic_ctc_loss.py:528-543 -> r [B,T,U]: logproba of re-emitting label[l-1] (-inf if it is blank)."""
        def f():
            mask = np.ones((self.V,), dtype=bool)
            if self.V > 0:
                mask[self.blank] = False
            nb = apply_logarithmic_mask(self.logproba, mask[None, None, :])
            return self._gather_tokens(nb, self.preceded_label)
        return self._c("r", f)

    @property
    def previous_label_token_logproba(self):
        """classic_ctc_loss.py:545-558 -> [B,T,U] (blank not masked)."""
        return self._c("plp", lambda: self._gather_tokens(self.logproba, self.preceded_label))

    @property
    def horizontal_step(self):
        """classic_ctc_loss.py:503-526 -> [B,T,U,next,prev]."""
        def f():
            h = self.blank_logproba
            blank_term = np.broadcast_to(h[:, :, None, None], (self.B, self.T, self.U, 2))
            non_blank = np.stack([np.full_like(self.not_blank_horizontal, NEG_INF), self.not_blank_horizontal], axis=3)
            return np.stack([blank_term, non_blank], axis=3)
        return self._c("hstep", f)

    # ---- alpha / beta ---------------------------------------------------------------------------
    @property
    def alpha(self):
        """classic_ctc_loss.py:379-462 ([B,T+1,U,2]) / simplified_ctc_loss.py:358-438 ([B,T+1,U])."""
        return self._c("alpha", self._alpha)

    def _alpha(self):
        B, T, U = self.B, self.T, self.U
        if self.variant == SIMPLIFIED:
            out = np.full((B, T + 1, U), NEG_INF, dtype=self.dtype)
            out[:, 0, 0] = 0.0                                   # simplified_ctc_loss.py:426-438
            h, d = self.blank_logproba, self.expected_token_logproba
            for t in range(T):                                   # unfold, tools.py:191-277
                out[:, t + 1] = self.simplified_alpha_step(out[:, t], h[:, t], d[:, t])
            return out
        out = np.full((B, T + 1, U, 2), NEG_INF, dtype=self.dtype)
        out[:, 0, 0, 0] = 0.0                                    # classic_ctc_loss.py:453-462
        for t in range(T):
            out[:, t + 1] = self.classic_alpha_step(out[:, t], t)
        return out

    @staticmethod
    def simplified_alpha_step(prev, h_t, d_t):
        """simplified_ctc_loss.py:393-424.  prev [..., U]; h_t [...]; d_t [..., U]."""
        horizontal = h_t[..., None] + prev
        diagonal = np.roll(d_t + prev, 1, axis=-1)
        return logsumexp2(horizontal, diagonal)

    def classic_alpha_step(self, prev, t):
        """classic_ctc_loss.py:415-451.  prev [B,U,2] -> [B,U,2]."""
        temp = self.horizontal_step[:, t] + prev[:, :, None, :]            # [B,U,next,prev]
        horizontal = reduce_logsumexp(temp, axis=3)                          # [B,U,2]
        diag = reduce_logsumexp(self.any_to_open_diagonal[:, t] + prev, axis=2)   # [B,U]
        moved = np.roll(diag, 1, axis=1)
        diagonal = np.stack([np.full_like(moved, NEG_INF), moved], axis=2)   # out state is always open
        return logsumexp2(horizontal, diagonal)

    @property
    def beta(self):
        """classic_ctc_loss.py:310-377 / simplified_ctc_loss.py:291-356."""
        return self._c("beta", self._beta)

    def _beta(self):
        B, T, U = self.B, self.T, self.U
        with np.errstate(divide="ignore"):
            last = np.log((np.arange(U)[None, :] == self._label_length[:, None]).astype(self.dtype))
        if self.variant == SIMPLIFIED:
            out = np.full((B, T + 1, U), NEG_INF, dtype=self.dtype)
            out[:, T] = last                                     # simplified_ctc_loss.py:345-356
            h, d = self.blank_logproba, self.expected_token_logproba
            for t in range(T - 1, -1, -1):
                out[:, t] = self.simplified_beta_step(out[:, t + 1], h[:, t], d[:, t])
            return out
        out = np.full((B, T + 1, U, 2), NEG_INF, dtype=self.dtype)
        out[:, T] = last[:, :, None]                             # classic_ctc_loss.py:366-377
        for t in range(T - 1, -1, -1):
            out[:, t] = self.classic_beta_step(out[:, t + 1], t)
        return out

    @staticmethod
    def simplified_beta_step(prev, h_t, d_t):
        """simplified_ctc_loss.py:327-343."""
        horizontal = h_t[..., None] + prev
        diagonal = d_t + np.roll(prev, -1, axis=-1)
        return logsumexp2(horizontal, diagonal)

    def classic_beta_step(self, prev, t):
        """classic_ctc_loss.py:349-364.  prev [B,U,2] (time t+1) -> [B,U,2] (time t)."""
        horizontal = reduce_logsumexp(self.horizontal_step[:, t] + prev[:, :, :, None], axis=2)
        diagonal = self.any_to_open_diagonal[:, t] + np.roll(prev[:, :, 1:], -1, axis=1)
        return logsumexp2(horizontal, diagonal)

    # ---- loss -----------------------------------------------------------------------------------
    @property
    def loss(self):
        """classic_ctc_loss.py:152-165 / simplified_ctc_loss.py:73-83 -> [B]."""
        def f():
            params = self.alpha[:, -1]
            if self.variant == CLASSIC:
                params = reduce_logsumexp(params, axis=-1)
            if self.B == 0:
                return np.zeros((0,), dtype=self.dtype)
            return -np.take_along_axis(params, self._label_length[:, None], axis=1)[:, 0]
        return self._c("loss", f)

    # ---- combine / gradient ---------------------------------------------------------------------
    def select_from_act(self, act, label):
        """_select_from_act, base_loss.py:420-468.  act [B,A,T,U,D], label [B,U] -> [B,A,T,V,D]."""
        B, A, T, U, D = act.shape
        data = np.transpose(act, (0, 3, 2, 1, 4)).reshape(B * U, T, A, D)
        seg = (label + np.arange(B)[:, None] * self.V).reshape(-1)
        out = unsorted_segment_logsumexp(data, seg, B * self.V)          # [B*V, T, A, D]
        return np.transpose(out.reshape(B, self.V, T, A, D), (0, 3, 2, 1, 4))

    def combine_transition_probabilities(self, a, b):
        """_combine_transition_probabilities: classic_ctc_loss.py:565-669 / simplified_ctc_loss.py:456-534.

        a [B, *DA, T, U(,2)], b [B, T, U(,2), *DB] -> [B, *DA, T, V, *DB].
        """
        B, T, U, V = self.B, self.T, self.U, self.V
        blank_mask = (np.arange(V) == self.blank)[None, None, None, :, None]
        if self.variant == SIMPLIFIED:
            dims_a, dims_b = a.shape[1:-2], b.shape[3:]
            a = a.reshape(B, int(np.prod(dims_a, dtype=np.int64)), T, U, 1)
            b = b.reshape(B, 1, T, U, int(np.prod(dims_b, dtype=np.int64)))
            ab = a + b
            blank_term = self.blank_logproba[:, None, :, None] + reduce_logsumexp(ab, axis=3)
            act = a + self.expected_token_logproba[:, None, :, :, None] + np.roll(b, -1, axis=3)
            non_blank = self.select_from_act(act, self.label)
        else:
            dims_a, dims_b = a.shape[1:-3], b.shape[4:]
            a = a.reshape(B, int(np.prod(dims_a, dtype=np.int64)), T, U, 2, 1)
            b = b.reshape(B, 1, T, U, 2, int(np.prod(dims_b, dtype=np.int64)))
            ab = reduce_logsumexp(a, axis=4) + b[:, :, :, :, 0]
            blank_term = self.blank_logproba[:, None, :, None] + reduce_logsumexp(ab, axis=3)
            act_h = a[:, :, :, :, 1] + self.previous_label_token_logproba[:, None, :, :, None] + b[:, :, :, :, 1]
            horizontal_non_blank = self.select_from_act(act_h, self.preceded_label)
            inp = a + self.any_to_open_diagonal[:, None, :, :, :, None] + np.roll(b[:, :, :, :, 1:], -1, axis=3)
            act_d = reduce_logsumexp(inp, axis=4)
            diagonal_non_blank = self.select_from_act(act_d, self.label)
            non_blank = logsumexp2(horizontal_non_blank, diagonal_non_blank)
        out = np.where(blank_mask, blank_term[:, :, :, None, :], non_blank)
        return out.reshape((B,) + tuple(dims_a) + (T, V) + tuple(dims_b))

    @property
    def logarithmic_logproba_gradient(self):
        """base_loss.py:270-298 -> [B,T,V]."""
        def f():
            with np.errstate(invalid="ignore"):
                lg = self.loss[:, None, None] + self.combine_transition_probabilities(self.alpha[:, :-1], self.beta[:, 1:])
            lg = np.where((self.loss == np.inf)[:, None, None], NEG_INF, lg)
            return apply_logarithmic_mask(lg, self.logit_length_mask[:, :, None])
        return self._c("lg", f)

    @property
    def gradient(self):
        """base_loss.py:262-268: d loss / d logproba = -exp(lg) -> [B,T,V]."""
        return self._c("gradient", lambda: -np.exp(self.logarithmic_logproba_gradient))

    # ---- gamma / hessian (literal; only for tiny shapes) ----------------------------------------
    @property
    def gamma(self):
        """classic_ctc_loss.py:167-308 ([B,T+1,U,2,T+1,U,2]) / simplified_ctc_loss.py:85-191,279-289 ([B,T+1,U,T+1,U])."""
        return self._c("gamma", self._gamma)

    def _gamma(self):
        B, T, U = self.B, self.T, self.U
        t_idx = np.arange(T + 1)
        with np.errstate(divide="ignore"):
            if self.variant == SIMPLIFIED:
                diag = np.log(np.eye(U, dtype=self.dtype))[None, None]            # [1,1,U,U]
                cur = np.broadcast_to(diag, (B, T + 1, U, U)).copy()
                slices = [cur]
                h, d = self.blank_logproba, self.expected_token_logproba
                for i in range(T):                                                  # gamma_step :146-191
                    horizontal = h[:, i][:, None, None, None] + cur
                    diagonal = np.roll(d[:, i][:, None, None, :] + cur, 1, axis=3)
                    new = logsumexp2(horizontal, diagonal)
                    cur = np.where((t_idx <= i)[None, :, None, None], new, diag)
                    slices.append(cur)
                fwd = np.transpose(np.stack(slices, 0), (1, 2, 3, 0, 4))            # [B,T+1,U,T+1,U]
                mask = (t_idx[:, None] <= t_idx[None, :])[None, :, None, :, None]
                return apply_logarithmic_mask(fwd, mask)
            diag = np.log(np.eye(2 * U, dtype=self.dtype)).reshape(1, 1, U, 2, U, 2)
            cur = np.broadcast_to(diag, (B, T + 1, U, 2, U, 2)).copy()
            slices = [cur]
            for i in range(T):                                                      # gamma_step :219-284
                hs = self.horizontal_step[:, i][:, None, None, None] + cur[:, :, :, :, :, None, :]
                horizontal = reduce_logsumexp(hs, axis=6)
                dsl = reduce_logsumexp(self.any_to_open_diagonal[:, i][:, None, None, None] + cur, axis=5)
                moved = np.roll(dsl, 1, axis=4)
                diagonal = np.stack([np.full_like(moved, NEG_INF), moved], axis=5)
                new = logsumexp2(horizontal, diagonal)
                cur = np.where((t_idx <= i)[None, :, None, None, None, None], new, diag)
                slices.append(cur)
            fwd = np.transpose(np.stack(slices, 0), (1, 2, 3, 4, 0, 5, 6))
            mask = (t_idx[:, None] <= t_idx[None, :])[None, :, None, None, :, None, None]
            return apply_logarithmic_mask(fwd, mask)

    @property
    def hessian(self):
        """base_loss.py:186-260, literal (materialises gamma): d2 loss / d logproba2 -> [B,T,V,T,V]."""
        return self._c("hessian", self._hessian_literal)

    def _hessian_literal(self):
        B, T, V = self.B, self.T, self.V
        with np.errstate(invalid="ignore", over="ignore"):
            ag = self.combine_transition_probabilities(self.alpha[:, :-1], self.gamma[:, 1:])
            agb = self.combine_transition_probabilities(ag[:, :, :, :-1], self.beta[:, 1:])
            first = self.loss[:, None, None, None, None] + agb
            n = T * V
            first = first.reshape(B, n, n).copy()
            idx = np.arange(n)
            first[:, idx, idx] = self.logarithmic_logproba_gradient.reshape(B, n)
            first = first.reshape(B, T, V, T, V)
            t_idx = np.arange(T)
            upper = (t_idx[:, None] <= t_idx[None, :])[None, :, None, :, None]
            sym = np.where(upper, first, np.transpose(first, (0, 3, 4, 1, 2)))
            g = self.gradient
            hess = -np.exp(sym) + g[:, :, :, None, None] * g[:, None, None, :, :]
        hess = np.where((self.loss == np.inf)[:, None, None, None, None], 0.0, hess)
        m = self.logit_length_mask
        hess = np.where(m[:, :, None, None, None], hess, 0.0)
        hess = np.where(m[:, None, None, :, None], hess, 0.0)
        return hess.astype(self.dtype)

    # ---- matrix-free Hessian (same quantity, O(T^2 U) per token; usable at cfg-4 size) ----------
    def hessian_fast(self):
        """Same tensor as ``hessian`` without materialising gamma.

        For (t,k): push alpha[t] through "emit k at t", propagate with the alpha step, and combine with
        beta at every t' > t (this is what the two _combine calls of base_loss.py:192-198 contract to).
        """
        B, T, V, U = self.B, self.T, self.V, self.U
        out = np.zeros((B, T, V, T, V), dtype=self.dtype)
        g = self.gradient
        lg = self.logarithmic_logproba_gradient
        lab, plab = self.label, self.preceded_label
        h, d = self.blank_logproba, self.expected_token_logproba
        for b in range(B):
            if self.loss[b] == np.inf:
                continue
            n_t = int(min(max(self._logit_length[b], 0), T))
            toks = sorted(set(lab[b, : int(self._label_length[b])].tolist()) | {self.blank})
            for t in range(n_t):
                for k in toks:
                    w = self._emit(b, t, k, self.alpha[b, t])
                    for t2 in range(t + 1, n_t):
                        c = self._combine_one(b, t2, w, self.beta[b, t2 + 1])      # [V]
                        with np.errstate(over="ignore"):
                            out[b, t, k, t2, :] = -np.exp(self.loss[b] + c)
                        w = self._step_one(b, t2, w)
                out[b, t, :, t + 1 : n_t, :] += g[b, t, :, None, None] * g[b, None, t + 1 : n_t, :]
                out[b, t + 1 : n_t, :, t, :] = np.transpose(out[b, t, :, t + 1 : n_t, :], (1, 2, 0))
                # same-time block: diagonal -exp(lg) + g^2, off-diagonal g g'
                blk = g[b, t, :, None] * g[b, t, None, :]
                blk[np.arange(V), np.arange(V)] += -np.exp(lg[b, t])
                out[b, t, :, t, :] = blk
        return out

    def _step_one(self, b, t, w):
        if self.variant == SIMPLIFIED:
            return self.simplified_alpha_step(w, self.blank_logproba[b, t], self.expected_token_logproba[b, t])
        temp = self.horizontal_step[b, t] + w[:, None, :]
        horizontal = reduce_logsumexp(temp, axis=2)
        diag = np.roll(reduce_logsumexp(self.any_to_open_diagonal[b, t] + w, axis=1), 1)
        return logsumexp2(horizontal, np.stack([np.full_like(diag, NEG_INF), diag], axis=1))

    def _emit(self, b, t, k, a):
        """State after consuming frame t *with token k emitted* (a = state log-weights before frame t)."""
        U = self.U
        lab, plab = self.label[b], self.preceded_label[b]
        if self.variant == SIMPLIFIED:
            if k == self.blank:
                return self.blank_logproba[b, t] + a
            term = np.where(lab == k, a + self.expected_token_logproba[b, t], NEG_INF)
            return np.roll(term, 1)
        out = np.full((U, 2), NEG_INF, dtype=self.dtype)
        if k == self.blank:
            out[:, 0] = self.blank_logproba[b, t] + reduce_logsumexp(a, axis=1)
            return out
        stay = np.where(plab == k, a[:, 1] + self.previous_label_token_logproba[b, t], NEG_INF)
        move = reduce_logsumexp(a + self.any_to_open_diagonal[b, t], axis=1)
        move = np.roll(np.where(lab == k, move, NEG_INF), 1)
        out[:, 1] = logsumexp2(stay, move)
        return out

    def _combine_one(self, b, t, a, bt):
        """_combine_transition_probabilities for one (b,t): a, bt state vectors -> [V]."""
        res = np.full((self.V,), NEG_INF, dtype=self.dtype)
        toks = sorted(set(self.label[b, : int(self._label_length[b])].tolist()))
        for k in toks:
            if k == self.blank:
                continue
            res[k] = reduce_logsumexp(self._emit(b, t, k, a) + bt, axis=None)
        res[self.blank] = reduce_logsumexp(self._emit(b, t, self.blank, a) + bt, axis=None)
        return res


# --------------------------------------------------------------------------------------------------
# Loss driver (base_loss.py:38-99) and the log-softmax chain TF autodiff adds on top
# --------------------------------------------------------------------------------------------------
def ctc_loss_data(labels, logits, label_length, logit_length, blank_index=0, variant=CLASSIC, dtype=np.float64):
    """base_loss.py:38-68: log-softmax (tools.py:27-40) then the data class on the log-probabilities."""
    logprobas = logit_to_logproba(np.asarray(logits, dtype=dtype), axis=2)
    return CtcLossData(labels, logprobas, label_length, logit_length, blank_index, variant, dtype), logprobas


def loss_and_grad_logits(labels, logits, label_length, logit_length, blank_index=0, variant=CLASSIC,
                         d_loss=None, dtype=np.float64):
    """loss [B] and d(sum_b d_loss[b] loss[b]) / d logits [B,T,V].

    Chain of base_loss.py:150-153 (forward_fn.backprop) with TF's autodiff of tools.py:37-39:
    dlogit = g - softmax * sum_k g, where g = d_loss * gradient.
    """
    data, logprobas = ctc_loss_data(labels, logits, label_length, logit_length, blank_index, variant, dtype)
    g = data.gradient
    if d_loss is not None:
        g = g * np.asarray(d_loss, dtype=dtype)[:, None, None]
    grad = g - np.exp(logprobas) * g.sum(axis=2, keepdims=True)
    return data.loss, grad, data


def hessian_logits(data: CtcLossData, logprobas: np.ndarray, hessian: np.ndarray | None = None) -> np.ndarray:
    """d2 loss / d logits2 [B,T,V,T,V] = what tape.batch_jacobian(gradient, logits) returns in the reference
    (README.md:58-71, tests/test_hessian.py:185-211): the data-class Hessian (wrt logprobas) pushed through
    the log-softmax Jacobian J = I - 1 p^T, plus the softmax curvature term.
    """
    H = data.hessian_fast() if hessian is None else hessian
    g = data.gradient
    p = np.exp(logprobas)
    s = g.sum(axis=2)                                             # [B,T]
    r1 = H.sum(axis=2)                                            # sum_k  H[t,k,t',j]   -> [B,T,T',V]
    r2 = H.sum(axis=4)                                            # sum_k' H[t,i,t',k']  -> [B,T,V,T']
    r12 = r1.sum(axis=3)                                          # [B,T,T']
    out = (H
           - p[:, :, :, None, None] * r1[:, :, None, :, :]
           - p[:, None, None, :, :] * r2[:, :, :, :, None]
           + p[:, :, :, None, None] * p[:, None, None, :, :] * r12[:, :, None, :, None])
    B, T, V = p.shape
    for t in range(T):
        blk = np.einsum("bi,ij->bij", p[:, t], np.eye(V)) - p[:, t, :, None] * p[:, t, None, :]
        out[:, t, :, t, :] -= s[:, t, None, None] * blk
    return out


def greedy_decode(logits: np.ndarray, logit_length, blank_index: int = 0, merge_repeated: bool = True):
    """Best-path decoding as tf.nn.ctc_greedy_decoder defines it (not part of tf_seq2seq_losses; checker for
    ctcb200_greedy_decode): per frame t < logit_length the arg-max token (np.argmax: lowest index on ties), repeats merged,
    blanks dropped.  Returns (decoded [B,T] int32 padded with -1, decoded_length [B], neg_sum_logits [B])."""
    logits = np.asarray(logits)
    B, T, _ = logits.shape
    decoded = np.full((B, T), -1, dtype=np.int32)
    length = np.zeros((B,), dtype=np.int32)
    neg_sum = np.zeros((B,), dtype=np.float64)
    for b in range(B):
        n = max(0, min(int(logit_length[b]), T))
        best = np.argmax(logits[b, :n], axis=1) if n else np.zeros((0,), dtype=np.int64)
        neg_sum[b] = -float(np.sum(np.max(logits[b, :n].astype(np.float64), axis=1))) if n else 0.0
        k, prev = 0, -1
        for t in range(n):
            tok = int(best[t])
            if tok != blank_index and not (merge_repeated and tok == prev):
                decoded[b, k] = tok
                k += 1
            prev = tok
        length[b] = k
    return decoded, length, neg_sum

