"""Test infrastructure: a numpy stand-in for the part of the TensorFlow API that tf_seq2seq_losses uses, so that the
REFERENCE'S OWN CODE (/root/reference/tf_seq2seq_losses/*.py, imported unmodified) can be executed in a container that
has no TensorFlow wheel.  ``install()`` registers this module as ``tensorflow`` in ``sys.modules``;
``tests/golden/make_reference_golden.py`` then imports the reference and writes its outputs as fixtures.

Every function follows the documented semantics of the TensorFlow op of the same name for the argument patterns the
reference uses (tf_seq2seq_losses/{tools,base_loss,classic_ctc_loss,simplified_ctc_loss}.py); nothing here knows about CTC.
Tensors are plain ``numpy.ndarray``s.  ``float32`` is the module-level float type: ``install(np.float64)`` -- what the
fixture generator uses -- runs the reference in double precision (``tf.float32`` then names float64, so the reference's
dtype assertion and casts follow), which allows comparisons at 1e-12.  (With ``install(np.float32)`` numpy's promotion
rules differ from TensorFlow's where the reference adds a numpy float64 scalar, tools.py:70, so that mode is not a
faithful float32 run.)

Never imported by the product package, by bench.py or by the GPU tests: only by the fixture generator."""
from __future__ import annotations

import builtins
import contextlib
import sys
import types

import numpy as np

class Tensor(np.ndarray):
    """A numpy array that also answers ``.numpy()`` (the reference's own tests call it); arithmetic, comparisons and slices
    of a Tensor stay Tensors, and an element access gives a 0-d Tensor instead of a numpy scalar."""

    def numpy(self):
        return np.asarray(self)

    def __getitem__(self, key):
        r = super().__getitem__(key)
        return r if isinstance(r, np.ndarray) else np.asarray(r).view(Tensor)


def _t(x):
    return np.asarray(x).view(Tensor)


int32 = np.dtype(np.int32)
int64 = np.dtype(np.int64)
bool = np.dtype(np.bool_)          # noqa: A001  (tf.bool)
float32 = np.dtype(np.float32)     # rebound by install()
inf = np.inf
_pybool = builtins.bool


class Variable:                    # only used in an isinstance() test (base_loss.py:121)
    pass


def _i(x):
    return int(np.asarray(x))


def _shape(shape):
    if isinstance(shape, np.ndarray):
        return tuple(int(v) for v in shape.reshape(-1))
    if isinstance(shape, (list, tuple)):
        return tuple(_i(v) for v in shape)
    return (_i(shape),)


def TensorShape(dims):             # noqa: N802
    return tuple(dims)


def constant(value, dtype=None):
    if dtype is None:
        a = np.asarray(value)
        if a.dtype.kind == "f":
            dtype = float32
        elif a.dtype.kind in "iu":
            dtype = int32
        else:
            dtype = a.dtype
    return np.asarray(value, dtype=dtype)


def convert_to_tensor(value, dtype=None):
    return value if isinstance(value, np.ndarray) and dtype is None else constant(value, dtype)


def cast(x, dtype):
    return np.asarray(x).astype(dtype)


def shape(input):                  # noqa: A002
    return np.asarray(np.shape(input), dtype=np.int32)


def reshape(tensor, shape):        # noqa: A002
    return np.reshape(tensor, _shape(shape))


def transpose(a, perm=None):
    return np.transpose(a, perm)


def expand_dims(input, axis):      # noqa: A002
    return np.expand_dims(input, axis)


def squeeze(input, axis=None):     # noqa: A002
    return np.squeeze(input, axis=axis)


def stack(values, axis=0):
    return np.stack([np.asarray(v) for v in values], axis=axis)


def concat(values, axis):
    return np.concatenate([np.asarray(v) for v in values], axis=axis)


def tile(input, multiples):        # noqa: A002
    return np.tile(input, _shape(multiples))


def roll(input, shift, axis):      # noqa: A002
    return np.roll(input, shift, axis=axis)


def pad(tensor, paddings, constant_values=0):
    widths = [(_i(lo), _i(hi)) for lo, hi in paddings]
    return np.pad(tensor, widths, mode="constant", constant_values=np.asarray(constant_values).astype(np.asarray(tensor).dtype))


def zeros(shape, dtype=None):      # noqa: A002
    return np.zeros(_shape(shape), dtype=float32 if dtype is None else dtype)


def ones(shape, dtype=None):       # noqa: A002
    return np.ones(_shape(shape), dtype=float32 if dtype is None else dtype)


def zeros_like(input):             # noqa: A002
    return np.zeros_like(input)


def ones_like(input):              # noqa: A002
    return np.ones_like(input)


def eye(num_rows, dtype=None):
    return np.eye(_i(num_rows), dtype=float32 if dtype is None else dtype)


def range(limit):                  # noqa: A001
    return np.arange(_i(limit), dtype=np.int32)


def one_hot(indices, depth, dtype=None):
    idx = np.asarray(indices)
    out = (idx[..., None] == np.arange(_i(depth))).astype(float32 if dtype is None else dtype)
    return out


def sequence_mask(lengths, maxlen=None):
    lengths = np.asarray(lengths)
    n = _i(maxlen) if maxlen is not None else int(lengths.max())
    return np.arange(n) < lengths[..., None]


def where(condition, x=None, y=None):
    return np.where(condition, x, y)


def cond(pred, true_fn, false_fn):
    return true_fn() if _pybool(np.asarray(pred)) else false_fn()


def exp(x):
    with np.errstate(all="ignore"):
        return np.exp(x)


def reduce_sum(input_tensor, axis=None, keepdims=False):
    return np.sum(input_tensor, axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims)


def _axis(axis):
    return tuple(axis) if isinstance(axis, list) else axis


def reduce_max(input_tensor, axis=None, keepdims=False):
    return np.max(input_tensor, axis=_axis(axis), keepdims=keepdims)


def reduce_mean(input_tensor, axis=None, keepdims=False):
    return np.mean(input_tensor, axis=_axis(axis), keepdims=keepdims)


def reduce_all(input_tensor, axis=None, keepdims=False):
    return np.all(input_tensor, axis=_axis(axis), keepdims=keepdims)


def abs(x):                        # noqa: A001
    return np.abs(x)


def norm(tensor, ord="euclidean", axis=None):   # noqa: A002
    x = np.asarray(tensor)
    if ord == np.inf:
        return np.max(np.abs(x), axis=_axis(axis))
    assert ord in ("euclidean", 2)
    return np.sqrt(np.sum(x * x, axis=_axis(axis)))


def reduce_logsumexp(input_tensor, axis=None, keepdims=False):
    """tf.reduce_logsumexp: log(sum(exp(x - m))) + m with m = max where that is finite, else 0."""
    x = np.asarray(input_tensor)
    axis = _axis(axis)
    with np.errstate(all="ignore"):
        raw = np.max(x, axis=axis, keepdims=True)
        m = np.where(np.isfinite(raw), raw, np.zeros_like(raw))
        out = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
    return out if keepdims else np.squeeze(out, axis=axis)


def cumsum(x, axis=0, exclusive=False):
    c = np.cumsum(x, axis=axis)
    return c - x if exclusive else c


def meshgrid(*args):
    return np.meshgrid(*args)


def gather(params, indices, axis=None, batch_dims=0):
    params, indices = np.asarray(params), np.asarray(indices)
    axis = batch_dims if axis is None else _i(axis)       # tf.gather: axis defaults to batch_dims
    if batch_dims == 0:
        return np.take(params, indices, axis=axis)
    assert batch_dims == 1 and axis >= 1 and params.shape[0] == indices.shape[0]
    return np.stack([np.take(params[b], indices[b], axis=axis - 1) for b in np.arange(params.shape[0])], axis=0)


def scatter_nd(indices, updates, shape):   # noqa: A002
    out = np.zeros(_shape(shape), dtype=np.asarray(updates).dtype)
    np.add.at(out, tuple(np.asarray(indices).T), updates)
    return out


def stop_gradient(input):          # noqa: A002
    return input


def custom_gradient(f):
    """The forward value only; the backward closure is kept on the wrapper for inspection."""
    def wrapper(*args, **kwargs):
        value, grad_fn = f(*args, **kwargs)
        wrapper.last_grad_fn = grad_fn
        return value
    wrapper.__wrapped__ = f
    return wrapper


@contextlib.contextmanager
def name_scope(name):
    yield name


class TensorArray:
    """tf.TensorArray as the reference uses it (tools.py:221-229): fixed size, write / read / stack."""

    def __init__(self, dtype, size, element_shape=None, clear_after_read=False, infer_shape=True, dynamic_size=False):
        self._items = [None] * _i(size)

    def write(self, index, value):
        self._items[_i(index)] = np.asarray(value)
        return self

    def read(self, index):
        return _t(self._items[_i(index)])

    def stack(self):
        return _t(np.stack(self._items, axis=0))


def while_loop(cond, body, loop_vars, maximum_iterations=None, swap_memory=False, name=None):   # noqa: A002
    state = tuple(loop_vars)
    n = 0
    while (maximum_iterations is None or n < _i(maximum_iterations)) and _pybool(np.asarray(cond(*state))):
        state = tuple(body(*state))
        n += 1
    return state


def _log(x):
    with np.errstate(all="ignore"):
        return np.log(x)


def _softplus(x):
    with np.errstate(all="ignore"):
        return np.logaddexp(np.zeros_like(x), x)


def _expm1(x):
    with np.errstate(all="ignore"):
        return np.expm1(x)


def _segment_reduce(data, segment_ids, num_segments, ufunc, init):
    data, ids = np.asarray(data), np.asarray(segment_ids)
    out = np.full((_i(num_segments),) + data.shape[ids.ndim:], init, dtype=data.dtype)
    ufunc.at(out, ids.reshape(-1), data.reshape((-1,) + data.shape[ids.ndim:]))
    return out


def _unsorted_segment_max(data, segment_ids, num_segments):
    """Empty segments (and segments holding only -inf) give the lowest finite value of the type, as TensorFlow does."""
    return _segment_reduce(data, segment_ids, num_segments, np.maximum, np.finfo(np.asarray(data).dtype).min)


def _unsorted_segment_sum(data, segment_ids, num_segments):
    return _segment_reduce(data, segment_ids, num_segments, np.add, 0)


def _band_part(input, num_lower, num_upper):   # noqa: A002
    x = np.asarray(input)
    m, n = x.shape[-2:]
    i, j = np.arange(m)[:, None], np.arange(n)[None, :]
    keep = ((num_lower < 0) | (i - j <= num_lower)) & ((num_upper < 0) | (j - i <= num_upper))
    return np.where(keep, x, np.zeros_like(x))


def _set_diag(input, diagonal):    # noqa: A002
    out = np.array(input, copy=True)
    n = min(out.shape[-2:])
    idx = np.arange(n)
    out[..., idx, idx] = diagonal
    return out


# ---- only the reference's own TESTS need what follows (tests/golden/run_reference_tests.py) ----
_rng = np.random.default_rng(0)


def _set_seed(seed):
    global _rng
    _rng = np.random.default_rng(seed)


def _random_normal(shape, mean=0.0, stddev=1.0, dtype=None):   # noqa: A002
    return (_rng.standard_normal(_shape(shape)) * stddev + mean).astype(float32 if dtype is None else dtype)


def _random_uniform(shape, minval=0, maxval=None, dtype=None):   # noqa: A002
    dtype = np.dtype(float32 if dtype is None else dtype)
    if dtype.kind in "iu":
        return _rng.integers(minval, maxval, size=_shape(shape)).astype(dtype)
    return (_rng.random(_shape(shape)) * ((1.0 if maxval is None else maxval) - minval) + minval).astype(dtype)


def function(func=None, **kwargs):     # tf.function: eager execution is all there is
    return func if func is not None else (lambda f: f)


def TensorSpec(shape=None, dtype=None, name=None):   # noqa: N802, A002
    return (shape, dtype)


math = types.SimpleNamespace(log=_log, softplus=_softplus, expm1=_expm1, exp=exp, unsorted_segment_max=_unsorted_segment_max,
                             unsorted_segment_sum=_unsorted_segment_sum)
linalg = types.SimpleNamespace(band_part=_band_part, set_diag=_set_diag)


random = types.SimpleNamespace(set_seed=_set_seed, normal=_random_normal, uniform=_random_uniform)


def _as_tensor_result(f):
    """Results leave the shim as Tensor views (so that ``.numpy()`` works on them); the values are untouched."""
    def g(*args, **kwargs):
        r = f(*args, **kwargs)
        if isinstance(r, (np.ndarray, np.generic)):
            return _t(r)
        if isinstance(r, list):
            return [_t(v) for v in r]
        return r
    g.__name__ = getattr(f, "__name__", "op")
    g.__doc__ = f.__doc__
    return g


for _name in ("constant", "convert_to_tensor", "cast", "shape", "reshape", "transpose", "expand_dims", "squeeze", "stack",
              "concat", "tile", "roll", "pad", "zeros", "ones", "zeros_like", "ones_like", "eye", "range", "one_hot",
              "sequence_mask", "where", "exp", "reduce_sum", "reduce_max", "reduce_mean", "reduce_all", "reduce_logsumexp",
              "abs", "norm", "cumsum", "meshgrid", "gather", "scatter_nd"):
    globals()[_name] = _as_tensor_result(globals()[_name])
for _ns in (math, linalg, random):
    for _name, _f in list(vars(_ns).items()):
        if callable(_f) and _name != "set_seed":
            setattr(_ns, _name, _as_tensor_result(_f))


def install(float_dtype=np.float32):
    """Registers this module as ``tensorflow`` and sets the float type that ``tf.float32`` names."""
    global float32
    float32 = np.dtype(float_dtype)
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    for name in [n for n in sys.modules if n == "tf_seq2seq_losses" or n.startswith("tf_seq2seq_losses.")]:
        del sys.modules[name]       # module-level constants of the reference (tools.inf) are typed at import time
    return me
