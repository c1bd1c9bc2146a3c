"""jax.numpy stand-in on CPU torch tensors (see ../README.md).  Only what the reference hot path calls.

JAX semantics kept on purpose: float32 default dtype, Python scalars are weakly typed, gathers clamp
out-of-bounds indices, `median` is the linear-interpolation 0.5-quantile, `.transpose()` is numpy's.
Set REFSHIM_WIDEN=1 to alias float32 -> float64 (dtype widened, source otherwise unmodified): used to pin the
fp64 closed forms of oracle/ to ~1e-12 instead of fp32 round-off."""
import builtins as _b
import math
import os

import numpy as _onp
import torch

WIDEN = os.environ.get("REFSHIM_WIDEN", "0") == "1"
float32 = torch.float64 if WIDEN else torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool_ = torch.bool
torch.set_default_dtype(float32)
ndarray = torch.Tensor
pi = math.pi
inf = math.inf
newaxis = None


def _t(x, dtype=None):
    """to tensor with JAX-like dtype defaults (python float -> default float dtype, python int -> int)."""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (list, tuple)):
        if len(x) and _b.any(isinstance(e, torch.Tensor) for e in x):
            out = torch.stack([_t(e) for e in x])
            return out if dtype is None else out.to(dtype)
        x = _onp.asarray(x)
    if isinstance(x, _onp.ndarray):
        if x.dtype == _onp.float64 or x.dtype == _onp.float32:
            return torch.as_tensor(x).to(dtype or float32)
        return torch.as_tensor(x) if dtype is None else torch.as_tensor(x).to(dtype)
    if isinstance(x, bool):
        return torch.tensor(x, dtype=dtype or torch.bool)
    if isinstance(x, int):
        return torch.tensor(x, dtype=dtype or torch.int64)
    if isinstance(x, float):
        return torch.tensor(x, dtype=dtype or float32)
    return torch.as_tensor(x)


def _f(x):
    """operand of a float function: ints become the default float dtype (JAX promotes int -> float32)."""
    x = _t(x)
    return x if x.is_floating_point() else x.to(float32)


def array(x, dtype=None, copy=True):
    return _t(x, dtype)


asarray = array


def _axis(kw, axis):
    return {} if axis is None else {"dim": axis}


def exp(x): return torch.exp(_f(x))
def log(x): return torch.log(_f(x))
def sqrt(x): return torch.sqrt(_f(x))
def abs(x): return torch.abs(_t(x))
def square(x): return torch.square(_t(x))
def tanh(x): return torch.tanh(_f(x))


def sum(x, axis=None, keepdims=False):
    x = _t(x)
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis, keepdim=keepdims)


def mean(x, axis=None, keepdims=False):
    x = _f(x)
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis, keepdim=keepdims)


def std(x, axis=None, ddof=0):
    x = _f(x)
    return torch.std(x, correction=ddof) if axis is None else torch.std(x, dim=axis, correction=ddof)


def var(x, axis=None, ddof=0):
    x = _f(x)
    return torch.var(x, correction=ddof) if axis is None else torch.var(x, dim=axis, correction=ddof)


def prod(x, axis=None):
    x = _t(x)
    return torch.prod(x) if axis is None else torch.prod(x, dim=axis)


def all(x, axis=None):
    x = _t(x)
    return torch.all(x) if axis is None else torch.all(x, dim=axis)


def any(x, axis=None):
    x = _t(x)
    return torch.any(x) if axis is None else torch.any(x, dim=axis)


def where(c, a, b):
    a, b = _t(a), _t(b)
    return torch.where(_t(c), a, b)


def inner(a, b):
    a, b = _t(a), _t(b)
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    return torch.sum(a * b, dim=-1) if a.ndim == 1 and b.ndim == 1 else torch.inner(a, b)


def dot(a, b):
    a, b = _t(a), _t(b)
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    return torch.matmul(a, b)


matmul = dot


def einsum(spec, *ops):
    return torch.einsum(spec, *[_t(o) for o in ops])


def trace(x):
    return torch.diagonal(_t(x), dim1=0, dim2=1).sum(-1)


def diag(x):
    x = _t(x)
    return torch.diagonal(x) if x.ndim == 2 else torch.diag(x)


def transpose(x, axes=None):
    return _t(x).transpose() if axes is None else _t(x).permute(*axes)


def reshape(x, newshape):
    if isinstance(newshape, int):
        newshape = (newshape,)
    return _t(x).reshape(tuple(newshape))


def squeeze(x, axis=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def tile(x, reps):
    x = _t(x)
    if isinstance(reps, int):
        reps = (reps,)
    return x.repeat(*reps) if len(reps) >= x.ndim else x.repeat(*((1,) * (x.ndim - len(reps)) + tuple(reps)))


def repeat(x, repeats, axis=None):
    x = _t(x)
    repeats = _t(repeats) if not isinstance(repeats, int) else repeats
    return torch.repeat_interleave(x if axis is not None else x.reshape(-1), repeats, dim=axis if axis is not None else 0)


def concatenate(xs, axis=0):
    return torch.cat([_t(x) for x in xs], dim=axis)


def stack(xs, axis=0):
    return torch.stack([_t(x) for x in xs], dim=axis)


def append(a, b, axis=None):
    a, b = _f(a), _f(b)
    if axis is None:
        return torch.cat([a.reshape(-1), b.reshape(-1)])
    return torch.cat([a, b], dim=axis)


def ones(shape, dtype=None):
    return torch.ones(shape if not isinstance(shape, int) else (shape,), dtype=dtype or float32)


def zeros(shape, dtype=None):
    return torch.zeros(shape if not isinstance(shape, int) else (shape,), dtype=dtype or float32)


def ones_like(x): return torch.ones_like(_t(x))
def zeros_like(x): return torch.zeros_like(_t(x))
def eye(n, dtype=None): return torch.eye(n, dtype=dtype or float32)
def arange(*a, dtype=None): return torch.arange(*a, dtype=dtype)
def linspace(a, b, n): return torch.linspace(a, b, n, dtype=float32)
def clip(x, a_min=None, a_max=None): return torch.clamp(_t(x), a_min, a_max) if isinstance(x, torch.Tensor) else max(x, a_min) if a_max is None else min(max(x, a_min), a_max)
def maximum(a, b): return torch.maximum(_t(a), _t(b))
def minimum(a, b): return torch.minimum(_t(a), _t(b))
def isnan(x): return torch.isnan(_t(x))


def median(x, axis=None):
    """jnp.median = quantile(x, 0.5) with linear interpolation: for an even count
    low*0.5 + high*0.5 (jax/_src/numpy/lax_numpy.py `_quantile`); NaN if any NaN."""
    x = _f(x)
    assert axis is None
    v, _ = torch.sort(x.reshape(-1))
    n = v.numel()
    if bool(torch.isnan(v).any()):
        return torch.tensor(float("nan"), dtype=v.dtype)
    q = 0.5 * (n - 1)
    lo, hi = int(math.floor(q)), int(math.ceil(q))
    hi_w = q - lo
    lo_w = 1.0 - hi_w
    return v[lo] * lo_w + v[hi] * hi_w


class _Linalg:
    @staticmethod
    def norm(x, axis=None):
        x = _f(x)
        return torch.linalg.norm(x) if axis is None else torch.linalg.norm(x, dim=axis)


linalg = _Linalg()


# ---------------------------------------------------------------------------------------------------------------
# numpy/JAX method semantics on torch.Tensor (process-wide; the golden generator is its own process)
# ---------------------------------------------------------------------------------------------------------------
_installed = False


def _install_tensor_patches():
    global _installed
    if _installed:
        return
    _installed = True
    _orig_transpose = torch.Tensor.transpose
    _orig_getitem = torch.Tensor.__getitem__

    def np_transpose(self, *axes):
        if len(axes) == 0:
            return self.permute(*reversed(range(self.ndim)))
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            return self.permute(*axes[0])
        return _orig_transpose(self, *axes)

    def jax_getitem(self, idx):
        # JAX gather semantics for integer-array indices: negative wraps once, out-of-bounds is CLAMPED
        # ("promise_in_bounds"/clip default of x[idx]); utils.remove_diagonal relies on this for (n,d) input, d<n-1.
        if isinstance(idx, tuple) and len(idx) and _b.all(isinstance(i, _onp.ndarray) and i.dtype.kind in "iu" for i in idx):
            fixed = []
            for ax, i in enumerate(idx):
                size = self.shape[ax]
                i = _onp.where(i < 0, i + size, i)
                fixed.append(torch.as_tensor(_onp.clip(i, 0, size - 1)))
            return _orig_getitem(self, tuple(fixed))
        return _orig_getitem(self, idx)

    torch.Tensor.transpose = np_transpose
    torch.Tensor.__getitem__ = jax_getitem

