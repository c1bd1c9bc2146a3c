"""Drop-in for the three hot-path helpers of nvgd/src/utils.py plus the function combinators
KernelGradient uses (negative, mul) and a `vmap` that folds a per-point phi* closure into ONE
engine call (the role jax.vmap + the FFI batching rule play in the reference)."""
from __future__ import annotations

import torch


class BatchedFunction:
    """A function of one point (d,) that also knows how to evaluate a whole (m,d) batch in one
    engine call.  `scale` implements utils.negative / utils.mul without another pass."""

    def __init__(self, batch_fn, scale=1.0, base=None):
        self._batch_fn = batch_fn
        self.scale = scale
        self.base = base if base is not None else self

    def batch(self, xs):
        out = self._batch_fn(xs)
        if isinstance(self.scale, torch.Tensor) or self.scale != 1.0:
            out = out * (self.scale.to(out.dtype) if isinstance(self.scale, torch.Tensor) else self.scale)
        return out

    def __call__(self, x):
        if x.ndim != 1:
            raise ValueError(f"Shape of xi must be (d,). Instead, received shape {tuple(x.shape)}")
        return self.batch(x[None, :])[0]

    def scaled(self, factor):
        s = factor * self.scale if not isinstance(factor, torch.Tensor) else factor * self.scale
        return BatchedFunction(self._batch_fn, s, self.base)


def vmap(fun, in_axes=0):
    """vmap over the leading axis.  BatchedFunction -> one engine call; anything else -> torch.func.vmap."""
    if isinstance(fun, BatchedFunction):
        return fun.batch
    return torch.func.vmap(fun, in_dims=in_axes)


def negative(fun):
    """f --> -f                                                        utils.py:515-519"""
    if isinstance(fun, BatchedFunction):
        return fun.scaled(-1.0)

    def negfun(*args, **kwargs):
        return -fun(*args, **kwargs)
    return negfun


def mul(fun, factor):
    """f --> factor * fun                                              utils.py:542-544"""
    if isinstance(fun, BatchedFunction):
        return fun.scaled(factor)
    return lambda *args, **kwargs: factor * fun(*args, **kwargs)


def l2_norm_squared(samples, fun):
    """mean of fun^T fun over samples                                  utils.py:483-487"""
    vals = vmap(fun)(samples)
    return torch.mean(torch.sum(vals * vals, dim=1))


def remove_diagonal(matrix):
    """(n,n) -> (n,n-1) without the diagonal                           utils.py:653-658"""
    n = matrix.shape[0]
    mask = ~torch.eye(n, dtype=torch.bool, device=matrix.device)
    return matrix[mask].reshape(n, n - 1)


def squared_distance_matrix(x):
    """utils.py:178-197.  Convenience only (materialises n x n); the engine's median never does."""
    if x.ndim == 1:
        x = x[:, None]
    assert x.ndim == 2
    diff = x[:, None, :] - x[None, :, :]
    return torch.sum(diff * diff, dim=-1)
