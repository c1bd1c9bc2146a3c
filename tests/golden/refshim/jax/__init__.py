"""torch.func-backed stand-in for the parts of JAX the reference hot path uses (see ../README.md)."""
import torch
from torch import func as _F
from . import numpy, ops, random, lax, flatten_util, scipy  # noqa: F401
from .numpy import _install_tensor_patches

_install_tensor_patches()


def jit(fun=None, static_argnums=None, static_argnames=None, **_kw):
    """Tracing compiler -> identity (eager)."""
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    if isinstance(in_axes, list):
        in_axes = tuple(in_axes)
    return _F.vmap(fun, in_dims=in_axes, out_dims=out_axes)


def grad(fun, argnums=0, has_aux=False):
    return _F.grad(fun, argnums=argnums, has_aux=has_aux)


def value_and_grad(fun, argnums=0, has_aux=False):
    return _F.grad_and_value(fun, argnums=argnums, has_aux=has_aux) if False else _vag(fun, argnums, has_aux)


def _vag(fun, argnums, has_aux):
    gv = _F.grad_and_value(fun, argnums=argnums, has_aux=has_aux)

    def wrapped(*a, **k):
        g, v = gv(*a, **k)
        return v, g
    return wrapped


def jacfwd(fun, argnums=0):
    return _F.jacfwd(fun, argnums=argnums)


def jacrev(fun, argnums=0):
    return _F.jacrev(fun, argnums=argnums)


def tree_map(f, tree, *rest):
    from torch.utils._pytree import tree_map as _tm
    return _tm(f, tree, *rest)
