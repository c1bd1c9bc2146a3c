"""Differentiable forms of the hot-path operators: torch.autograd.Functions whose backward passes are the
closed-form CUDA rules of csrc/vjp.cu and csrc/pair_engine.cu (MODE_KSD with the d/dh accumulator) -- what a
jax.custom_vjp around the FFI ops would bind.  The reference differentiates through these operators with jax.grad /
jacfwd: the KSD w.r.t. the bandwidth (kernel learning, experiments/sd_maxing.py:71-75) and phi* w.r.t. its inputs
(stein.py:99-106, 180-203)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, ld, load, ptr, require_cuda_f32, stream_ptr, workspace, workspace_bytes


def _h_tensor(h, device):
    if isinstance(h, torch.Tensor):
        if h.numel() != 1:
            raise ValueError(f"Bandwidth needs to be a scalar here. Instead it has shape {tuple(h.shape)}")
        return h
    return torch.tensor(float(h), dtype=torch.float32, device=device)


class _KsdFn(torch.autograd.Function):
    """KSD^2 (stein.py:273-349) as a differentiable function of the bandwidth."""

    @staticmethod
    def forward(ctx, h, X, Sx, Y, Sy, variant):
        X = require_cuda_f32(X.detach(), "X").contiguous(); Sx = require_cuda_f32(Sx.detach(), "Sx").contiguous()
        n, d = X.shape
        m = n
        if variant == "rect":
            Y = require_cuda_f32(Y.detach(), "Y").contiguous(); Sy = require_cuda_f32(Sy.detach(), "Sy").contiguous()
            m = Y.shape[0]
        h_dev = h.detach().to(device=X.device, dtype=torch.float32).reshape(1).contiguous()
        out = torch.empty(1, dtype=torch.float64, device=X.device)
        dh = torch.empty(1, dtype=torch.float64, device=X.device)
        ws = workspace(workspace_bytes("ksd", n, m, d), X.device)
        check(load().nsvgd_ksd_dh_f32(stream_ptr(X.device), ptr(X), ptr(Sx), n, d, ptr(Y) if variant == "rect" else 0,
                                      ptr(Sy) if variant == "rect" else 0, m, d, d, ptr(h_dev), _lib.KSD_VARIANTS[variant],
                                      ptr(out), ptr(dh), ptr(ws), ws.numel()))
        ctx.save_for_backward(dh)
        ctx.h_meta = (h.shape, h.dtype, h.device)
        return out[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (dh,) = ctx.saved_tensors
        shape, dtype, device = ctx.h_meta
        return (g.to(torch.float64) * dh[0]).to(device=device, dtype=dtype).reshape(shape), None, None, None, None, None


def ksd_squared(X, Sx, h, variant="v", Y=None, Sy=None):
    """KSD^2 (variant "u" | "v" | "rect"), differentiable w.r.t. the bandwidth tensor h."""
    return _KsdFn.apply(_h_tensor(h, X.device), X, Sx, Y, Sy, variant)


class _PhiFn(torch.autograd.Function):
    """phi*(Q; X, S, h) (stein.py:180-222), differentiable w.r.t. Q, X, S and h."""

    @staticmethod
    def forward(ctx, Q, X, S, h):
        from . import engine
        Qc = require_cuda_f32(Q.detach(), "Q").contiguous(); Xc = require_cuda_f32(X.detach(), "X").contiguous()
        Sc = require_cuda_f32(S.detach(), "S").contiguous()
        h_dev = h.detach().to(device=Xc.device, dtype=torch.float32).reshape(1).contiguous()
        phi = engine.phistar(Qc, Xc, Sc, h_dev)
        ctx.save_for_backward(Qc, Xc, Sc, h_dev)
        ctx.h_meta = (h.shape, h.dtype, h.device)
        return phi

    @staticmethod
    def backward(ctx, G):
        Q, X, S, h_dev = ctx.saved_tensors
        G = require_cuda_f32(G, "cotangent").contiguous()
        m, d = Q.shape
        n = X.shape[0]
        need = [ctx.needs_input_grad[i] for i in range(4)]
        dQ = torch.empty_like(Q) if need[0] else None
        dX = torch.empty_like(X) if need[1] else None
        dS = torch.empty_like(S) if need[2] else None
        dh = torch.empty(1, dtype=torch.float64, device=X.device) if need[3] else None
        nbytes = C.c_size_t(0)
        check(load().nsvgd_phi_vjp_workspace_bytes(m, n, d, C.byref(nbytes)))
        ws = workspace(int(nbytes.value), X.device)
        check(load().nsvgd_phi_vjp_f32(stream_ptr(X.device), ptr(Q), ptr(G), m, ld(Q), ptr(X), ptr(S), n, ld(X), d, ptr(h_dev),
                                       ptr(dQ), d, ptr(dX), d, ptr(dS), d, ptr(dh), ptr(ws), ws.numel()))
        shape, dtype, device = ctx.h_meta
        gh = dh[0].to(device=device, dtype=dtype).reshape(shape) if need[3] else None
        return dQ, dX, dS, gh


def phistar(Q, X, S, h):
    """phi* at the rows of Q, differentiable w.r.t. every argument (pass the same tensor for Q and X in the SVGD case:
    autograd adds the two cotangents)."""
    return _PhiFn.apply(Q, X, S, _h_tensor(h, X.device))
