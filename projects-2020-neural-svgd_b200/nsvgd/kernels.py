"""Drop-in for nvgd/src/kernels.py on the SVGD hot path: RBF kernel + median heuristic
(+ the funnel kernel used by the MMD tracer).  Other kernels of the reference file (tophat,
IMQ, ARD, ...) are not on the path and raise NotImplementedError when used with the engine."""
from __future__ import annotations

import math

import torch

from . import engine


def _check_xy(x, y, dim=None):
    """kernels.py:10-29 — same ValueErrors."""
    x, y = [torch.as_tensor(v) for v in (x, y)]
    if x.shape != y.shape:
        raise ValueError(f"Shapes of particles x and y need to match. "
                         f"Recieved shapes x: {tuple(x.shape)}, y: {tuple(y.shape)}")
    elif x.ndim > 1:
        raise ValueError(f"Input particles x and y can't have more than one "
                         f"dimension. Instead they have rank {x.ndim}")
    if dim is not None:
        if dim > 1:
            if x.ndim != 1 or x.shape[0] != dim:
                raise ValueError(f"x must have shape {(dim,)}. Instead received shape {tuple(x.shape)}.")
        elif dim == 1:
            if not (x.ndim == 0 or x.shape[0] == dim):
                raise ValueError(f"x must have shape (1,) or scalar. Instead received shape {tuple(x.shape)}.")
        else:
            raise ValueError("dim must be a natural nr")
    return x, y


def _check_bandwidth(bandwidth, dim=None):
    """kernels.py:31-42 — a scalar or a (d,) vector."""
    if isinstance(bandwidth, torch.Tensor):
        b = bandwidth.squeeze()
    else:
        b = torch.as_tensor(bandwidth, dtype=torch.float32).squeeze()
    if b.ndim > 1:
        raise ValueError(f"Bandwidth needs to be a scalar or a d-dim vector. Instead it has shape {tuple(b.shape)}")
    if dim is not None:
        if not (b.ndim == 0 or b.shape[0] in (0, 1, dim)):
            raise ValueError(f"Bandwidth has shape {tuple(b.shape)}.")
    return b


class RBFKernel:
    """Callable k(x, y) = exp(-sum((x-y)^2 / h^2) / 2) (kernels.py:52-65), tagged so that
    stein.get_phistar / ksd_squared* / metrics.get_mmd dispatch to the CUDA engine.
    `bandwidth` may be a python float or a 1-element device tensor (median heuristic)."""
    is_rbf = True
    funnel = False

    def __init__(self, bandwidth, squared=True, normalize=False, dim=None):
        self.bandwidth = _check_bandwidth(bandwidth, dim)
        self.squared, self.normalize, self.dim = squared, normalize, dim
        if normalize:
            raise NotImplementedError("normalize=True is not on the SVGD hot path")

    def is_ard(self):
        """a (d,) bandwidth vector (kernels.py:52-65 with `bandwidth` of shape (d,)): SURVEY 8f N3"""
        return self.bandwidth.ndim == 1 and self.bandwidth.numel() > 1

    def ard_scale(self, device):
        """per-coordinate length scales h_a: k(x, y) = exp(-sum_a (x_a - y_a)^2 / (2 h_a^2)).  The engine then runs on
        the rescaled cloud x' = x / h with unit bandwidth (stein.py / metrics.py do the three elementwise ops)."""
        b = self.bandwidth.to(device=device, dtype=torch.float32)
        return b if self.squared else torch.sqrt(b)

    def h(self):
        """effective bandwidth h with k = exp(-r^2/(2 h^2))"""
        if self.is_ard():
            raise NotImplementedError("this entry point needs a scalar bandwidth; vector (ARD) bandwidths are supported by "
                                      "stein.get_phistar / phistar / phistar_i and metrics.get_mmd (SURVEY 8f N3)")
        b = self.bandwidth.reshape(())
        return b if self.squared else torch.sqrt(b)

    def __call__(self, x, y):
        x, y = _check_xy(x, y, self.dim)
        h_squared = self.bandwidth ** 2 if self.squared else self.bandwidth
        h_squared = h_squared.to(x.device) if isinstance(h_squared, torch.Tensor) else h_squared
        return torch.exp(-torch.sum((x - y) ** 2 / h_squared) / 2)


def get_rbf_kernel(bandwidth, squared=True, normalize=False, dim=None):
    """kernels.py:52-65"""
    return RBFKernel(bandwidth, squared, normalize, dim)


def defunnelize(z):
    """kernels.py:112-116"""
    x, y = z[..., :-1], z[..., -1:]
    return torch.cat([x * torch.exp(-y / 2), y / 3], dim=-1)


def funnelize(v):
    """kernels.py:105-110"""
    x, y = v[..., :-1], v[..., -1:]
    return torch.cat([x * torch.exp(3 * y / 2), 3 * y], dim=-1)


class FunnelKernel(RBFKernel):
    """kernels.py:118-123: rbf(defunnelize(x), defunnelize(y))"""
    funnel = True

    def __call__(self, x, y):
        return super().__call__(defunnelize(x), defunnelize(y))


def get_funnel_kernel(bandwidth):
    return FunnelKernel(bandwidth)


class FullGaussianKernel:
    """kernels.py:83-88: k(x, y) = exp(-|Sigma (x - y)|^2 / 2) for a (d, d) matrix Sigma (the "full" parametrisation of
    nets.RBFKernel, nets.py:44-48).  With x' = Sigma x it is the unit-bandwidth RBF kernel, so phi* and MMD dispatch to the
    same engine after a linear map of particles and scores (stein.py / metrics.py); Sigma must be invertible for phi*."""
    is_rbf = True
    funnel = False
    is_full = True

    def __init__(self, sigma, dim=None):
        sigma = torch.as_tensor(sigma, dtype=torch.float32)
        if sigma.ndim != 2 or sigma.shape[0] != sigma.shape[1]:
            raise ValueError(f"sigma needs to be a (d, d) matrix. Instead it has shape {tuple(sigma.shape)}")
        if dim is not None and sigma.shape[0] != dim:
            raise ValueError(f"sigma has shape {tuple(sigma.shape)}.")
        self.sigma, self.dim = sigma, dim
        self.bandwidth = torch.tensor(1.0)

    def is_ard(self):
        return False

    def h(self):
        raise NotImplementedError("this entry point needs a scalar bandwidth; the full-matrix Gaussian kernel is supported by "
                                  "stein.get_phistar / phistar / phistar_i and metrics.get_mmd (SURVEY 8f N3)")

    def __call__(self, x, y):
        x, y = _check_xy(x, y, self.dim)
        s = self.sigma.to(x.device) @ (x - y)
        return torch.exp(-torch.sum(s * s) / 2)


def get_multivariate_gaussian_kernel(sigma, dim=None):
    """kernels.py:83-88"""
    return FullGaussianKernel(sigma, dim)


def get_rbf_kernel_logscaled(logh, normalize=False):
    """kernels.py:77-81 (the reference passes `normalize` positionally into `squared`: reproduced -> squared=normalize)"""
    return get_rbf_kernel(torch.exp(torch.as_tensor(logh, dtype=torch.float32)), normalize)


def median_heuristic(x, mode=None):
    """kernels.py:160-172.  Returns a 0-dim float32 DEVICE tensor (no host sync).
    mode None -> engine.DEFAULT_MEDIAN_MODE ("literal": exactly what the reference executes)."""
    return engine.median_heuristic(x, mode).reshape(())


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"kernels.{name} is outside the SVGD hot path covered by nsvgd (SURVEY.md 8)")
    f.__name__ = name
    return f


get_tophat_kernel = _unsupported("get_tophat_kernel")
get_imq_kernel = _unsupported("get_imq_kernel")
get_inverse_log_kernel = _unsupported("get_inverse_log_kernel")
