"""Drop-in for the learnable kernel of nvgd/src/nets.py:24-49 (`RBFKernel`, a haiku module there, a torch.nn.Module here).
Three parametrisations, equivalent at initialisation (nets.py:27-33):
  "diagonal"      one bandwidth per dimension, init 1                       -> kernels.get_rbf_kernel(bandwidth)
  "log_diagonal"  log-bandwidths, init 0, clipped to [-5, 5]                -> kernels.get_rbf_kernel_logscaled(log_bandwidth)
  "full"          a (d, d) matrix, init identity                            -> kernels.get_multivariate_gaussian_kernel(sigma)
`module(xy)` evaluates scale * k(x, y) for xy of shape (2, d) exactly like the reference's __call__; `module.kernel()`
returns the TAGGED kernel object that stein.get_phistar / phistar / ksd_squared* / metrics.get_mmd dispatch to the CUDA
engine.  The rest of nets.py (MLPs, DeepKernel, CNNs: the neural learner) is outside the hot path."""
from __future__ import annotations

import torch

from . import kernels


class RBFKernel(torch.nn.Module):
    def __init__(self, scale_param=False, parametrization="diagonal", dim=None, name=None):
        super().__init__()
        if parametrization not in ("diagonal", "log_diagonal", "full"):
            raise ValueError(f"unknown parametrization {parametrization!r}")
        self.parametrization, self.scale_param, self.name = parametrization, scale_param, name
        self.scale = torch.nn.Parameter(torch.ones(())) if scale_param else None
        self._built = False
        if dim is not None:
            self._build(int(dim))

    def _build(self, d):
        """haiku creates the parameters at the first call, when d is known (nets.py:38-48)"""
        if self.parametrization == "log_diagonal":
            self.log_bandwidth = torch.nn.Parameter(torch.zeros(d))
        elif self.parametrization == "diagonal":
            self.bandwidth = torch.nn.Parameter(torch.ones(d))
        else:
            self.sigma = torch.nn.Parameter(torch.eye(d))          # bandwidth_init, nets.py:9-12
        self._built = True

    def kernel(self, dim=None):
        """the tagged kernel for the current parameter values (detached: the engine evaluates it; gradients w.r.t. the
        parameters flow through nsvgd.autodiff when requested explicitly)"""
        if not self._built:
            if dim is None:
                raise ValueError("RBFKernel parameters are created at the first call; pass dim")
            self._build(int(dim))
        if self.parametrization == "log_diagonal":
            return kernels.get_rbf_kernel_logscaled(torch.clamp(self.log_bandwidth.detach(), -5, 5))
        if self.parametrization == "diagonal":
            return kernels.get_rbf_kernel(self.bandwidth.detach())
        return kernels.get_multivariate_gaussian_kernel(self.sigma.detach())

    def forward(self, xy):
        """xy should have shape (2, d)                                            nets.py:35-48"""
        d = xy.shape[-1]
        if not self._built:
            self._build(d)
            self.to(xy.device)
        scale = self.scale if self.scale_param else 1.
        x, y = xy[0], xy[1]
        if self.parametrization == "log_diagonal":
            h = torch.exp(torch.clamp(self.log_bandwidth, -5, 5))
            # get_rbf_kernel_logscaled passes normalize=False positionally into `squared` (kernels.py:77-81): h is then
            # used UNSQUARED, k = exp(-sum((x-y)^2 / h) / 2)
            return scale * torch.exp(-torch.sum((x - y) ** 2 / h) / 2)
        if self.parametrization == "diagonal":
            return scale * torch.exp(-torch.sum((x - y) ** 2 / self.bandwidth ** 2) / 2)
        s = self.sigma @ (x - y)
        return scale * torch.exp(-torch.sum(s * s) / 2)
