"""Minimal targets/proposals for the harness (the reference's distributions.py is out of scope;
it only has to hand the hot path a `logpdf`).  Each logpdf object is a callable on (d,) tensors
(usable with torch.func.grad) and also exposes batched analytic `.score(X)` / `.batch_logp(X)`."""
from __future__ import annotations

import math

import torch


class _LogPdf:
    def __init__(self, logp, score, batch_logp):
        self._logp, self.score, self.batch_logp = logp, score, batch_logp

    def __call__(self, x):
        return self._logp(x)


class Funnel:
    """Neal's funnel (distributions.py:275-319): y ~ N(0, 9), x_k | y ~ N(0, e^y)."""

    def __init__(self, d):
        self.d = d

        def logp(x):
            xs, y = x[:-1], x[-1]
            return -y ** 2 / 18.0 - 0.5 * (d - 1) * y - 0.5 * torch.exp(-y) * torch.sum(xs ** 2)

        def score(X):
            xs, y = X[:, :-1], X[:, -1:]
            gx = -xs * torch.exp(-y)
            gy = -y / 9.0 - (d - 1) / 2.0 + 0.5 * torch.exp(-y) * torch.sum(xs * xs, dim=1, keepdim=True)
            return torch.cat([gx, gy], dim=1)

        def batch_logp(X):
            xs, y = X[:, :-1], X[:, -1]
            return -y ** 2 / 18.0 - 0.5 * (d - 1) * y - 0.5 * torch.exp(-y) * torch.sum(xs ** 2, dim=1)

        self.logpdf = _LogPdf(logp, score, batch_logp)

    def sample(self, n_samples, key=None):
        y = torch.randn(n_samples, 1, generator=key) * 3
        x = torch.randn(n_samples, self.d - 1, generator=key) * torch.exp(y / 2)
        return torch.cat([x, y], dim=1)


class Gaussian:
    """N(mean, cov) with cov a (d,) diagonal or a (d,d) matrix (distributions.py:89-154)."""

    def __init__(self, mean, cov):
        self.mean = torch.as_tensor(mean, dtype=torch.float32).reshape(-1)
        self.d = self.mean.numel()
        cov = torch.as_tensor(cov, dtype=torch.float32)
        if cov.ndim == 0:
            cov = cov * torch.ones(self.d)
        self.cov = torch.diag(cov) if cov.ndim == 1 else cov
        self.prec = torch.linalg.inv(self.cov.double()).float()
        self.chol = torch.linalg.cholesky(self.cov.double()).float()
        mean_, prec_ = self.mean, self.prec

        def logp(x):
            c = x - mean_.to(x.device)
            return -0.5 * torch.dot(c, prec_.to(x.device) @ c)

        def score(X):
            return -(X - mean_.to(X.device)) @ prec_.to(X.device).T

        def batch_logp(X):
            c = X - mean_.to(X.device)
            return -0.5 * torch.sum(c * (c @ prec_.to(X.device).T), dim=1)

        self.logpdf = _LogPdf(logp, score, batch_logp)

    def sample(self, n_samples, key=None):
        z = torch.randn(n_samples, self.d, generator=key)
        return z @ self.chol.T + self.mean


class Setup:
    """distributions.py:544-578: a (target, proposal) pair."""

    def __init__(self, target, proposal):
        self.target, self.proposal = target, proposal

    def get(self):
        return self.target, self.proposal


funnel = Setup(Funnel(2), Gaussian([-3.0, 0.0], [9.0, 9.0]))      # distributions.py:594-596
