"""Minimal targets/proposals for the harness (the reference's distributions.py is out of scope;
it only has to hand the hot path a `logpdf`).  Each logpdf object is a callable on (d,) tensors
(usable with torch.func.grad) and also exposes batched analytic `.score(X)` / `.batch_logp(X)`.
On CUDA tensors `.score` / `.batch_logp` run the device kernels of csrc/scores.cu: the step
scores -> bandwidth -> phi* -> update is then enqueued without autodiff; CPU tensors (harness set-up, sampling)
use the same closed forms in torch."""
from __future__ import annotations

import math

import torch


class _LogPdf:
    def __init__(self, logp, score, batch_logp, device_spec=None):
        self._logp, self.score, self.batch_logp = logp, score, batch_logp
        # (target id, parameter tensor | None, diagonal precision?, log normaliser): lets the persistent small-n loop
        # (csrc/small_loop.cu) evaluate the scores inside the kernel
        self.device_spec = device_spec

    def __call__(self, x):
        return self._logp(x)


class Funnel:
    """Neal's funnel (distributions.py:275-319): y ~ N(0, 9), x_k | y ~ N(0, e^y)."""

    def __init__(self, d):
        self.d = d

        # normalised like the reference (stats.norm.logpdf(y, 0, 3) + multivariate_normal.logpdf(x, 0, e^y I), :311-319)
        lognorm = -math.log(3.0) - 0.5 * d * math.log(2.0 * math.pi)

        def logp(x):
            xs, y = x[:-1], x[-1]
            return -y ** 2 / 18.0 - 0.5 * (d - 1) * y - 0.5 * torch.exp(-y) * torch.sum(xs ** 2) + lognorm

        def score(X):
            if X.is_cuda:
                from . import engine
                return engine.score_funnel(X)
            xs, y = X[:, :-1], X[:, -1:]
            gx = -xs * torch.exp(-y)
            gy = -y / 9.0 - (d - 1) / 2.0 + 0.5 * torch.exp(-y) * torch.sum(xs * xs, dim=1, keepdim=True)
            return torch.cat([gx, gy], dim=1)

        def batch_logp(X):
            xs, y = X[:, :-1], X[:, -1]
            return -y ** 2 / 18.0 - 0.5 * (d - 1) * y - 0.5 * torch.exp(-y) * torch.sum(xs ** 2, dim=1) + lognorm

        self.logpdf = _LogPdf(logp, score, batch_logp, device_spec=(0, None, 0, lognorm))

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
        # normalised like stats.multivariate_normal.logpdf (:148-150)
        lognorm = float(-0.5 * (self.d * math.log(2.0 * math.pi) + torch.logdet(self.cov.double())))

        def logp(x):
            c = x - mean_.to(x.device)
            return -0.5 * torch.dot(c, prec_.to(x.device) @ c) + lognorm

        dev_cache = {}

        def score(X):
            if X.is_cuda:
                from . import engine
                if X.device not in dev_cache:
                    dev_cache[X.device] = (mean_.to(X.device), prec_.to(X.device).contiguous())
                m, p_ = dev_cache[X.device]
                return engine.score_gaussian(X, m, p_)
            return -(X - mean_.to(X.device)) @ prec_.to(X.device).T

        def batch_logp(X):
            c = X - mean_.to(X.device)
            return -0.5 * torch.sum(c * (c @ prec_.to(X.device).T), dim=1) + lognorm

        self.logpdf = _LogPdf(logp, score, batch_logp,
                              device_spec=(1, torch.cat([mean_.reshape(-1), prec_.reshape(-1)]).contiguous(), 0, lognorm))

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



class LogisticRegressionPosterior:
    """The target of experiments/bayesian_logistic_regression.py: params = [w (F features incl. intercept), log alpha],
    alpha ~ Gamma(a0, rate b0), w | alpha ~ N(0, 1/alpha), y | x, w ~ Bernoulli(sigmoid(x.w))   (:31, :73-114).
    `get_minibatch_logp(x, y)` mirrors :169-190 -- including its literal scaling num_datapoints * SUM over the batch --
    and returns a logp object that is callable on one (d,) parameter vector (torch.func.grad works on it) and exposes
    the batched device-side `.score(P)` / `.batch_logp(P)` (nsvgd_score_logreg_f32)."""

    def __init__(self, num_features, num_datapoints, a0=1.0, b0=0.01):
        self.F, self.N, self.a0, self.b0 = int(num_features), float(num_datapoints), float(a0), float(b0)
        self.d = self.F + 1

    def sample_from_prior(self, num, key=None):
        """ravel(*sample_from_prior) (:66-71, :209): alpha ~ Gamma(a0) / b0, w ~ N(0, I)"""
        g = torch.distributions.Gamma(torch.tensor(self.a0), torch.tensor(1.0))
        if key is not None:
            torch.manual_seed(int(key) if not isinstance(key, torch.Generator) else key.initial_seed())
        alpha = g.sample((num,)) / self.b0
        w = torch.randn(num, self.F)
        return torch.cat([w, torch.log(alpha)[:, None]], dim=1)

    def get_minibatch_logp(self, x, y):
        F, N, a0, b0 = self.F, self.N, self.a0, self.b0
        x = torch.as_tensor(x, dtype=torch.float32)
        y = torch.as_tensor(y, dtype=torch.float32)
        if x.shape[0] != y.shape[0] or x.ndim <= y.ndim:
            raise AssertionError("x, y must be minibatches of the same length")       # the asserts of :176-177

        def logp(params):
            w, la = params[:-1], params[-1]
            alpha = torch.exp(la)
            lg = (a0 - 1.0) * la - b0 * alpha + a0 * math.log(b0) - math.lgamma(a0)
            lw = torch.sum(-0.5 * alpha * w * w + 0.5 * la - 0.5 * math.log(2 * math.pi))
            ys = (y.to(params.device) - 0.5) * 2.0
            ll = torch.sum(torch.nn.functional.logsigmoid(ys * (x.to(params.device) @ w)))
            return lg + lw + N * ll

        def score(P):
            from . import engine
            return engine.score_logreg(P, x.to(P.device), y.to(P.device), N, a0, b0)

        def batch_logp(P):
            from . import engine
            return engine.score_logreg(P, x.to(P.device), y.to(P.device), N, a0, b0, want_logp=True)[1]

        return _LogPdf(logp, score, batch_logp)
