"""Structural mirror of nvgd/src/stein.py, kernels.py, metrics.py, models.py in torch.func.
TEST INFRASTRUCTURE ONLY (see oracle/svgd_oracle.py header; parity pin: tests/golden/refshim + ref_*.npz).

The reference writes the hot path as nested ``vmap`` over per-pair closures with
``grad`` / ``jacfwd`` inside.  JAX is not installable here, so this file re-states
the same *structure* with ``torch.func.{vmap, grad, jacfwd}`` in fp64, for an
arbitrary ``logp`` callable and small n.  It shares no algebra with the closed
forms in svgd_oracle.py, which makes agreement between the two a real check, and
it generates the golden vectors under tests/golden.
"""
from __future__ import annotations

import torch
from torch.func import grad, jacfwd, vmap


def get_rbf_kernel(bandwidth):
    """kernels.py:52-65 (squared=True, normalize=False)."""
    def rbf(x, y):
        h_squared = bandwidth ** 2
        return torch.exp(-torch.sum((x - y) ** 2 / h_squared) / 2)
    return rbf


def defunnelize(z):
    """kernels.py:112-116"""
    x, y = z[:-1], z[-1:]
    return torch.cat([x * torch.exp(-y / 2), y / 3])


def get_funnel_kernel(bandwidth):
    """kernels.py:118-123"""
    rbf = get_rbf_kernel(bandwidth)

    def funnel_kernel(x, y):
        return rbf(defunnelize(x), defunnelize(y))
    return funnel_kernel


def stein_operator(fun, x, logp, transposed=False, aux=False):
    """stein.py:8-68 (scalar / vector branches used on the hot path)."""
    fx = fun(x)
    if transposed:
        assert fx.ndim == 1
        drift_term = torch.inner(grad(logp)(x), fx)                      # stein.py:40
        repulsive_term = torch.trace(jacfwd(fun)(x).transpose(0, 1))     # stein.py:41
    else:
        assert fx.ndim == 0
        drift_term = grad(logp)(x) * fx                                  # stein.py:50
        repulsive_term = grad(fun)(x)                                    # stein.py:51
    out = drift_term + repulsive_term
    if aux:
        return out, torch.stack([drift_term, repulsive_term])
    return out


def stein_expectation(fun, xs, logp, transposed=False, aux=False):
    """stein.py:71-96"""
    out = vmap(lambda x: stein_operator(fun, x, logp, transposed, aux))(xs)
    if aux:
        steins, auxdata = out
        return steins.mean(0), auxdata.mean(0)
    return out.mean(0)


def stein_discrepancy(xs, logp, f):
    """stein.py:99-106"""
    return stein_expectation(f, xs, logp, transposed=True)


def phistar_i(xi, x, logp, kernel, aux=True):
    """stein.py:180-197"""
    def kx(y):
        return kernel(y, xi)
    return stein_expectation(kx, x, logp, aux=aux)


def get_phistar(kernel, logp, samples):
    """stein.py:200-203"""
    def phistar(x):
        return phistar_i(x, samples, logp, kernel, aux=False)
    return phistar


def phistar(particles, leaders, logp, kernel):
    """stein.py:206-222"""
    return vmap(lambda p: phistar_i(p, leaders, logp, kernel, True))(particles)


def _g(x, y, logp, k):
    """the pair function of stein.py:286-291 / 311-315 / 339-343"""
    def inner(x_):
        return stein_operator(lambda y_: k(x_, y_), y, logp)
    return stein_operator(inner, x, logp, transposed=True)


def ksd_squared(xs, ys, logp, k):
    """stein.py:273-295"""
    gv = vmap(lambda x, y: _g(x, y, logp, k), (0, None))
    gvv = vmap(gv, (None, 0))
    return gvv(xs, ys).mean()


def ksd_squared_u(xs, logp, k):
    """stein.py:299-323"""
    gv = vmap(lambda x, y: _g(x, y, logp, k), (0, None))
    m = vmap(gv, (None, 0))(xs, xs)
    n = xs.shape[0]
    m = m - torch.diag(torch.diag(m))
    return m.sum() / (n * (n - 1))


def ksd_squared_v(xs, logp, k):
    """stein.py:327-349"""
    gv = vmap(lambda x, y: _g(x, y, logp, k), (0, None))
    m = vmap(gv, (None, 0))(xs, xs)
    return m.sum() / xs.shape[0] ** 2


def l2_norm_squared(samples, fun):
    """utils.py:483-487"""
    return vmap(lambda x: torch.inner(fun(x), fun(x)))(samples).mean()


def remove_diagonal(matrix):
    """utils.py:653-658 for a genuine (n,n) matrix."""
    n = matrix.shape[0]
    mask = ~torch.eye(n, dtype=torch.bool)
    return matrix[mask].reshape(n, n - 1)


def get_mmd(kernel):
    """metrics.py:88-105"""
    kernel_matrix = vmap(vmap(kernel, (0, None)), (None, 0))

    def mmd(xs, ys):
        kxx = remove_diagonal(kernel_matrix(xs, xs))
        kyy = remove_diagonal(kernel_matrix(ys, ys))
        kxy = kernel_matrix(xs, ys)
        return kxx.mean() + kyy.mean() - 2 * kxy.mean()
    return mmd


def get_field_scaled(particles, logp, bandwidth, lambda_reg=0.5):
    """models.py:677-689 with a given bandwidth: returns (alpha, phi values, ksd, l2)."""
    kernel = get_rbf_kernel(bandwidth)
    phi = get_phistar(kernel, logp, particles)
    l2 = l2_norm_squared(particles, phi)
    ksd = stein_discrepancy(particles, logp, phi)
    alpha = ksd / (2 * lambda_reg * l2)
    return alpha, vmap(phi)(particles), ksd, l2


# ---- log-densities used by the golden script and tests ----------------------
def funnel_logp(x):
    """Neal's funnel, distributions.py:311-319 (up to a constant)."""
    xs, y = x[:-1], x[-1]
    dm1 = xs.shape[0]
    return -y ** 2 / 18.0 - 0.5 * dm1 * y - 0.5 * torch.exp(-y) * torch.sum(xs ** 2)


def make_gaussian_logp(mean, precision):
    def logp(x):
        c = x - mean
        return -0.5 * torch.dot(c, precision @ c)
    return logp


# ---------------------------------------------------------------------------------------------------------------
# Line-for-line torch mirror of the logistic-regression target (experiments/bayesian_logistic_regression.py), fp64.
# Differentiated with torch.func.grad it is the witness the closed-form scores of oracle/svgd_oracle.py
# (and through them the CUDA kernel csrc/scores.cu) are pinned against (tests/test_scores.py).
# ---------------------------------------------------------------------------------------------------------------
def _gamma_logpdf(x, a, scale):
    """scipy.stats.gamma.logpdf(x, a, scale=scale)"""
    import math
    return (a - 1) * torch.log(x) - x / scale - a * math.log(scale) - math.lgamma(a)


def _norm_logpdf(x, scale):
    """scipy.stats.norm.logpdf(x, scale=scale)"""
    import math
    return -0.5 * (x / scale) ** 2 - torch.log(scale) - 0.5 * math.log(2 * math.pi)


def logreg_prior_logp(w, log_alpha, a0=1.0, b0=0.01):
    """prior_logp, single-parameter branch                                                  :73-92"""
    alpha = torch.exp(log_alpha)
    logp_alpha = torch.sum(_gamma_logpdf(alpha, a0, scale=1 / b0))
    logp_w = torch.sum(_norm_logpdf(w, scale=1 / torch.sqrt(alpha)))
    return logp_alpha + logp_w


def logreg_loglikelihood(y, x, w):
    """loglikelihood: log p(y | x, w) = sum_i(logp(yi | xi, w))                             :103-114"""
    y = ((y - 1 / 2) * 2)
    logits = x @ w
    prob_y = torch.special.expit(logits * y)
    return torch.sum(torch.log(prob_y))


def logreg_logp_literal(x, y, num_datapoints, a0=1.0, b0=0.01):
    """get_minibatch_logp(x, y) -> logp(params), w.ndim == 1 branch                          :169-190"""
    def logp(params):
        w, log_alpha = params[:-1], params[-1]                       # unravel  :162-166
        log_prior = logreg_prior_logp(w, log_alpha, a0, b0)
        mean_loglikelihood = logreg_loglikelihood(y, x, w)
        return log_prior + num_datapoints * mean_loglikelihood
    return logp
