"""Drop-in for the hot-path functions of nvgd/src/stein.py.  The reference's nested
vmap/grad/jacfwd closures are replaced by closed forms evaluated in CUDA;
`kernel` must be a tagged RBF kernel (kernels.get_rbf_kernel), `logp` a callable on (d,)
torch tensors (scores through torch.func.vmap(grad)) or an object with a batched `.score(X)`."""
from __future__ import annotations

import torch

from . import engine
from .utils import BatchedFunction


def score(logp, xs):
    """S = vmap(grad(logp))(xs), (n,d) float32.  The reference computes this inside its double
    vmap (stein.py:50); it depends on x_j only, so it is hoisted."""
    if hasattr(logp, "score"):
        s = logp.score(xs)
    else:
        s = torch.func.vmap(torch.func.grad(logp))(xs)
    return s.to(torch.float32).contiguous()


def _require_rbf(kernel, who):
    if not getattr(kernel, "is_rbf", False):
        raise NotImplementedError(f"{who}: only kernels.get_rbf_kernel kernels dispatch to the CUDA engine "
                                  "(north_star: no fallback path)")
    if getattr(kernel, "funnel", False):
        raise NotImplementedError(f"{who}: the funnel kernel is only supported by metrics.get_mmd")
    return kernel.h()


def _phistar_any_bandwidth(kernel, who, Q, X, S, aux=False):
    """phi* for a scalar bandwidth (one engine call) or a (d,) ARD bandwidth vector (kernels.py:52-65 with a vector):
    with x' = x / h (elementwise) the kernel has unit bandwidth, the scores w.r.t. x' are s h, and
    phi_a = phi'_a / h_a -- three elementwise ops around the same engine call."""
    if getattr(kernel, "is_rbf", False) and not getattr(kernel, "funnel", False) and kernel.is_ard():
        sc = kernel.ard_scale(X.device)
        if sc.numel() != X.shape[1]:
            raise ValueError(f"Bandwidth has shape {tuple(sc.shape)}.")
        out = engine.phistar((Q / sc).contiguous(), (X / sc).contiguous(), (S * sc).contiguous(), 1.0, aux=aux)
        if aux:
            return out[0] / sc, out[1] / sc
        return out / sc
    h = _require_rbf(kernel, who)
    return engine.phistar(Q, X, S, h, aux=aux)


class PhiStar(BatchedFunction):
    """phi* as a function of one point; carries what the closed forms need."""

    def __init__(self, kernel, logp, samples, scores=None):
        self.kernel, self.logp, self.samples = kernel, logp, samples
        self.ard = bool(getattr(kernel, "is_rbf", False) and not getattr(kernel, "funnel", False) and kernel.is_ard())
        self.h = None if self.ard else _require_rbf(kernel, "get_phistar")
        self.scores = score(logp, samples) if scores is None else scores
        super().__init__(self._phi_batch)

    def _phi_batch(self, xs):
        if self.ard:
            return _phistar_any_bandwidth(self.kernel, "get_phistar", xs, self.samples, self.scores)
        return engine.phistar(xs, self.samples, self.scores, self.h)


def get_phistar(kernel, logp, samples):
    """stein.py:200-203: returns x -> phi*(x) estimated from `samples`."""
    return PhiStar(kernel, logp, samples)


def phistar_i(xi, x, logp, kernel, aux=True):
    """stein.py:180-197"""
    if xi.ndim > 1:
        raise ValueError(f"Shape of xi must be (d,). Instead, received shape {tuple(xi.shape)}")
    out = _phistar_any_bandwidth(kernel, "phistar_i", xi[None, :], x, score(logp, x), aux=aux)
    if aux:
        return out[0][0], out[1][0]
    return out[0]


def phistar(particles, leaders, logp, kernel):
    """stein.py:206-222: (n,d) updates and (n,2,d) [drift, repulsion] aux; rectangular n x l."""
    return _phistar_any_bandwidth(kernel, "phistar", particles, leaders, score(logp, leaders), aux=True)


def phistar_u(followers, leaders, logp, kernel):
    """stein.py:225-270: followers get the V-statistic over the leaders, leaders the U-statistic (their own
    k(x_i, x_i) term left out, divisor l-1).  Returns (l+m, d) updates for concat(leaders, followers) and (l+m, 2, d) aux.
    Closed form: one rectangular engine call, then for the leader rows phi_U = (l phi_V - s_i) / (l - 1) because the
    diagonal term is k(x_i,x_i) s_i + grad k = s_i.  The reference zeroes the diagonal of the phi matrix only — its aux
    keeps the diagonal terms and is merely divided by l-1 (stein.py:263-269); reproduced as is."""
    if getattr(kernel, "is_ard", lambda: False)():
        raise NotImplementedError("phistar_u: scalar bandwidth only")
    particles = torch.cat([leaders, followers], dim=0)
    s_lead = score(logp, leaders)
    phi, aux = _phistar_any_bandwidth(kernel, "phistar_u", particles, leaders, s_lead, aux=True)
    l = leaders.shape[0]
    phi = phi.clone(); aux = aux.clone()
    phi[:l] = (l * phi[:l] - s_lead) / (l - 1)
    aux[:l] = aux[:l] * (l / (l - 1))
    return phi, aux


def ksd_squared_l(samples, logp, k, return_stddev=False):
    """stein.py:353-380: the linear-time estimator mean_i u_p(x_i, y_i) over the two halves of `samples` (an odd count
    drops the last sample).  O(n d): plain elementwise device ops, no pair engine involved."""
    h = _require_rbf(k, "ksd_squared_l")
    n = samples.shape[0] - (samples.shape[0] % 2)
    xs, ys = samples[: n // 2], samples[n // 2: n]
    sx, sy = score(logp, xs), score(logp, ys)
    h2 = (h * h) if isinstance(h, torch.Tensor) else float(h) ** 2
    diff = xs - ys
    r2 = torch.sum(diff * diff, dim=1)
    kk = torch.exp(-r2 / (2 * h2))
    d = samples.shape[1]
    outs = kk * (torch.sum(sx * sy, dim=1) + torch.sum((sx - sy) * diff, dim=1) / h2 + d / h2 - r2 / (h2 * h2))
    if return_stddev:
        return torch.mean(outs), torch.std(outs, unbiased=True) / xs.shape[0]
    return torch.mean(outs)


def stein_discrepancy(xs, logp, f, aux=False):
    """stein.py:99-106 for a witness f built from get_phistar (optionally through
    utils.negative / utils.mul): mean_i[ s_i . f(x_i) + div f(x_i) ] in closed form.
    Arbitrary callables f would need jacfwd through user code -> NotImplementedError."""
    base = getattr(f, "base", None)
    if not isinstance(base, PhiStar):
        raise NotImplementedError("stein_discrepancy: f must come from stein.get_phistar")
    if base.ard:
        raise NotImplementedError("stein_discrepancy: the closed-form divergence is built for a scalar bandwidth")
    if aux:
        raise NotImplementedError("stein_discrepancy(aux=True) is not on the hot path")
    s_q = score(logp, xs)
    _, stats = engine.phistar(xs, base.samples, base.scores, base.h, Sq=s_q, stats=True)
    val = stats[0] / stats[2]
    scale = f.scale
    return (val * scale).to(torch.float32) if isinstance(scale, torch.Tensor) else (val * scale).to(torch.float32)


def ksd_squared(xs, ys, logp, k):
    """stein.py:273-295"""
    h = _require_rbf(k, "ksd_squared")
    return engine.ksd(xs, score(logp, xs), h, "rect", ys, score(logp, ys)).to(torch.float32)


def ksd_squared_u(xs, logp, k):
    """stein.py:299-323"""
    h = _require_rbf(k, "ksd_squared_u")
    return engine.ksd(xs, score(logp, xs), h, "u").to(torch.float32)


def ksd_squared_v(xs, logp, k, dummy_arg1=None, dummy_arg2=None):
    """stein.py:327-349"""
    h = _require_rbf(k, "ksd_squared_v")
    return engine.ksd(xs, score(logp, xs), h, "v").to(torch.float32)
