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


def _is_full(kernel):
    return bool(getattr(kernel, "is_full", False))


def _is_ard(kernel):
    return bool(getattr(kernel, "is_rbf", False) and not getattr(kernel, "funnel", False) and not _is_full(kernel) and kernel.is_ard())


def _phistar_any_bandwidth(kernel, who, Q, X, S, aux=False):
    """phi* for a scalar bandwidth (one engine call), a (d,) ARD bandwidth vector (kernels.py:52-65 with a vector) or a
    full (d, d) matrix (kernels.py:83-88).  With x' = A x (A = diag(1/h) or Sigma) the kernel has unit bandwidth, the
    scores w.r.t. x' are A^-T s, and phi = A^T phi': a linear map on the device (nsvgd_scale_cols_f32 / nsvgd_matmul_f32)
    before and after the same engine call."""
    if _is_ard(kernel):
        sc = kernel.ard_scale(X.device)
        if sc.numel() != X.shape[1]:
            raise ValueError(f"Bandwidth has shape {tuple(sc.shape)}.")
        out = engine.phistar(engine.scale_cols(Q, sc, True), engine.scale_cols(X, sc, True), engine.scale_cols(S, sc, False), 1.0, aux=aux)
        if aux:
            d = X.shape[1]
            return engine.scale_cols(out[0], sc, True), engine.scale_cols(out[1].reshape(-1, d), sc, True).reshape(out[1].shape)
        return engine.scale_cols(out, sc, True)
    if _is_full(kernel):
        sig = kernel.sigma.to(device=X.device, dtype=torch.float32)
        if sig.shape[0] != X.shape[1]:
            raise ValueError(f"sigma has shape {tuple(sig.shape)}.")
        sig_t = sig.T.contiguous()
        sig_inv = torch.linalg.inv(sig.double()).float().contiguous()       # d x d set-up on the caller's matrix
        out = engine.phistar(engine.matmul(Q, sig_t), engine.matmul(X, sig_t), engine.matmul(S, sig_inv), 1.0, aux=aux)
        if aux:
            d = X.shape[1]
            return engine.matmul(out[0], sig.contiguous()), engine.matmul(out[1].reshape(-1, d), sig.contiguous()).reshape(out[1].shape)
        return engine.matmul(out, sig.contiguous())
    h = _require_rbf(kernel, who)
    return engine.phistar(Q, X, S, h, aux=aux)


class PhiStar(BatchedFunction):
    """phi* as a function of one point; carries what the closed forms need."""

    def __init__(self, kernel, logp, samples, scores=None):
        self.kernel, self.logp, self.samples = kernel, logp, samples
        self.ard = _is_ard(kernel) or _is_full(kernel)           # a non-scalar bandwidth: linear map around the engine
        self.ard_vector = _is_ard(kernel)
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
    drops the last sample).  One warp per pair on the device (nsvgd_ksd_linear_f32)."""
    h = _require_rbf(k, "ksd_squared_l")
    n = samples.shape[0] - (samples.shape[0] % 2)
    out = engine.ksd_linear(samples[:n], score(logp, samples[:n]), h)
    if return_stddev:
        return out[0].to(torch.float32), out[1].to(torch.float32)
    return out[0].to(torch.float32)


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


def _wants_grad(h):
    return isinstance(h, torch.Tensor) and h.requires_grad and torch.is_grad_enabled()


def ksd_squared(xs, ys, logp, k):
    """stein.py:273-295.  Differentiable w.r.t. a bandwidth tensor that requires grad (autodiff.py); a (d,)
    ARD bandwidth vector uses the per-coordinate closed form (nsvgd_ksd_ard_f32)."""
    if _is_ard(k):
        return engine.ksd_ard(xs, score(logp, xs), k.ard_scale(xs.device), "rect", ys, score(logp, ys)).to(torch.float32)
    h = _require_rbf(k, "ksd_squared")
    if _wants_grad(h):
        from . import autodiff
        return autodiff.ksd_squared(xs, score(logp, xs), h, "rect", ys, score(logp, ys))
    return engine.ksd(xs, score(logp, xs), h, "rect", ys, score(logp, ys)).to(torch.float32)


def ksd_squared_u(xs, logp, k):
    """stein.py:299-323"""
    if _is_ard(k):
        return engine.ksd_ard(xs, score(logp, xs), k.ard_scale(xs.device), "u").to(torch.float32)
    h = _require_rbf(k, "ksd_squared_u")
    if _wants_grad(h):
        from . import autodiff
        return autodiff.ksd_squared(xs, score(logp, xs), h, "u")
    return engine.ksd(xs, score(logp, xs), h, "u").to(torch.float32)


def ksd_squared_v(xs, logp, k, dummy_arg1=None, dummy_arg2=None):
    """stein.py:327-349"""
    if _is_ard(k):
        return engine.ksd_ard(xs, score(logp, xs), k.ard_scale(xs.device), "v").to(torch.float32)
    h = _require_rbf(k, "ksd_squared_v")
    if _wants_grad(h):
        from . import autodiff
        return autodiff.ksd_squared(xs, score(logp, xs), h, "v")
    return engine.ksd(xs, score(logp, xs), h, "v").to(torch.float32)
