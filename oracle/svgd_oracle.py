"""CPU oracle for the SVGD particle-interaction hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
path (``projects-2020-neural-svgd_b200/``) never does: it fails loudly when the
CUDA library is missing.

PARITY PIN: the reference (ratschlab/projects-2020-Neural-SVGD) ships no
tests, golden vectors or fixtures for this path, and JAX/jaxlib/haiku/optax are
absent (no network).  The pin is therefore the reference's own UNMODIFIED
sources executed in this container under torch-backed stand-ins for those
dependencies (tests/golden/refshim/, tests/golden/make_reference_golden.py ->
tests/golden/ref_*.npz); tests/test_reference_golden.py checks every function
of this file against those fixtures.  What stays unpinned is XLA's last-ulp
floating-point behaviour (the stand-in computes with torch).  A second witness:
oracle/structural.py, an independent torch.func mirror of nvgd/src/stein.py
(tests/golden/make_golden.py).

Every function cites the reference file:line it restates (paths relative to
the reference checkout).

Conventions: X (n,d) particles, S (n,d) scores grad log p(x_j), Q (m,d) query
points, r2_ij = |q_i - x_j|^2, k_ij = exp(-r2_ij / (2 h^2))   [kernels.py:60-64].
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32

MEDIAN_MODES = ("literal", "offdiag", "full")


# --------------------------------------------------------------------------
# A1  RBF kernel                                            kernels.py:52-65
# --------------------------------------------------------------------------
def rbf(x, y, h):
    """k(x,y) = exp(-sum((x-y)^2 / h^2) / 2)                 kernels.py:60-64"""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return np.exp(-np.sum((x - y) ** 2 / (float(h) ** 2)) / 2)


def sqdist(Q, X):
    """fp64 pairwise squared distances by direct differences (m,n)."""
    Q = np.asarray(Q, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    out = np.empty((Q.shape[0], X.shape[0]))
    for i in range(Q.shape[0]):  # row loop keeps memory at O(n d)
        diff = Q[i][None, :] - X
        out[i] = np.einsum("jd,jd->j", diff, diff)
    return out


def kernel_matrix(Q, X, h):
    return np.exp(-sqdist(Q, X) / (2.0 * float(h) ** 2))


# --------------------------------------------------------------------------
# A2  median heuristic                 kernels.py:160-172, utils.py:178-197,
#                                      utils.py:653-658
# --------------------------------------------------------------------------
def remove_diagonal_literal(x):
    """What ``utils.remove_diagonal`` (utils.py:653-658) yields when it is applied
    to the PARTICLES (n,d) as kernels.py:169 does: the (row, col) index pairs of
    the off-diagonal of an n x n matrix, with JAX's clamp-on-out-of-bounds gather
    for col >= d.  Result (n, n-1)."""
    x = np.asarray(x)
    n, d = x.shape
    rows, cols = np.nonzero(~np.eye(n, dtype=bool))
    cols = np.minimum(cols, d - 1)  # JAX gather clamps; NumPy would raise
    return x[rows, cols].reshape(n, n - 1)


def literal_z(x):
    """O(nd) closed form of the literal mode: the pairwise squared
    distances of ``remove_diagonal_literal(x)`` equal those of z_a with the last
    coordinate weighted by w.  Returns (Z (n, dz) same dtype as x, w float).

    n > d : z_a = [x_a without column a (a < d-1) | x_a[:d-1] (a >= d-1), x_a[d-1]],
            dz = d, w = n - d
    n <= d: z_a = x_a[:n] without column a, dz = n - 1, w = 1
    """
    x = np.asarray(x)
    n, d = x.shape
    if n > d:
        z = np.empty((n, d), dtype=x.dtype)
        for a in range(n):
            if a < d - 1:
                z[a, : d - 1] = np.delete(x[a], a)[: d - 1]
            else:
                z[a, : d - 1] = x[a, : d - 1]
            z[a, d - 1] = x[a, d - 1]
        return z, float(n - d)
    z = np.empty((n, n - 1), dtype=x.dtype)
    for a in range(n):
        z[a] = np.delete(x[a, :n], a)
    return z, 1.0


def pair_sqdist_f32(z, w_last=1.0):
    """DEFINED-ORDER fp32 squared distances of all unique pairs a<b (the bit-exact
    contract shared with the CUDA select kernel, csrc/pair_engine.cuh):

        acc = 0
        for c in 0..dz-2:  acc = fl(acc + fl(diff_c * diff_c))       (no FMA)
        t = fl(diff_last * diff_last); t = fl(w * t); acc = fl(acc + t)

    Restates utils.squared_distance_matrix (utils.py:178-197) with an explicit
    summation order (XLA's own order is unknowable here).  Returns the (P,)
    float32 vector, P = n(n-1)/2, pairs in row-major upper-triangular order."""
    z = np.ascontiguousarray(z, dtype=F32)
    n, dz = z.shape
    w = F32(w_last)
    iu, ju = np.triu_indices(n, k=1)
    out = np.empty(iu.shape[0], dtype=F32)
    step = 1 << 20
    for s in range(0, iu.shape[0], step):
        a = z[iu[s:s + step]]
        b = z[ju[s:s + step]]
        acc = np.zeros(a.shape[0], dtype=F32)
        for c in range(dz - 1):
            diff = a[:, c] - b[:, c]
            acc = acc + diff * diff          # each op rounds to fp32, no fusion
        diff = a[:, dz - 1] - b[:, dz - 1]
        t = diff * diff
        t = w * t
        acc = acc + t
        out[s:s + step] = acc
    return out


def median_ranks(n, mode):
    """Appendix B of SURVEY.md: which order statistics of the P unique pair
    distances make up ``np.median`` (kernels.py:170) in each mode.

    Returns (r_lo, r_hi) where each r is either -1 (the value 0.0 coming from the
    n diagonal zeros) or a 0-based rank into the ascending pair distances; the
    median is (v(r_lo) + v(r_hi)) * 0.5 in fp32 (r_lo == r_hi for an odd count,
    in which case the midpoint of two equal numbers is exact)."""
    P = n * (n - 1) // 2
    if mode == "offdiag":
        return (P - 1) // 2, P // 2
    if mode in ("full", "literal"):
        N = n * n

        def f(r):
            return -1 if r < n else (r - n) // 2
        if N % 2 == 0:
            return f(N // 2 - 1), f(N // 2)
        return f((N - 1) // 2), f((N - 1) // 2)
    raise ValueError(f"unknown median mode {mode!r}")


def median_sq(x, mode="literal"):
    """The squared-distance median (float32) of kernels.py:169-170 per mode."""
    x = np.asarray(x, dtype=F32)
    if x.ndim == 1:
        x = x[:, None]
    n = x.shape[0]
    if mode == "literal":
        z, w = literal_z(x)
    else:
        z, w = x, 1.0
    p = np.sort(pair_sqdist_f32(z, w))
    if np.isnan(p).any():
        return F32(np.nan)
    r_lo, r_hi = median_ranks(n, mode)
    lo = F32(0.0) if r_lo < 0 else p[r_lo]
    hi = F32(0.0) if r_hi < 0 else p[r_hi]
    return F32((lo + hi) * F32(0.5))


def log_np1_f32(n):
    """log(n+1) evaluated in fp64 on the host and rounded once to fp32 — the value
    both the oracle and the CUDA finalize kernel use (kernels.py:171)."""
    return F32(math.log(float(n + 1)))


def median_heuristic(x, mode="literal"):
    """h = sqrt(0.5 * medsq / log(n + 1))  in fp32            kernels.py:160-172"""
    x = np.asarray(x, dtype=F32)
    n = x.shape[0]
    med = median_sq(x, mode)
    with np.errstate(divide="ignore", invalid="ignore"):
        return F32(np.sqrt(F32(F32(0.5) * med) / log_np1_f32(n)))


def median_heuristic_bruteforce_literal(x):
    """kernels.py:169-171 executed literally in fp64 (O(n^3) memory-free loop form):
    remove_diagonal on the particles, full (n,n) distance matrix incl. diagonal,
    np.median over all n^2 entries.  Small n only; validates literal_z + ranks."""
    y = remove_diagonal_literal(np.asarray(x, dtype=np.float64))
    D = sqdist(y, y)
    med = np.median(D)
    return math.sqrt(0.5 * med / math.log(x.shape[0] + 1))


# --------------------------------------------------------------------------
# A3/A4  Stein operator, phi*                       stein.py:8-96, 180-222
# --------------------------------------------------------------------------
def phistar(Q, X, S, h, aux=False):
    """phi*(q_i) = (1/n) sum_j [ k_ij s_j + k_ij (q_i - x_j)/h^2 ]  (V-statistic)
    stein.py:180-203 (get_phistar / phistar_i) with the non-transposed scalar
    branch of stein_operator, stein.py:49-53.  fp64.

    aux=True also returns the (m,2,d) [drift, repulsion] split of stein.phistar
    (stein.py:206-222)."""
    Q = np.asarray(Q, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    n = X.shape[0]
    h2 = float(h) ** 2
    K = kernel_matrix(Q, X, h)                     # (m,n)
    drift = K @ S / n
    rep = (K.sum(1, keepdims=True) * Q - K @ X) / (n * h2)
    if aux:
        return drift + rep, np.stack([drift, rep], axis=1)
    return drift + rep


def div_phistar(Q, X, S, h):
    """div_x phi*(x) at each query (needed by stein_discrepancy with f = phi*):
    (1/n) sum_j k_ij [ s_j.(x_j - q_i)/h^2 + d/h^2 - r2_ij/h^4 ]."""
    Q = np.asarray(Q, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    n, d = X.shape
    h2 = float(h) ** 2
    r2 = sqdist(Q, X)
    K = np.exp(-r2 / (2 * h2))
    c = np.einsum("jd,jd->j", S, X)
    t1 = (K @ c - np.einsum("id,id->i", Q, K @ S)) / h2
    return (t1 + (d / h2) * K.sum(1) - (K * r2).sum(1) / h2 ** 2) / n


def stein_discrepancy_phistar(X, S, h):
    """stein.stein_discrepancy(xs, logp, f=phi*) = mean_i[ s_i.phi_i + div phi(x_i) ]
    stein.py:99-106 with the transposed vector branch stein.py:39-43."""
    phi = phistar(X, X, S, h)
    return float(np.mean(np.einsum("id,id->i", np.asarray(S, np.float64), phi) + div_phistar(X, X, S, h)))


def l2_norm_squared(phi):
    """utils.l2_norm_squared: mean_i |phi(x_i)|^2               utils.py:483-487"""
    phi = np.asarray(phi, dtype=np.float64)
    return float(np.mean(np.einsum("id,id->i", phi, phi)))


# --------------------------------------------------------------------------
# A6  KSD family                                           stein.py:273-349
# --------------------------------------------------------------------------
def ksd_matrix(X, Sx, Y, Sy, h):
    """u_p(x_i, y_j) = k [ s_x.s_y + (s_x - s_y).(x - y)/h^2 + d/h^2 - r2/h^4 ]
    (n,m) — the closed form of the nested stein_operator in stein.py:286-291."""
    X = np.asarray(X, np.float64); Y = np.asarray(Y, np.float64)
    Sx = np.asarray(Sx, np.float64); Sy = np.asarray(Sy, np.float64)
    d = X.shape[1]
    h2 = float(h) ** 2
    r2 = sqdist(X, Y)
    K = np.exp(-r2 / (2 * h2))
    ss = Sx @ Sy.T
    # (s_x - s_y).(x - y) = s_x.x - s_x.y - s_y.x + s_y.y
    cross = (np.einsum("id,id->i", Sx, X)[:, None] - Sx @ Y.T
             - X @ Sy.T + np.einsum("jd,jd->j", Sy, Y)[None, :])
    return K * (ss + cross / h2 + d / h2 - r2 / h2 ** 2)


def ksd_squared(X, Sx, Y, Sy, h):
    """stein.ksd_squared: mean over the n x m matrix           stein.py:273-295"""
    return float(np.mean(ksd_matrix(X, Sx, Y, Sy, h)))


def ksd_squared_u(X, S, h):
    """stein.ksd_squared_u: sum_{i != j} / (n(n-1))            stein.py:299-323"""
    M = ksd_matrix(X, S, X, S, h)
    n = M.shape[0]
    np.fill_diagonal(M, 0.0)
    return float(M.sum() / (n * (n - 1)))


def ksd_squared_v(X, S, h):
    """stein.ksd_squared_v: sum_{i,j} / n^2                    stein.py:327-349"""
    M = ksd_matrix(X, S, X, S, h)
    return float(M.sum() / M.shape[0] ** 2)


# --------------------------------------------------------------------------
# A7/A8  KernelGradient + Particles._step            models.py:143-167, 622-689
# --------------------------------------------------------------------------
def kernel_gradient(X, S, bandwidth=None, scaled=False, lambda_reg=0.5, median_mode="literal"):
    """KernelGradient.gradient (models.py:662-675): returns (grads (n,d), h, extras).
    grads = -phi (get_field, models.py:654-660) or -alpha*phi with
    alpha = ksd / (2 lambda l2) (get_field_scaled, models.py:677-689)."""
    h = bandwidth if bandwidth else float(median_heuristic(X, median_mode))   # truthiness test: models.py:657
    phi = phistar(X, X, S, h)
    if not scaled:
        return -phi, h, {}
    ksd = stein_discrepancy_phistar(X, S, h)
    l2 = l2_norm_squared(phi)
    alpha = ksd / (2 * lambda_reg * l2)
    return -alpha * phi, h, {"ksd": ksd, "l2": l2, "alpha": alpha}


def svgd_step_sgd(X, S, lr, **kw):
    """Particles._step with optax.sgd (models.py:158-160): x <- x - lr * grads."""
    grads, h, extras = kernel_gradient(X, S, **kw)
    return np.asarray(X, np.float64) - lr * grads, grads, h, extras


# --------------------------------------------------------------------------
# A9  MMD                                  metrics.py:88-127, kernels.py:112-123
# --------------------------------------------------------------------------
def defunnelize(Z):
    """kernels.defunnelize, batched: [x * exp(-y/2), y/3]      kernels.py:112-116"""
    Z = np.asarray(Z, dtype=np.float64)
    x, y = Z[:, :-1], Z[:, -1:]
    return np.concatenate([x * np.exp(-y / 2), y / 3], axis=1)


def mmd(X, Y, h=1.0, funnel=False):
    """metrics.get_mmd(kernel)(xs, ys): mean(offdiag Kxx) + mean(offdiag Kyy) - 2 mean(Kxy)
    metrics.py:96-104; funnel=True applies get_funnel_kernel (kernels.py:118-123)."""
    X = np.asarray(X, np.float64); Y = np.asarray(Y, np.float64)
    if funnel:
        X, Y = defunnelize(X), defunnelize(Y)
    n, m = X.shape[0], Y.shape[0]
    kxx = kernel_matrix(X, X, h); kyy = kernel_matrix(Y, Y, h); kxy = kernel_matrix(X, Y, h)
    sxx = (kxx.sum() - np.trace(kxx)) / (n * (n - 1))
    syy = (kyy.sum() - np.trace(kyy)) / (m * (m - 1))
    return float(sxx + syy - 2 * kxy.mean())


# --------------------------------------------------------------------------
# Synthetic workloads — shared by tests and bench.py so that both
# arms see identical inputs.  Analytic scores; numpy RNG, seeded.
# --------------------------------------------------------------------------
def funnel_score(X):
    """grad log p of Neal's funnel (distributions.py:311-319): x_k | y ~ N(0, e^y), y ~ N(0, 9)."""
    X = np.asarray(X, dtype=np.float64)
    x, y = X[:, :-1], X[:, -1:]
    dm1 = X.shape[1] - 1
    gx = -x * np.exp(-y)
    gy = -y / 9.0 - dm1 / 2.0 + 0.5 * np.exp(-y) * np.sum(x * x, axis=1, keepdims=True)
    return np.concatenate([gx, gy], axis=1)


def gaussian_diag_workload(n, d, seed=0, shift=0.5):
    """cfg4-style workload: X ~ N(0, I) + shift, S = -(X - mu) / diag(Sigma),
    Sigma = diag(logspace(-2, 0, d)), mu = 0.  fp32 arrays."""
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) + shift).astype(F32)
    var = np.logspace(-2, 0, d)
    S = (-(X.astype(np.float64)) / var[None, :]).astype(F32)
    return X, S


# --------------------------------------------------------------------------
# fp32 blocked CPU path (BLAS) — the "port" timed as cpu_baseline by bench.py.
# Same math as phistar(), Gram-trick + sgemm, row-blocked to bound memory.
# --------------------------------------------------------------------------
def phistar_f32_blocked(Q, X, S, h, block=2048):
    Q = np.ascontiguousarray(Q, dtype=F32); X = np.ascontiguousarray(X, dtype=F32)
    S = np.ascontiguousarray(S, dtype=F32)
    n = X.shape[0]
    h2 = F32(h) * F32(h)
    mu = X.mean(0, dtype=np.float64).astype(F32)
    Xc = X - mu
    xn = np.einsum("jd,jd->j", Xc, Xc)
    V = np.concatenate([S, Xc], axis=1)
    out = np.empty_like(Q)
    d = X.shape[1]
    for s in range(0, Q.shape[0], block):
        q = Q[s:s + block] - mu
        qn = np.einsum("id,id->i", q, q)
        r2 = np.maximum(qn[:, None] + xn[None, :] - F32(2) * (q @ Xc.T), F32(0))
        K = np.exp(r2 * (F32(-0.5) / h2))
        KV = K @ V
        rs = K.sum(1, keepdims=True)
        out[s:s + block] = (KV[:, :d] + (rs * q - KV[:, d:]) / h2) / F32(n)
    return out


# --------------------------------------------------------------------------
# fp64 blocked evaluation of phi* and the KSD / l2 statistics at BASELINE sizes
# (n = 65536): torch CPU on all cores, row-blocked so that memory stays at
# O(block * n).  Same formulas as phistar() / div_phistar() above
# (stein.py:180-203, 99-106; utils.py:483-487; models.py:677-689).
# --------------------------------------------------------------------------
def phistar_stats_f64_blocked(X, S, h, block=512, rows=None, lambda_reg=0.5):
    """Returns dict(phi=(n,d) fp64 [or only the requested `rows`], ksd, l2, alpha).
    ksd / l2 are means over ALL n query rows, so every block is evaluated; `rows`
    only limits what is kept of phi."""
    import torch
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    Xt = torch.as_tensor(np.asarray(X, np.float64)); St = torch.as_tensor(np.asarray(S, np.float64))
    n, d = Xt.shape
    h2 = float(h) ** 2
    xn = (Xt * Xt).sum(1)
    c = (St * Xt).sum(1)
    V = torch.cat([St, Xt, torch.ones(n, 1, dtype=torch.float64), c[:, None]], dim=1)      # K @ [S | X | 1 | c]
    keep = None if rows is None else np.asarray(rows)
    phi_keep = np.empty((n if keep is None else len(keep), d))
    pos = None if keep is None else {int(r): i for i, r in enumerate(keep)}
    t_sum, l_sum = 0.0, 0.0
    for s in range(0, n, block):
        q = Xt[s:s + block]
        r2 = (xn[s:s + block, None] + xn[None, :] - 2.0 * (q @ Xt.T)).clamp_(min=0.0)
        K = torch.exp(r2 * (-0.5 / h2))
        KV = K @ V
        KS, KX, rs, Kc = KV[:, :d], KV[:, d:2 * d], KV[:, 2 * d], KV[:, 2 * d + 1]
        phi = (KS + (rs[:, None] * q - KX) / h2) / n
        b = (K * r2).sum(1)
        div = ((Kc - (q * KS).sum(1)) / h2 + (d / h2) * rs - b / h2 ** 2) / n
        t_sum += float(((St[s:s + block] * phi).sum(1) + div).sum())
        l_sum += float((phi * phi).sum())
        if keep is None:
            phi_keep[s:s + block] = phi.numpy()
        else:
            for r in range(s, min(s + block, n)):
                if r in pos:
                    phi_keep[pos[r]] = phi[r - s].numpy()
    ksd, l2 = t_sum / n, l_sum / n
    return {"phi": phi_keep, "ksd": ksd, "l2": l2, "alpha": ksd / (2 * lambda_reg * l2)}


def phistar_f32_blocked_torch(Q, X, S, h, block=2048):
    """The same fp32 Gram-trick port as phistar_f32_blocked, on torch CPU so that exp and the GEMMs use every host core
    (NumPy's exp is single-threaded).  bench.py's CPU arm; inputs are torch float32 CPU tensors or arrays."""
    import torch
    Q = torch.as_tensor(Q, dtype=torch.float32); X = torch.as_tensor(X, dtype=torch.float32)
    S = torch.as_tensor(S, dtype=torch.float32)
    n, d = X.shape
    h2 = float(h) ** 2
    mu = X.double().mean(0).float()
    Xc = X - mu
    xn = (Xc * Xc).sum(1)
    V = torch.cat([S, Xc], dim=1)
    out = torch.empty_like(Q)
    for s in range(0, Q.shape[0], block):
        q = Q[s:s + block] - mu
        qn = (q * q).sum(1)
        r2 = torch.addmm(qn[:, None] + xn[None, :], q, Xc.T, alpha=-2.0).clamp_(min=0.0)
        K = r2.mul_(-0.5 / h2).exp_()
        KV = K @ V
        rs = K.sum(1, keepdim=True)
        out[s:s + block] = (KV[:, :d] + (rs * q - KV[:, d:]) / h2) / n
    return out


# --------------------------------------------------------------------------
# N4  scores of the canonical targets (closed forms, fp64).  The logistic-
# regression posterior restates experiments/bayesian_logistic_regression.py:
#   prior_logp              :73-92    gamma.logpdf(alpha; a0, scale=1/b0) + sum norm.logpdf(w; scale=alpha^-1/2)
#   loglikelihood           :103-114  SUM_i log expit(y'_i x_i.w),  y' = (y - 1/2) * 2
#   get_minibatch_logp.logp :179-189  log_prior + num_datapoints * loglikelihood   (single-parameter branch: the
#                                     plug-in differentiates one particle at a time, models.py:662-675 / stein.py:86-89)
# params = ravel(w, log_alpha) = [w, log_alpha]  (:158-159); alpha = exp(log_alpha), no Jacobian term (:84-85).
# oracle/structural.py::logreg_logp_literal is the line-for-line torch mirror these closed forms are pinned against.
# --------------------------------------------------------------------------
def logreg_logp(P, Xb, yb, num_datapoints, a0=1.0, b0=0.01):
    P = np.asarray(P, np.float64); Xb = np.asarray(Xb, np.float64)
    ys = (np.asarray(yb, np.float64) - 0.5) * 2.0
    w, la = P[:, :-1], P[:, -1]
    F = w.shape[1]
    alpha = np.exp(la)
    lg = (a0 - 1.0) * la - b0 * alpha + a0 * math.log(b0) - math.lgamma(a0)
    lw = -0.5 * alpha * (w * w).sum(1) + 0.5 * F * la - 0.5 * F * math.log(2 * math.pi)
    t = (w @ Xb.T) * ys[None, :]
    ll = -(np.logaddexp(0.0, -t)).sum(1)                         # log expit(t) = -log(1 + e^-t)
    return lg + lw + num_datapoints * ll


def logreg_score(P, Xb, yb, num_datapoints, a0=1.0, b0=0.01):
    P = np.asarray(P, np.float64); Xb = np.asarray(Xb, np.float64)
    ys = (np.asarray(yb, np.float64) - 0.5) * 2.0
    w, la = P[:, :-1], P[:, -1]
    F = w.shape[1]
    alpha = np.exp(la)
    t = (w @ Xb.T) * ys[None, :]
    g = num_datapoints * ys[None, :] * (1.0 / (1.0 + np.exp(t)))   # d/dz log expit(y' z) = y' expit(-y' z)
    s_w = g @ Xb - w * alpha[:, None]
    s_la = (a0 - 1.0) - b0 * alpha + 0.5 * F - 0.5 * alpha * (w * w).sum(1)
    return np.concatenate([s_w, s_la[:, None]], axis=1)


def gaussian_score(X, mean, cov):
    """-(x - mean) cov^-1; cov a (d,) diagonal or (d,d) matrix        distributions.py:89-154"""
    X = np.asarray(X, np.float64); mean = np.asarray(mean, np.float64); cov = np.asarray(cov, np.float64)
    prec = np.diag(1.0 / cov) if cov.ndim == 1 else np.linalg.inv(cov)
    return -(X - mean) @ prec


def logreg_workload(n, F, B=128, seed=0, num_datapoints=418_000):
    """cfg3-style synthetic stand-in for the covertype run (no dataset in the image): standardised features + intercept
    column (:41), labels from a random ground-truth weight vector, particles from the prior (:66-71, a0=1, b0=0.01).
    num_datapoints ~ the reference's len(x_train) = 0.8 * 0.9 * 581012."""
    rng = np.random.default_rng(seed)
    Xb = np.concatenate([rng.standard_normal((B, F - 1)), np.ones((B, 1))], axis=1).astype(F32)
    w_true = rng.standard_normal(F) * 0.5
    yb = (rng.random(B) < 1.0 / (1.0 + np.exp(-(Xb @ w_true)))).astype(F32)
    alpha = rng.gamma(1.0, size=n) / 0.01
    P = np.concatenate([rng.standard_normal((n, F)), np.log(alpha)[:, None]], axis=1).astype(F32)
    return P, Xb, yb, float(num_datapoints)
