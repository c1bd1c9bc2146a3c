"""N4: scores of the canonical targets.  CPU: the closed forms of the oracle against autodiff of the
line-for-line restatement of the reference's target (oracle/structural.py) and against scipy.stats.  GPU: the device
kernels (csrc/scores.cu, through the C ABI) against the oracle, and BASELINE configs[1] / configs[2] as workloads."""
import numpy as np
import pytest
import torch

from oracle import structural as st
from oracle import svgd_oracle as so


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("F", [54, 55])
def test_logreg_closed_form_vs_autodiff_of_literal_restatement(F):
    P, Xb, yb, N = so.logreg_workload(12, F, B=40, seed=F)
    x64, y64 = torch.tensor(Xb, dtype=torch.float64), torch.tensor(yb, dtype=torch.float64)
    logp = st.logreg_logp_literal(x64, y64, N)
    Pt = torch.tensor(P, dtype=torch.float64)
    lp_ref = torch.stack([logp(p) for p in Pt]).numpy()
    s_ref = torch.func.vmap(torch.func.grad(logp))(Pt).numpy()
    assert np.allclose(so.logreg_logp(P, Xb, yb, N), lp_ref, rtol=1e-12, atol=1e-9)
    assert rel(so.logreg_score(P, Xb, yb, N), s_ref) < 1e-12


def test_logreg_literal_restatement_uses_scipy_logpdfs():
    from scipy import stats
    rng = np.random.default_rng(0)
    w, la = rng.standard_normal(7), 0.3
    alpha = np.exp(la)
    want = stats.gamma.logpdf(alpha, 1.0, scale=1 / 0.01) + np.sum(stats.norm.logpdf(w, scale=1 / np.sqrt(alpha)))
    got = st.logreg_prior_logp(torch.tensor(w), torch.tensor(la, dtype=torch.float64)).item()
    assert got == pytest.approx(want, rel=1e-12)


def test_funnel_and_gaussian_closed_forms_vs_autodiff():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((9, 5))
    fl = lambda x: -x[-1] ** 2 / 18.0 - 0.5 * 4 * x[-1] - 0.5 * torch.exp(-x[-1]) * torch.sum(x[:-1] ** 2)
    s_ref = torch.func.vmap(torch.func.grad(fl))(torch.tensor(X)).numpy()
    assert rel(so.funnel_score(X), s_ref) < 1e-12
    A = rng.standard_normal((5, 5)); cov = A @ A.T + np.eye(5); mean = rng.standard_normal(5)
    prec = torch.tensor(np.linalg.inv(cov))
    gl = lambda x: -0.5 * (x - torch.tensor(mean)) @ prec @ (x - torch.tensor(mean))
    s_ref = torch.func.vmap(torch.func.grad(gl))(torch.tensor(X)).numpy()
    assert rel(so.gaussian_score(X, mean, cov), s_ref) < 1e-12


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("n,F,B", [(8192, 55, 128), (8192, 54, 128), (300, 7, 33), (1000, 130, 64)])
def test_logreg_score_kernel_vs_oracle(n, F, B):
    from nsvgd import engine
    P, Xb, yb, N = so.logreg_workload(n, F, B=B, seed=n + F)
    S, lp = engine.score_logreg(dev(P), dev(Xb), dev(yb), N, want_logp=True)
    assert rel(S.cpu(), so.logreg_score(P, Xb, yb, N)) < 2e-5
    ref_lp = so.logreg_logp(P, Xb, yb, N)
    assert np.abs(lp.cpu().numpy() - ref_lp).max() < 2e-5 * np.abs(ref_lp).max()
    with pytest.raises(ValueError):
        engine.score_logreg(dev(P[:, :-1]), dev(Xb), dev(yb), N)


@pytest.mark.gpu
def test_funnel_and_gaussian_score_kernels_vs_oracle():
    from nsvgd import distributions, engine
    rng = np.random.default_rng(2)
    for d in (2, 5, 40):
        X = (rng.standard_normal((777, d)) * 1.5).astype(np.float32)
        S, lp = engine.score_funnel(dev(X), want_logp=True)
        assert rel(S.cpu(), so.funnel_score(X)) < 1e-5
        tgt = distributions.Funnel(d)
        assert rel(tgt.logpdf.score(dev(X)).cpu(), so.funnel_score(X)) < 1e-5
        from scipy import stats
        ref_lp = stats.norm.logpdf(X[:, -1], 0, 3) + np.sum(stats.norm.logpdf(X[:, :-1], 0, np.exp(X[:, -1:] / 2)), axis=1)
        assert np.abs(lp.cpu().numpy() - ref_lp).max() < 1e-4 * (1 + np.abs(ref_lp).max())
    d = 50
    A = rng.standard_normal((d, d)); Q, _ = np.linalg.qr(A)
    var = np.logspace(-2, 0, d)
    cov = Q.T @ np.diag(var) @ Q                                # the rotated Gaussian of distributions.py:633-639
    mean = rng.standard_normal(d)
    X = (rng.standard_normal((2048, d))).astype(np.float32)
    for c in (cov, var):
        tgt = distributions.Gaussian(mean, c)
        assert rel(tgt.logpdf.score(dev(X)).cpu(), so.gaussian_score(X, mean, c)) < 2e-5


@pytest.mark.gpu
def test_cfg2_gaussian_n2048_d50_step_with_ksd_u_and_mmd():
    """BASELINE configs[1]: SVGD on the 50-D correlated Gaussian, n = 2048, KSD_U + MMD tracked every step
    (flows.py:70-110 with the metrics of metrics.py:88-117); 5 steps against the fp64 oracle replay."""
    from nsvgd import distributions, engine, kernels, models, stein, metrics
    rng = np.random.default_rng(3)
    d, n, m = 50, 2048, 500
    A = rng.standard_normal((d, d)); Q, _ = np.linalg.qr(A)
    cov = Q.T @ np.diag(np.logspace(-2, 0, d)) @ Q
    tgt = distributions.Gaussian(np.zeros(d), cov)
    X0 = rng.standard_normal((n, d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) @ np.linalg.cholesky(cov).T).astype(np.float32)
    kg = models.KernelGradient(target_logp=tgt.logpdf, bandwidth=None, scaled=True, lambda_reg=0.5)
    mmd = metrics.get_mmd(kernels.get_rbf_kernel(1.0))
    Xd, Yd = dev(X0).clone(), dev(Y)
    X = X0.astype(np.float64)
    lr = 0.1
    for step in range(5):
        grads, aux = kg.step_sgd(Xd, None, lr)
        ksd_u = float(stein.ksd_squared_u(Xd, tgt.logpdf, kernels.get_rbf_kernel(aux["bandwidth"])))
        mmd_v = float(mmd(Xd, Yd))
        # oracle replay on the previous particles
        S = so.gaussian_score(X.astype(np.float32), np.zeros(d), cov).astype(np.float32)
        Xn, g_ref, h_ref, extras = so.svgd_step_sgd(X.astype(np.float32), S, lr, scaled=True, lambda_reg=0.5)
        assert np.float32(aux["bandwidth"].item()).tobytes() == np.float32(h_ref).tobytes()
        assert rel(grads.cpu(), g_ref) < 2e-4
        X = np.asarray(Xd.cpu().numpy(), np.float64)            # continue from the GPU state (no error compounding)
        Sn = so.gaussian_score(X.astype(np.float32), np.zeros(d), cov).astype(np.float32)
        ref_u = so.ksd_squared_u(X.astype(np.float32), Sn, float(h_ref))
        assert ksd_u == pytest.approx(ref_u, rel=1e-3, abs=1e-4 * abs(so.ksd_squared_v(X.astype(np.float32), Sn, float(h_ref))))
        assert mmd_v == pytest.approx(so.mmd(X.astype(np.float32), Y, 1.0), abs=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("F", [54, 55])
def test_cfg3_logreg_n8192_step(F):
    """BASELINE configs[2]: Bayesian logistic regression posterior, d = 55 (BASELINE) and 56 (the reference's
    55 features + log alpha), n = 8192 particles, minibatch 128, median bandwidth, scaled sgd
    (bayesian_logistic_regression.py:209-225).  Scores on the device; one step against the oracle on sampled rows."""
    from nsvgd import distributions, models
    n, B = 8192, 128
    P, Xb, yb, N = so.logreg_workload(n, F, B=B, seed=F)
    tgt = distributions.LogisticRegressionPosterior(F, N)
    kg = models.KernelGradient(get_target_logp=lambda batch: tgt.get_minibatch_logp(*batch), scaled=True)
    Pd = dev(P).clone()
    grads, aux = kg.step_sgd(Pd, (torch.tensor(Xb), torch.tensor(yb)), 1e-6)
    S = so.logreg_score(P, Xb, yb, N).astype(np.float32)
    h_ref = so.median_heuristic(P, "literal") if False else None   # 33M pairs x 56: the oracle median is too slow here
    h = float(aux["bandwidth"].item())
    ref = so.phistar_stats_f64_blocked(P, S, h, block=1024)
    g_ref = -ref["alpha"] * ref["phi"]
    assert rel(grads.cpu(), g_ref) < 2e-4
    assert rel(Pd.cpu(), P.astype(np.float64) - 1e-6 * g_ref) < 1e-5      # |lr * grads| ~ |P| here: 2e-4 on the gradients -> ~2e-6
    assert np.isfinite(aux["logp"].cpu().numpy()).all()
    assert np.abs(aux["logp"].cpu().numpy() - so.logreg_logp(P, Xb, yb, N)).max() < 2e-5 * np.abs(so.logreg_logp(P, Xb, yb, N)).max()
