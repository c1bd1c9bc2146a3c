"""GPU: the CUDA path, called through the reference-facing API (nsvgd.{kernels,stein,metrics,models,flows}), against
fixtures produced by executing the unmodified reference sources (tests/golden/ref_*.npz; generator and the
dependency stand-ins: tests/golden/make_reference_golden.py, tests/golden/refshim/README.md).
Tolerance: 1e-4 relative (north star) for phi/KSD/MMD; the bandwidth must match the reference's own fp32 execution
to <= 2 ulp (summation order differs) and is bit-exact against the oracle's defined order (test_gpu_parity.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, GOLDEN_DIR
from golden.make_reference_golden import CASES, inputs

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _load(name, tag="f64"):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"ref_{name}_{tag}.npz")))


def _rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _logp(name):
    """the same log-density the generator handed to the reference, as a torch callable on (d,) CUDA tensors"""
    (_, n, d, m, seed, kind, _), = [c for c in CASES if c[0] == name]
    if kind == "funnel":
        def logp(x):
            xs, y = x[:-1], x[-1]
            return -y ** 2 / 18.0 - 0.5 * xs.shape[0] * y - 0.5 * torch.exp(-y) * torch.sum(xs ** 2)
        return logp
    _, _, mean, prec = inputs(n, d, m, seed, kind)
    mean_t = torch.tensor(mean, dtype=torch.float32, device="cuda")
    prec_t = torch.tensor(prec, dtype=torch.float32, device="cuda")
    return lambda x: -0.5 * torch.dot(x - mean_t, prec_t @ (x - mean_t))


def _ulp_diff(a, b):
    ia = np.float32(a).view(np.int32).astype(np.int64)
    ib = np.float32(b).view(np.int32).astype(np.int64)
    return int(abs(ia - ib))


def test_median_heuristic_vs_reference_execution():
    from nsvgd import kernels
    r32 = dict(np.load(os.path.join(GOLDEN_DIR, "ref_median_f32.npz")))
    r64 = dict(np.load(os.path.join(GOLDEN_DIR, "ref_median_f64.npz")))
    for tag in [k[2:] for k in r32 if k.startswith("X_") and not k.startswith("X_1d")]:
        h = kernels.median_heuristic(torch.tensor(r32["X_" + tag], device="cuda")).item()
        assert _ulp_diff(h, r32["h_" + tag]) <= 2, (tag, h, float(r32["h_" + tag]))
        assert h == pytest.approx(float(r64["h_" + tag]), rel=3e-7), tag


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_stein_and_metrics_vs_reference_execution(name):
    from nsvgd import kernels, stein, metrics, utils, vmap
    r = _load(name)
    logp = _logp(name)
    X = torch.tensor(r["X"], device="cuda")
    Y = torch.tensor(r["Y"], device="cuda")
    assert _rel(stein.score(logp, X), r["Sx"]) < 1e-5
    for ht in ("hfix", "hmed"):
        h = float(r[f"h_{ht}"])
        k = kernels.get_rbf_kernel(h)
        assert float(k(X[0], Y[1])) == pytest.approx(float(r[f"K_{ht}"][1, 0]), rel=1e-5)      # A1: K[b,a]=k(x_a,y_b)
        phi, aux = stein.phistar(X, X, logp, k)                                                    # A3/A4
        assert _rel(phi, r[f"phi_{ht}"]) < RTOL and _rel(aux, r[f"aux_{ht}"]) < RTOL
        phi_r, aux_r = stein.phistar(Y, X, logp, k)
        assert _rel(phi_r, r[f"phi_rect_{ht}"]) < RTOL and _rel(aux_r, r[f"aux_rect_{ht}"]) < RTOL
        f = stein.get_phistar(k, logp, X)
        assert _rel(vmap(f)(X), r[f"phi_closure_{ht}"]) < RTOL
        assert _rel(f(X[5]), r[f"phi_closure_{ht}"][5]) < RTOL
        assert float(stein.stein_discrepancy(X, logp, f)) == pytest.approx(float(r[f"sd_phistar_{ht}"]), rel=RTOL)   # A5
        assert float(utils.l2_norm_squared(X, f)) == pytest.approx(float(r[f"l2_{ht}"]), rel=RTOL)
        scale = abs(float(r[f"ksd_v_{ht}"]))                                                       # A6
        assert float(stein.ksd_squared_u(X, logp, k)) == pytest.approx(float(r[f"ksd_u_{ht}"]), abs=RTOL * scale)
        assert float(stein.ksd_squared_v(X, logp, k, None, None)) == pytest.approx(float(r[f"ksd_v_{ht}"]), rel=RTOL)
        assert float(stein.ksd_squared(X, Y, logp, k)) == pytest.approx(float(r[f"ksd_rect_{ht}"]), abs=RTOL * scale)
    # A9: tolerance relative to the kernel means (~1), MMD itself can be ~0
    assert float(metrics.get_mmd(kernels.get_rbf_kernel(1.))(X, Y)) == pytest.approx(float(r["mmd_rbf"]), abs=RTOL * 1e-1)
    assert float(metrics.get_mmd_tracer(Y)(X)["mmd"]) == pytest.approx(float(r["mmd_default_tracer"]), abs=RTOL * 1e-1)
    ft = metrics.get_funnel_tracer(Y)(X)
    assert float(ft["rbf_mmd"]) == pytest.approx(float(r["mmd_tracer_rbf"]), abs=RTOL * 1e-1)
    assert float(ft["funnel_mmd"]) == pytest.approx(float(r["mmd_funnel"]), abs=RTOL * 1e-1)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_kernel_gradient_vs_reference_execution(name):
    """A7: the plug-in, four (bandwidth, scaled) corners; aux dict contract {"bandwidth", "logp"}."""
    from nsvgd import models
    r = _load(name)
    r32 = _load(name, "f32")
    logp = _logp(name)
    X = torch.tensor(r["X"], device="cuda")
    hfix = float(r["h_hfix"])
    for gt, bw, scaled in (("med_scaled", None, True), ("med_plain", None, False),
                           ("fix_scaled", hfix, True), ("fix_plain", hfix, False)):
        kg = models.KernelGradient(target_logp=logp, bandwidth=bw, scaled=scaled, lambda_reg=0.5)
        g, aux = kg.gradient(None, X, aux=True)
        assert _rel(g, r[f"grad_{gt}"]) < RTOL, gt
        assert _rel(aux["logp"], r[f"grad_{gt}_logp"]) < 1e-5
        assert float(aux["bandwidth"]) == pytest.approx(float(r[f"grad_{gt}_h"]), rel=3e-7)
        if bw is None:
            assert _ulp_diff(float(aux["bandwidth"]), r32[f"grad_{gt}_h"]) <= 2


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_particles_step_vs_reference_execution(name, opt):
    """A8: three Particles.step calls (median bandwidth, scaled) and the tracer's first MMD."""
    from nsvgd import models, metrics
    r = _load(name)
    logp = _logp(name)
    Y = torch.tensor(r["Y"], device="cuda")
    kg = models.KernelGradient(target_logp=logp, bandwidth=None, scaled=True, lambda_reg=0.5)
    p = models.Particles(key=0, gradient=kg.gradient, init_samples=torch.tensor(r["X"]),
                         learning_rate=float(r[f"step3_{opt}_lr"]), optimizer=opt,
                         compute_metrics=metrics.get_mmd_tracer(Y))
    for _ in range(3):
        p.step(None)
    assert _rel(p.particles, r[f"step3_{opt}_X"]) < 1e-5
    moved = _rel(p.particles - torch.tensor(r["X"], device="cuda"), r[f"step3_{opt}_X"] - r["X"].astype(np.float64))
    assert moved < (RTOL if opt == "sgd" else 1e-3)        # the displacement itself, not hidden behind |X|
    hs = [float(h) for h in p.rundata["bandwidth"]]
    np.testing.assert_allclose(hs, r[f"step3_{opt}_h"], rtol=1e-5)
    gn = [float(g) for g in p.rundata["global_grad_norm"]]
    np.testing.assert_allclose(gn, r[f"step3_{opt}_gnorm"], rtol=RTOL if opt == "sgd" else 1e-3)
    assert float(p.rundata["mmd"][0][1]) == pytest.approx(float(r[f"step3_{opt}_mmd0"]), abs=1e-5)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_svgd_flow_vs_reference_execution(name):
    """flows.svgd_flow with its defaults (bandwidth=1, scaled, sgd lr=0.1), four steps."""
    from nsvgd import flows
    r = _load(name)
    logp = _logp(name)
    X0, Y0 = torch.tensor(r["X"]), torch.tensor(r["Y"])

    class _T:
        logpdf = staticmethod(logp)
        sample = staticmethod(lambda k, key=None: Y0[:k])

    class _P:
        sample = staticmethod(lambda k, key=None: X0[:k])

    class _Setup:
        get = staticmethod(lambda: (_T, _P))

    _, parts, err = flows.svgd_flow(0, _Setup, n_particles=X0.shape[0], n_steps=4, catch_exceptions=False)
    assert err is None
    assert _rel(parts.particles, r["flow4_X"]) < 1e-5
    assert _rel(parts.particles - X0.cuda(), r["flow4_X"] - r["X"].astype(np.float64)) < RTOL
    mmd0 = parts.rundata["mmd"][0]
    assert float(mmd0[1]) == pytest.approx(float(r["flow4_mmd0"]), abs=1e-5)


def test_ard_bandwidth_vs_reference_execution():
    """N3: vector (ARD) bandwidth through stein.phistar / get_phistar / phistar_i and metrics.get_mmd"""
    from nsvgd import kernels, stein, metrics, models, vmap
    r = dict(np.load(os.path.join(GOLDEN_DIR, "ref_ard_f64.npz")))
    logp = _logp("gauss_n64_d5")
    X = torch.tensor(r["X"], device="cuda")
    Y = torch.tensor(r["Y"], device="cuda")
    k = kernels.get_rbf_kernel(torch.tensor(r["bandwidth"], dtype=torch.float32))
    assert float(k(X[0].cpu(), Y[1].cpu())) == pytest.approx(float(r["K"][1, 0]), rel=1e-5)
    phi, aux = stein.phistar(X, X, logp, k)
    assert _rel(phi, r["phi"]) < RTOL and _rel(aux, r["aux"]) < RTOL
    phi_r, aux_r = stein.phistar(Y, X, logp, k)
    assert _rel(phi_r, r["phi_rect"]) < RTOL and _rel(aux_r, r["aux_rect"]) < RTOL
    f = stein.get_phistar(k, logp, X)
    assert _rel(vmap(f)(X), r["phi_closure"]) < RTOL
    assert _rel(f(X[3]), r["phi_closure"][3]) < RTOL
    one, _ = stein.phistar_i(X[7], X, logp, k)
    assert _rel(one, r["phi"][7]) < RTOL
    assert float(metrics.get_mmd(k)(X, Y)) == pytest.approx(float(r["mmd"]), abs=1e-5)
    with pytest.raises(NotImplementedError):
        stein.stein_discrepancy(X, logp, f)
    # the closed-form KSD under a vector bandwidth (round 2) against the structural mirror of stein.py
    from oracle import structural as st
    hv = torch.tensor(r["bandwidth"], dtype=torch.float64)
    X64 = torch.tensor(r["X"], dtype=torch.float64)
    (_, n_, d_, m_, seed_, kind_, _), = [c for c in CASES if c[0] == "gauss_n64_d5"]
    _, _, mean_, prec_ = inputs(n_, d_, m_, seed_, kind_)
    mean64, prec64 = torch.tensor(mean_, dtype=torch.float64), torch.tensor(prec_, dtype=torch.float64)
    logp64 = lambda x: -0.5 * torch.dot(x - mean64, prec64 @ (x - mean64))
    ref_u = float(st.ksd_squared_u(X64, logp64, st.get_rbf_kernel(hv)))
    ref_v = float(st.ksd_squared_v(X64, logp64, st.get_rbf_kernel(hv)))
    assert float(stein.ksd_squared_u(X, logp, k)) == pytest.approx(ref_u, abs=RTOL * abs(ref_v))
    # the plug-in (an extension: the reference's KernelGradient raises on a bandwidth vector, models.py:657)
    kg = models.KernelGradient(target_logp=logp, bandwidth=torch.tensor(r["bandwidth"], dtype=torch.float32), scaled=False)
    g = kg.gradient(None, X)
    assert _rel(g, -r["phi"]) < RTOL


def test_phistar_u_and_linear_ksd():
    """N3: leader/follower U-statistic (stein.phistar_u) vs the reference's execution; linear-time KSD vs the oracle's u_p"""
    from nsvgd import kernels, stein
    from oracle import svgd_oracle as so
    r = dict(np.load(os.path.join(GOLDEN_DIR, "ref_ard_f64.npz")))
    logp = _logp("gauss_n64_d5")
    X = torch.tensor(r["X"], device="cuda")
    l, h = int(r["u_leaders"]), float(r["u_h"])
    k = kernels.get_rbf_kernel(h)
    phi_u, aux_u = stein.phistar_u(X[l:], X[:l], logp, k)
    assert _rel(phi_u, r["phi_u"]) < RTOL and _rel(aux_u, r["aux_u"]) < RTOL
    S = r["Sx"].astype(np.float64)
    Xn = r["X"].astype(np.float64)
    for n_use in (64, 63):
        half = n_use // 2
        M = so.ksd_matrix(Xn[:half], S[:half], Xn[half:2 * half], S[half:2 * half], h)
        want = float(np.mean(np.diag(M)))
        got, sd = stein.ksd_squared_l(X[:n_use], logp, k, return_stddev=True)
        assert float(got) == pytest.approx(want, rel=1e-4, abs=1e-6)
        assert float(sd) == pytest.approx(float(np.std(np.diag(M), ddof=1) / half), rel=1e-3)


def test_cfg1_thousand_steps_vs_reference_execution():
    """BASELINE configs[0] on the GPU: 2-D funnel, n = 100, median bandwidth, scaled sgd, 1000 graph-replayed steps
    (Particles.step_many) vs 1000 steps of the reference's own loop (tests/golden/ref_cfg1_f64.npz).  First 20 steps
    tight; beyond that the map amplifies rounding (the reference's float32 and float64 runs differ by 1-3 %), so the long
    horizon is compared through bandwidth, moments and the MMD tracers (tests/test_reference_golden.py has the CPU twin)."""
    from nsvgd import distributions, models, metrics
    r = dict(np.load(os.path.join(GOLDEN_DIR, "ref_cfg1_f64.npz")))
    tgt = distributions.Funnel(2)
    target = torch.tensor(r["target"], device="cuda")
    kg = models.KernelGradient(target_logp=tgt.logpdf, bandwidth=None, scaled=True, lambda_reg=0.5)
    p = models.Particles(key=0, gradient=kg.gradient, init_samples=torch.tensor(r["X0"]), learning_rate=float(r["lr"]),
                         optimizer="sgd", compute_metrics=metrics.get_funnel_tracer(target))
    p.step_many(int(r["n_steps"]), None)
    hs = np.asarray([float(h) for h in p.rundata["bandwidth"]])
    gn = np.asarray([float(g) for g in p.rundata["global_grad_norm"]])
    np.testing.assert_allclose(hs[:20], r["h"][:20], rtol=1e-5)
    np.testing.assert_allclose(gn[:20], r["gnorm"][:20], rtol=1e-4)
    X = p.particles.double().cpu().numpy()
    ref_end = r["snaps"][-1]
    assert hs[-1] == pytest.approx(float(r["h"][-1]), rel=0.02)
    np.testing.assert_allclose(X.mean(0), ref_end.mean(0), atol=0.03)
    np.testing.assert_allclose(X.std(0), ref_end.std(0), rtol=0.03)
    steps = [s for s, _ in p.rundata["rbf_mmd"]]
    assert steps == r["mmd_steps"].tolist()
    assert float(p.rundata["rbf_mmd"][0][1]) == pytest.approx(float(r["rbf_mmd"][0]), abs=1e-5)
    assert float(p.rundata["funnel_mmd"][0][1]) == pytest.approx(float(r["funnel_mmd"][0]), abs=1e-5)
    assert float(p.rundata["rbf_mmd"][-1][1]) == pytest.approx(float(r["rbf_mmd"][-1]), abs=2e-3)
    assert float(p.rundata["funnel_mmd"][-1][1]) == pytest.approx(float(r["funnel_mmd"][-1]), abs=2e-3)
