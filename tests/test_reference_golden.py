"""CPU: pins oracle/ against fixtures produced by EXECUTING the unmodified reference sources
(/root/reference/nvgd/src/*.py) under the torch-backed stand-ins for its missing third-party dependencies
(tests/golden/refshim/, generator tests/golden/make_reference_golden.py).  `*_f64.npz`: float32 aliased to float64
in the stand-in (source unmodified) -> tight tolerances on the closed forms; `*_f32.npz`: float32 as the reference
runs it -> fp32 round-off tolerances."""
import os

import numpy as np
import pytest

from oracle import svgd_oracle as so
from conftest import GOLDEN_CASES, GOLDEN_DIR


def _load(name, tag):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"ref_{name}_{tag}.npz")))


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


TOL = {"f64": 2e-7, "f32": 3e-5}   # f64: limited by the fp32-rounded scores the oracle consumes (6e-8 relative)


def _ulp_diff(a, b):
    ia = np.float32(a).view(np.int32).astype(np.int64)
    ib = np.float32(b).view(np.int32).astype(np.int64)
    return int(abs(ia - ib))


def test_literal_median_vs_reference_execution():
    """A2: kernels.median_heuristic as the reference executes it (gather clamping, (n,n,n-1) tile, jnp.median).
    The oracle's O(nd) z-transform + defined-order fp32 distances must agree with the reference's own fp32 run to
    rounding (different summation order => not bit-equal; <= 2 ulp) and with its fp64 run to fp32 precision."""
    r32 = dict(np.load(os.path.join(GOLDEN_DIR, "ref_median_f32.npz")))
    r64 = dict(np.load(os.path.join(GOLDEN_DIR, "ref_median_f64.npz")))
    shapes = [k[2:] for k in r32 if k.startswith("X_") and not k.startswith("X_1d")]
    assert len(shapes) >= 12
    for tag in shapes:
        X = r32["X_" + tag]
        np.testing.assert_array_equal(X, r64["X_" + tag].astype(np.float32))
        h = so.median_heuristic(X, "literal")
        assert _ulp_diff(h, r32["h_" + tag]) <= 2, (tag, h, r32["h_" + tag])
        assert float(h) == pytest.approx(float(r64["h_" + tag]), rel=3e-7), tag
        # brute-force emulation of the literal execution in fp64 == the reference in fp64
        assert so.median_heuristic_bruteforce_literal(X) == pytest.approx(float(r64["h_" + tag]), rel=1e-12), tag
    # (n,) input: the docstring allows it, the reference raises (remove_diagonal indexes two axes)
    assert str(r32["h_1d_40_error"]) == "IndexError"


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_vs_reference_execution(name, tag):
    r = _load(name, tag)
    tol = TOL[tag]
    X, Y = r["X"], r["Y"]
    Sx, Sy = r["Sx"].astype(np.float32), r["Sy"].astype(np.float32)      # what the CUDA path consumes
    assert _rel(Sx, r["Sx"]) < 1e-7
    for ht in ("hfix", "hmed"):
        h = float(r[f"h_{ht}"])
        assert _rel(so.kernel_matrix(Y, X, h), r[f"K_{ht}"]) < tol                      # A1
        phi, aux = so.phistar(X, X, Sx, h, aux=True)                                      # A3/A4
        assert _rel(phi, r[f"phi_{ht}"]) < tol
        assert _rel(phi, r[f"phi_closure_{ht}"]) < tol
        assert _rel(aux, r[f"aux_{ht}"]) < tol
        phi_r, aux_r = so.phistar(Y, X, Sx, h, aux=True)
        assert _rel(phi_r, r[f"phi_rect_{ht}"]) < tol
        assert _rel(aux_r, r[f"aux_rect_{ht}"]) < tol
        ktol = tol * 30                                                                   # sums of signed O(1) terms
        assert so.stein_discrepancy_phistar(X, Sx, h) == pytest.approx(float(r[f"sd_phistar_{ht}"]), rel=ktol)   # A5
        assert so.l2_norm_squared(phi) == pytest.approx(float(r[f"l2_{ht}"]), rel=ktol)
        assert so.ksd_squared_u(X, Sx, h) == pytest.approx(float(r[f"ksd_u_{ht}"]), rel=ktol)                     # A6
        assert so.ksd_squared_v(X, Sx, h) == pytest.approx(float(r[f"ksd_v_{ht}"]), rel=ktol)
        assert so.ksd_squared(X, Sx, Y, Sy, h) == pytest.approx(float(r[f"ksd_rect_{ht}"]), rel=ktol)
    mtol = 1e-12 if tag == "f64" else 2e-7                                                # A9 (absolute: MMD ~ 0)
    assert so.mmd(X, Y, 1.0) == pytest.approx(float(r["mmd_rbf"]), abs=mtol)
    assert so.mmd(X, Y, 1.0) == pytest.approx(float(r["mmd_default_tracer"]), abs=mtol)
    assert so.mmd(X, Y, 1.0) == pytest.approx(float(r["mmd_tracer_rbf"]), abs=mtol)
    assert so.mmd(X, Y, 1.0, funnel=True) == pytest.approx(float(r["mmd_funnel"]), abs=mtol)


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_kernel_gradient_vs_reference_execution(name, tag):
    """A7: KernelGradient.gradient in the four (bandwidth, scaled) corners."""
    r = _load(name, tag)
    tol = TOL[tag] * 30
    X, Sx = r["X"], r["Sx"].astype(np.float32)
    hfix = float(r["h_hfix"])
    for gt, bw, scaled in (("med_scaled", None, True), ("med_plain", None, False),
                           ("fix_scaled", hfix, True), ("fix_plain", hfix, False)):
        h_ref = float(r[f"grad_{gt}_h"])
        if bw is None:   # the reference's own median, to fp32 rounding
            assert float(so.median_heuristic(X, "literal")) == pytest.approx(h_ref, rel=3e-7)
        g, h, _ = so.kernel_gradient(X, Sx, bandwidth=bw if bw else h_ref, scaled=scaled, lambda_reg=0.5)
        assert _rel(g, r[f"grad_{gt}"]) < tol, gt


def _oracle_score(name, X):
    from golden.make_reference_golden import CASES, inputs
    (_, n, d, m, seed, kind, _), = [c for c in CASES if c[0] == name]
    if kind == "funnel":
        return so.funnel_score(X)
    _, _, mean, prec = inputs(n, d, m, seed, kind)
    mean, prec = mean.astype(np.float32).astype(np.float64), prec.astype(np.float32).astype(np.float64)
    return -(np.asarray(X, np.float64) - mean) @ (0.5 * (prec + prec.T))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_particle_steps_vs_reference_execution(name):
    """A8: three Particles.step calls (sgd) and four svgd_flow steps, re-played with the oracle."""
    r = _load(name, "f64")
    X = r["X"].astype(np.float64)
    assert _rel(_oracle_score(name, X), r["Sx"]) < 1e-6          # analytic scores == reference vmap(grad(logp))
    lr = float(r["step3_sgd_lr"])
    for i in range(3):
        h = float(so.median_heuristic(X.astype(np.float32), "literal"))
        assert h == pytest.approx(float(r["step3_sgd_h"][i]), rel=1e-6)
        X, g, _, _ = so.svgd_step_sgd(X, _oracle_score(name, X), lr, bandwidth=h, scaled=True, lambda_reg=0.5)
        assert np.linalg.norm(g) == pytest.approx(float(r["step3_sgd_gnorm"][i]), rel=1e-5)
    assert _rel(X, r["step3_sgd_X"]) < 1e-6
    assert so.mmd(r["X"], r["Y"], 1.0) == pytest.approx(float(r["step3_sgd_mmd0"]), abs=1e-10)
    # flows.svgd_flow defaults: bandwidth=1, scaled, lr=0.1, lambda=1/2 (flows.py:70-80)
    X = r["X"].astype(np.float64)
    for i in range(4):
        X, _, _, _ = so.svgd_step_sgd(X, _oracle_score(name, X), 1e-1, bandwidth=1.0, scaled=True, lambda_reg=0.5)
    assert _rel(X, r["flow4_X"]) < 1e-6


def test_ard_bandwidth_rescaling_vs_reference_execution():
    """N3: kernels.get_rbf_kernel with a (d,) bandwidth vector, executed by the reference (stein.phistar, get_phistar,
    metrics.get_mmd).  The product serves it by rescaling x' = x/h, s' = s h, phi = phi'/h around a unit-bandwidth engine
    call; this pins that identity (oracle on the rescaled cloud == the reference with the vector)."""
    r = dict(np.load(os.path.join(GOLDEN_DIR, "ref_ard_f64.npz")))
    X, Y, S, bw = r["X"].astype(np.float64), r["Y"].astype(np.float64), r["Sx"].astype(np.float64), r["bandwidth"]
    phi, aux = so.phistar(X / bw, X / bw, S * bw, 1.0, aux=True)
    assert _rel(phi / bw, r["phi"]) < 1e-12 and _rel(aux / bw, r["aux"]) < 1e-12
    assert _rel(phi / bw, r["phi_closure"]) < 1e-12
    phi_r, aux_r = so.phistar(Y / bw, X / bw, S * bw, 1.0, aux=True)
    assert _rel(phi_r / bw, r["phi_rect"]) < 1e-12 and _rel(aux_r / bw, r["aux_rect"]) < 1e-12
    assert so.mmd(X / bw, Y / bw, 1.0) == pytest.approx(float(r["mmd"]), abs=1e-13)
    assert _rel(so.kernel_matrix(Y / bw, X / bw, 1.0), r["K"]) < 1e-12


def test_phistar_u_closed_form_vs_reference_execution():
    """stein.phistar_u (stein.py:225-270): leaders' rows are (l phi_V - s_i)/(l-1); the reference's aux keeps its diagonal
    terms and is only divided by l-1 — both reproduced by the closed form the product uses."""
    r = dict(np.load(os.path.join(GOLDEN_DIR, "ref_ard_f64.npz")))
    X, S = r["X"].astype(np.float64), r["Sx"].astype(np.float64)
    l, h = int(r["u_leaders"]), float(r["u_h"])
    phi, aux = so.phistar(X, X[:l], S[:l], h, aux=True)
    phi[:l] = (l * phi[:l] - S[:l]) / (l - 1)
    aux[:l] = aux[:l] * (l / (l - 1))
    assert _rel(phi, r["phi_u"]) < 1e-12 and _rel(aux, r["aux_u"]) < 1e-12


def test_cfg1_thousand_steps_vs_reference_execution():
    """BASELINE configs[0]: 2-D funnel, n = 100, median-heuristic bandwidth, scaled field, sgd 5e-2, 1000 steps of the
    reference's own Particles / KernelGradient loop (tests/golden/ref_cfg1_*.npz), replayed with the oracle.
    The map amplifies rounding between steps ~20 and ~600 (the reference's own float32 and float64 runs differ by 1-3 %
    there and re-converge); so: the first 20 steps are compared tightly, the long horizon through what the experiment
    reports (bandwidth, moments, the two MMD tracers)."""
    r = dict(np.load(os.path.join(GOLDEN_DIR, "ref_cfg1_f64.npz")))
    r32 = dict(np.load(os.path.join(GOLDEN_DIR, "ref_cfg1_f32.npz")))
    X = r["X0"].astype(np.float64)
    lr = float(r["lr"])
    hs, gn = [], []
    for k in range(int(r["n_steps"])):
        h = float(so.median_heuristic(X.astype(np.float32), "literal"))
        X, g, _, _ = so.svgd_step_sgd(X, so.funnel_score(X), lr, bandwidth=h, scaled=True, lambda_reg=0.5)
        hs.append(h); gn.append(np.linalg.norm(g))
    hs, gn = np.asarray(hs), np.asarray(gn)
    np.testing.assert_allclose(hs[:20], r["h"][:20], rtol=1e-6)
    np.testing.assert_allclose(gn[:20], r["gnorm"][:20], rtol=1e-6)
    # the reference against itself (float32 vs float64) sets the scale of what "agreement" can mean later on
    self_dev = np.abs(r32["h"] - r["h"]).max() / r["h"].max()
    assert 1e-3 < self_dev < 0.1
    assert np.abs(hs - r["h"]).max() / r["h"].max() < 2 * max(self_dev, 0.02)
    ref_end = r["snaps"][-1]
    assert hs[-1] == pytest.approx(float(r["h"][-1]), rel=0.01)
    np.testing.assert_allclose(X.mean(0), ref_end.mean(0), atol=0.02)
    np.testing.assert_allclose(X.std(0), ref_end.std(0), rtol=0.02)
    assert _rel(X, ref_end) < 0.02
    # tracers at step 990 (every 10th step, models.py:180): the run ends one sgd step later
    assert so.mmd(X, r["target"], 1.0) == pytest.approx(float(r["rbf_mmd"][-1]), abs=1e-3)
    assert so.mmd(X, r["target"], 1.0, funnel=True) == pytest.approx(float(r["funnel_mmd"][-1]), abs=1e-3)
    assert float(r["funnel_mmd"][-1]) < 0.1 * float(r["funnel_mmd"][0])       # the flow did its job in the reference
