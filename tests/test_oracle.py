"""CPU: the oracle's closed forms against the golden vectors (produced by the independent
torch.func structural mirror, tests/golden/make_golden.py) and against its own brute-force forms."""
import numpy as np
import pytest

from oracle import svgd_oracle as so

from conftest import GOLDEN_CASES

RTOL = 1e-6   # closed form (fp64, fp32-rounded scores) vs structural mirror (fp64 scores)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("tag", ["hfix", "hmed"])
def test_closed_forms_match_structural_mirror(golden, name, tag):
    g = golden[name]
    X, Y, Sx, Sy = g["X"], g["Y"], g["Sx"], g["Sy"]
    h = float(g[f"h_{tag}"])
    phi, aux = so.phistar(X, X, Sx, h, aux=True)
    np.testing.assert_allclose(phi, g[f"phi_{tag}"], rtol=RTOL, atol=RTOL * np.abs(g[f"phi_{tag}"]).max())
    np.testing.assert_allclose(aux, g[f"aux_{tag}"], rtol=RTOL, atol=RTOL * np.abs(g[f"aux_{tag}"]).max())
    np.testing.assert_allclose(so.phistar(Y, X, Sx, h), g[f"phi_rect_{tag}"], rtol=RTOL,
                               atol=RTOL * np.abs(g[f"phi_rect_{tag}"]).max())
    assert so.ksd_squared_u(X, Sx, h) == pytest.approx(float(g[f"ksd_u_{tag}"]), rel=1e-5)
    assert so.ksd_squared_v(X, Sx, h) == pytest.approx(float(g[f"ksd_v_{tag}"]), rel=1e-5)
    assert so.ksd_squared(X, Sx, Y, Sy, h) == pytest.approx(float(g[f"ksd_rect_{tag}"]), rel=1e-5)
    sd = so.stein_discrepancy_phistar(X, Sx, h)
    assert sd == pytest.approx(float(g[f"sd_phistar_{tag}"]), rel=1e-5)
    assert sd == pytest.approx(so.ksd_squared_v(X, Sx, h), rel=1e-9)      # identity noted in SURVEY 8a-A5
    l2 = so.l2_norm_squared(phi)
    assert l2 == pytest.approx(float(g[f"l2_{tag}"]), rel=1e-5)
    _, _, extras = so.kernel_gradient(X, Sx, bandwidth=h, scaled=True)
    assert extras["alpha"] == pytest.approx(float(g[f"alpha_{tag}"]), rel=1e-5)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mmd_matches_structural_mirror(golden, name):
    g = golden[name]
    scale = 1e-7
    assert so.mmd(g["X"], g["Y"], 1.0) == pytest.approx(float(g["mmd_rbf"]), rel=1e-6, abs=scale)
    assert so.mmd(g["X"], g["Y"], 1.0, funnel=True) == pytest.approx(float(g["mmd_funnel"]), rel=1e-6, abs=scale)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_bandwidth_golden_bit_exact(golden, name):
    g = golden[name]
    for mode in so.MEDIAN_MODES:
        h = so.median_heuristic(g["X"], mode)
        assert h.dtype == np.float32
        assert h.tobytes() == np.float32(g[f"h_{mode}"]).tobytes(), mode
    # z-transform closed form of the literal mode vs the literal O(n^3)-style execution
    assert float(g["h_literal"]) == pytest.approx(float(g["h_literal_bruteforce"]), rel=2e-6)


@pytest.mark.parametrize("n,d", [(2, 1), (3, 2), (7, 3), (30, 50), (50, 50), (51, 50), (33, 1), (12, 12)])
def test_literal_z_equals_literal_execution(n, d):
    rng = np.random.default_rng(n * 100 + d)
    x = rng.standard_normal((n, d))
    y = so.remove_diagonal_literal(x)
    assert y.shape == (n, n - 1)
    z, w = so.literal_z(x)
    D_lit = so.sqdist(y, y)
    zz = z.copy()
    zz[:, -1] *= np.sqrt(w)
    np.testing.assert_allclose(so.sqdist(zz, zz), D_lit, rtol=1e-12, atol=1e-12)
    h_bf = so.median_heuristic_bruteforce_literal(x)
    h = float(so.median_heuristic(x.astype(np.float32), "literal"))
    assert h == pytest.approx(h_bf, rel=5e-6, abs=1e-7)


@pytest.mark.parametrize("n", list(range(2, 41)))
def test_median_ranks_reproduce_np_median(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, 3))
    D = so.sqdist(x, x)
    p = np.sort(D[np.triu_indices(n, 1)])

    def val(r):
        return 0.0 if r < 0 else p[r]
    lo, hi = so.median_ranks(n, "full")
    assert 0.5 * (val(lo) + val(hi)) == pytest.approx(np.median(D), rel=1e-14, abs=1e-300)
    lo, hi = so.median_ranks(n, "offdiag")
    off = D[~np.eye(n, dtype=bool)]
    assert 0.5 * (val(lo) + val(hi)) == pytest.approx(np.median(off), rel=1e-14)


def test_nan_and_degenerate_bandwidth():
    x = np.random.default_rng(0).standard_normal((10, 3)).astype(np.float32)
    x[3, 1] = np.nan
    assert np.isnan(so.median_heuristic(x, "offdiag"))
    same = np.ones((6, 2), dtype=np.float32)          # duplicates -> med = 0 -> h = 0 (Appendix A)
    assert float(so.median_heuristic(same, "offdiag")) == 0.0


def test_phistar_from_dlogp_is_a_different_estimator():
    """parity trap #3 (stein.py:163-177): score at x_i instead of x_j."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((40, 3)); S = -X * np.array([1.0, 4.0, 0.25])
    K = so.kernel_matrix(X, X, 1.0)
    n = X.shape[0]
    variant = (K.sum(1, keepdims=True) * S + (K.sum(1, keepdims=True) * X - K @ X)) / n
    assert np.linalg.norm(variant - so.phistar(X, X, S, 1.0)) > 1e-2


def test_blocked_f32_port_matches_fp64():
    X, S = so.gaussian_diag_workload(600, 24, seed=3)
    ref = so.phistar(X, X, S, 2.5)
    got = so.phistar_f32_blocked(X, X, S, 2.5, block=128)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-4
