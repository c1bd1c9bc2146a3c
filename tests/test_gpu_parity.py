"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors and the oracle.
Tolerances: bandwidth bit-exact; phi / KSD / MMD <= 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

from oracle import svgd_oracle as so

from conftest import GOLDEN_CASES

pytestmark = pytest.mark.gpu

TOL = 1e-4
ROW_TOL = 2 * TOL      # worst single row (relative to max(|row|, 1e-3 max|row|)); round 1 allowed 10 x TOL here


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).cuda()


def rel_fro(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)


def max_row_rel(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    num = np.linalg.norm(got - ref, axis=-1)
    den = np.maximum(np.linalg.norm(ref, axis=-1), 1e-3 * np.linalg.norm(ref, axis=-1).max())
    return float((num / den).max())


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_bandwidth_bit_exact_vs_golden(golden, name):
    from nsvgd import engine
    g = golden[name]
    X = dev(g["X"])
    for mode in so.MEDIAN_MODES:
        h = engine.median_heuristic(X, mode).cpu().numpy()[0]
        assert np.float32(h).tobytes() == np.float32(g[f"h_{mode}"]).tobytes(), (mode, h, g[f"h_{mode}"])


@pytest.mark.parametrize("n,d,seed", [(2, 1, 0), (3, 5, 1), (65, 3, 2), (64, 64, 3), (129, 17, 4), (500, 7, 5),
                                      (1000, 50, 6), (40, 100, 7), (777, 128, 8)])
@pytest.mark.parametrize("mode", so.MEDIAN_MODES)
def test_bandwidth_bit_exact_vs_oracle(n, d, seed, mode):
    from nsvgd import engine
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) * 1.7 + 0.3).astype(np.float32)
    if n > 10:
        X[5] = X[2]                      # duplicates (zero distance)
    want = so.median_heuristic(X, mode)
    got = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
    assert np.float32(got).tobytes() == np.float32(want).tobytes(), (got, want)


@pytest.mark.parametrize("n,d,seed", [(1000, 50, 6), (777, 128, 8), (1500, 200, 9), (700, 515, 10), (300, 1030, 11)])
@pytest.mark.parametrize("mode", so.MEDIAN_MODES)
def test_bandwidth_stored_distance_select(n, d, seed, mode, monkeypatch):
    """shapes outside the tensor-core filter (d > 128 or n < 8192) with d > 32: ONE exact pass keeps the distances and the
    later radix passes histogram the buffer -- bit-identical to the oracle and to the 3-pass recomputing select; NaN and
    all-equal inputs behave the same."""
    from nsvgd import engine
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d)) * 0.9 + 0.3).astype(np.float32)
    X[5] = X[2]
    want = so.median_heuristic(X, mode)
    got = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
    monkeypatch.setenv("NSVGD_MEDIAN_STORE", "0")
    three = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
    monkeypatch.delenv("NSVGD_MEDIAN_STORE")
    assert np.float32(got).tobytes() == np.float32(want).tobytes() == np.float32(three).tobytes(), (got, want, three)
    if mode == "offdiag":
        X[11, 2] = np.nan
        assert np.isnan(engine.median_heuristic(dev(X), mode).item())
        assert engine.median_heuristic(dev(np.ones((n, d), np.float32)), mode).item() == 0.0


@pytest.mark.parametrize("n,d,kind", [(2048, 132, "gauss"), (2500, 300, "gauss"), (3000, 1028, "gauss"), (2304, 512, "clusters"),
                                      (2100, 256, "ties"), (2200, 640, "bnn"),
                                      (2100, 1149, "bnn"), (2050, 4594, "bnn"), (2300, 202, "gauss")])     # d % 4 != 0 (4594: the reference's BNN)
@pytest.mark.parametrize("mode", so.MEDIAN_MODES)
def test_bandwidth_large_d_filtered_select(n, d, kind, mode, monkeypatch):
    """d > 128, n >= 2048: the tensor-core filtered select of csrc/median_large.cu (symmetric CTA-pair Gram with the classify
    epilogue, per-tile lists, tile-wise defined-order evaluation) is bit-identical to the exact SIMT select -- on a Gaussian
    cloud, on tight clusters (most pairs near, large positive Gram entries), on integer-valued coordinates (massive ties:
    the lists overflow and the device-side fallback takes over) and in the BNN weight regime."""
    from nsvgd import engine
    rng = np.random.default_rng(n + d)
    if kind == "gauss":
        X = (rng.standard_normal((n, d)) * 0.9 + 0.3).astype(np.float32)
    elif kind == "clusters":
        centres = rng.standard_normal((3, d)) * 2.0
        X = (centres[rng.integers(0, 3, n)] + rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    elif kind == "ties":
        X = rng.integers(0, 2, (n, d)).astype(np.float32)
    else:
        X = (rng.standard_normal((n, d)) * 0.15).astype(np.float32)
    X[5] = X[2]
    got = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
    monkeypatch.setenv("NSVGD_MEDIAN_LARGE", "0")
    want = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
    monkeypatch.delenv("NSVGD_MEDIAN_LARGE")
    assert np.float32(got).tobytes() == np.float32(want).tobytes(), (got, want)
    if (n, d) == (2048, 132):
        assert np.float32(got).tobytes() == np.float32(so.median_heuristic(X, mode)).tobytes()
    if mode == "offdiag" and kind == "gauss":
        Xn = X.copy(); Xn[11, 2] = np.nan
        assert np.isnan(engine.median_heuristic(dev(Xn), mode).item())


@pytest.mark.parametrize("n,d,kind", [(2500, 300, "gauss"), (2200, 640, "bnn"), (2304, 512, "clusters"), (2048, 4096, "bnn")])
def test_bandwidth_large_d_filter_radius_has_margin(n, d, kind, monkeypatch):
    """How conservative is the error radius of the large-d filter?  The classification is only trusted because the radius
    bounds the tensor-core Gram error (assumed <= 4 ulp of the largest partial sum per MMA, csrc/median_large.cu); with BOTH
    radii shrunk 8-fold the select must still return the identical bits on these clouds -- i.e. the actual error is at least
    8 x below the bound (a shrunk radius that misclassified a pair near the target rank would shift the order statistic)."""
    from nsvgd import engine
    rng = np.random.default_rng(n + d)
    if kind == "gauss":
        X = (rng.standard_normal((n, d)) * 0.9 + 0.3).astype(np.float32)
    elif kind == "clusters":
        centres = rng.standard_normal((3, d)) * 2.0
        X = (centres[rng.integers(0, 3, n)] + rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    else:
        X = (rng.standard_normal((n, d)) * 0.15).astype(np.float32)
    for mode in ("literal", "offdiag"):
        want = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
        monkeypatch.setenv("NSVGD_MEDIAN_LARGE_KAPPA_SCALE", "0.125")
        got = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
        monkeypatch.delenv("NSVGD_MEDIAN_LARGE_KAPPA_SCALE")
        assert np.float32(got).tobytes() == np.float32(want).tobytes(), (mode, got, want)


@pytest.mark.parametrize("n,d,h", [(4096, 50, 5.0), (5000, 128, 8.0), (4100, 3, 1.3), (4200, 200, 12.0)])
def test_ksd_u_v_through_phi_statistics(n, d, h, monkeypatch):
    """n >= 4096: ksd_squared_u / _v come out of the fused phi* kernels (stein_discrepancy(phi*) == V-statistic,
    U = (n^2 V - sum_i(|s_i|^2 + d/h^2)) / (n(n-1)); csrc/api.cu nsvgd_ksd_tc_f32) -- against the fp32 SIMT pair tiles
    (nsvgd_ksd_f32) and, for the V statistic, the fp64 oracle."""
    from nsvgd import engine
    X, S = so.gaussian_diag_workload(n, d, seed=n + d, shift=0.5)
    S = (S * 0.1).astype(np.float32)
    Xd, Sd = dev(X), dev(S)
    got = {v: engine.ksd(Xd, Sd, h, v).item() for v in ("u", "v")}
    monkeypatch.setenv("NSVGD_KSD_TC", "0")
    ref = {v: engine.ksd(Xd, Sd, h, v).item() for v in ("u", "v")}
    monkeypatch.delenv("NSVGD_KSD_TC")
    scale = abs(ref["v"])
    for v in ("u", "v"):
        assert got[v] == pytest.approx(ref[v], abs=TOL * scale), (v, got[v], ref[v])
    assert got["v"] == pytest.approx(so.stein_discrepancy_phistar(X, S, h), rel=TOL)


@pytest.mark.parametrize("mode", so.MEDIAN_MODES)
def test_bandwidth_warm_started_bracket_is_bit_identical(mode):
    """engine.median_heuristic(hint=): over an SVGD-like sequence of slowly moving clouds the warm-started bracket gives the
    same bits as the sampled one at every step and is actually used (hits); a jump of the cloud makes the hinted bracket miss
    -- the exact fallback still returns the right bits and the hint is dropped, then rebuilt."""
    from nsvgd import engine
    rng = np.random.default_rng(21)
    n, d = 8192, 64
    X = (rng.standard_normal((n, d)) * 1.1 + 0.2).astype(np.float32)
    V = (rng.standard_normal((n, d)) * 1e-4).astype(np.float32)
    hint = engine.new_median_hint("cuda")
    for step in range(8):
        if step == 5:
            X = (X * 1.5).astype(np.float32)            # the median of the squared distances moves by 2.25 x: a certain miss
        Xd = dev(X)
        want = engine.median_heuristic(Xd, mode).cpu().numpy()[0]
        got = engine.median_heuristic(Xd, mode, hint=hint).cpu().numpy()[0]
        assert np.float32(got).tobytes() == np.float32(want).tobytes(), (step, got, want)
        X = (X + V).astype(np.float32)
    st = engine.median_hint_stats(hint)
    # steps 0, 1 observe; 2, 3, 4 hit; 5 misses (exact fallback) and disarms the hint for 32 calls; 6, 7 observe under cool-down
    assert st["hits"] == 3 and st["misses"] == 1, st
    assert st["seen"] == 1 and st["valid"] == 0 and st["cooldown"] == 30 and st["penalty"] == 1, st
    # the same hint object handed to another problem (a different median mode / shape): keyed, so it is not used (no miss)
    other = [m for m in so.MEDIAN_MODES if m != mode][0]
    Xd = dev(X)
    for m_, x_ in ((other, Xd), (other, Xd), (other, Xd), (mode, Xd[:8192 - 64].contiguous())):
        want = engine.median_heuristic(x_, m_).cpu().numpy()[0]
        got = engine.median_heuristic(x_, m_, hint=hint).cpu().numpy()[0]
        assert np.float32(got).tobytes() == np.float32(want).tobytes(), (m_, got, want)
    st2 = engine.median_hint_stats(hint)
    assert st2["misses"] == 1 and st2["hits"] == 4, st2          # third call in the other mode hit; nothing missed


def test_bandwidth_warm_started_bracket_large_d():
    """the same warm start on the d > 128 filter (csrc/median_large.cu): drifting BNN-regime cloud, literal and offdiag modes"""
    from nsvgd import engine
    rng = np.random.default_rng(31)
    n, d = 2304, 512
    X = (rng.standard_normal((n, d)) * 0.15).astype(np.float32)
    V = (rng.standard_normal((n, d)) * 1e-5).astype(np.float32)
    hints = {m: engine.new_median_hint("cuda") for m in ("literal", "offdiag")}
    for step in range(6):
        Xd = dev(X)
        for m, hint in hints.items():
            want = engine.median_heuristic(Xd, m).cpu().numpy()[0]
            got = engine.median_heuristic(Xd, m, hint=hint).cpu().numpy()[0]
            assert np.float32(got).tobytes() == np.float32(want).tobytes(), (step, m, got, want)
        X = (X + V).astype(np.float32)
    for m, hint in hints.items():
        st = engine.median_hint_stats(hint)
        assert st["hits"] == 4 and st["misses"] == 0, (m, st)


def test_bandwidth_nan_and_zero():
    from nsvgd import engine
    X = np.random.default_rng(0).standard_normal((70, 3)).astype(np.float32)
    X[11, 2] = np.nan
    assert np.isnan(engine.median_heuristic(dev(X), "offdiag").item())
    same = np.ones((9, 4), dtype=np.float32)
    assert engine.median_heuristic(dev(same), "offdiag").item() == 0.0


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("tag", ["hfix", "hmed"])
def test_phi_ksd_against_golden(golden, name, tag):
    from nsvgd import engine
    g = golden[name]
    X, Y, Sx, Sy = dev(g["X"]), dev(g["Y"]), dev(g["Sx"]), dev(g["Sy"])
    h = float(g[f"h_{tag}"])
    phi, aux, stats = engine.phistar(X, X, Sx, h, Sq=Sx, aux=True, stats=True)
    assert rel_fro(phi.cpu(), g[f"phi_{tag}"]) < TOL
    assert max_row_rel(phi.cpu(), g[f"phi_{tag}"]) < ROW_TOL, max_row_rel(phi.cpu(), g[f"phi_{tag}"])
    assert rel_fro(aux.cpu(), g[f"aux_{tag}"]) < TOL
    assert rel_fro(engine.phistar(Y, X, Sx, h).cpu(), g[f"phi_rect_{tag}"]) < TOL
    stats = stats.cpu().numpy()
    n = X.shape[0]
    assert stats[2] == n
    assert stats[0] / n == pytest.approx(float(g[f"sd_phistar_{tag}"]), rel=TOL)
    assert stats[1] / n == pytest.approx(float(g[f"l2_{tag}"]), rel=TOL)
    # KSD U/V are differences of near-equal sums: tolerance relative to the V-statistic scale
    scale = abs(float(g[f"ksd_v_{tag}"]))
    assert engine.ksd(X, Sx, h, "v").item() == pytest.approx(float(g[f"ksd_v_{tag}"]), rel=TOL)
    assert engine.ksd(X, Sx, h, "u").item() == pytest.approx(float(g[f"ksd_u_{tag}"]), abs=TOL * scale)
    assert engine.ksd(X, Sx, h, "rect", Y, Sy).item() == pytest.approx(float(g[f"ksd_rect_{tag}"]), abs=TOL * scale)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mmd_against_golden(golden, name):
    from nsvgd import engine
    g = golden[name]
    X, Y = dev(g["X"]), dev(g["Y"])
    # MMD^2 is a difference of O(1) kernel means: tolerance relative to mean(K) ~ O(1)
    assert engine.mmd(X, Y, 1.0).item() == pytest.approx(float(g["mmd_rbf"]), abs=2e-6)
    assert engine.mmd(X, Y, 1.0, funnel=True).item() == pytest.approx(float(g["mmd_funnel"]), abs=2e-6)


PATH_CASES = [
    ("simt", 1000, 3, 1.3), ("simt", 333, 1, 0.7), ("simt", 257, 8, 2.0), ("simt", 300, 16, 3.0), ("simt", 130, 31, 4.0),
    ("generic", 700, 50, 5.0), ("generic", 300, 3, 1.3), ("generic", 129, 200, 12.0), ("generic", 64, 1030, 30.0),
    ("tc", 2048, 128, 8.0), ("tc", 1500, 100, 7.0), ("tc", 1024, 64, 5.5), ("tc", 1111, 50, 5.0), ("tc", 4096, 128, 3.0),
    ("tc", 2048, 128, 30.0), ("tc", 1500, 100, 40.0),      # wide kernels around the adaptive-precision threshold (max|xc|^2/h^2 = 0.21 / 0.09; <= 0.1 -> ONE Gram pass)
    ("large", 1500, 200, 12.0), ("large", 1100, 515, 20.0), ("large", 4500, 130, 9.0), ("large", 2100, 1000, 25.0),
]


def test_large_path_bnn_regime_self_term_dominated():
    """BASELINE configs[4] regime at a size the fp64 oracle finishes in seconds: d = 4096, weights ~ 0.15 N(0, 1), the
    bandwidth of the median heuristic (h^2 ~ |x|^2 / ln n), so every off-diagonal k is ~ e^-7 and the self term k_ii = 1
    carries half of each row.  The diagonal Gram entries accumulate monotonically over 768 MMAs and would pick up the TMEM
    truncation bias (k_ii = 1 - 2e-4) -- the kernel sets them in closed form.  Square (symmetric tiles + mirrored
    transposes, 5 column tiles) against the oracle, and the rectangular form on the same rows."""
    from nsvgd import engine
    rng = np.random.default_rng(11)
    n, d = 1200, 4096
    X = (rng.standard_normal((n, d)) * 0.15 + 0.01).astype(np.float32)
    S = (rng.standard_normal((n, d)) * 0.5).astype(np.float32)
    h = float(np.sqrt(0.5 * 2 * d * 0.15 ** 2 / np.log(n + 1)))
    ref = so.phistar(X, X, S, h)
    Xd, Sd = dev(X), dev(S)
    phi, st = engine.phistar(Xd, Xd, Sd, h, Sq=Sd, stats=True, path="large")
    assert rel_fro(phi.cpu(), ref) < 0.5 * TOL, rel_fro(phi.cpu(), ref)
    assert max_row_rel(phi.cpu(), ref) < TOL
    st = st.cpu().numpy()
    assert st[0] / n == pytest.approx(so.stein_discrepancy_phistar(X, S, h), rel=TOL)
    assert st[1] / n == pytest.approx(so.l2_norm_squared(ref), rel=TOL)
    # the row block of a row-sharded rank: a row range OF the particle array (self pairs at column 300 + i)
    blk = engine.phistar(Xd[300:811], Xd, Sd, h, path="large")
    assert rel_fro(blk.cpu(), ref[300:811]) < 0.5 * TOL, rel_fro(blk.cpu(), ref[300:811])
    Q = (rng.standard_normal((300, d)) * 0.15).astype(np.float32)
    assert rel_fro(engine.phistar(dev(Q), Xd, Sd, h, path="large").cpu(), so.phistar(Q, X, S, h)) < TOL


def test_large_path_variants_agree():
    """the single-CTA (M = 128) kernel and the full-rectangle (non-mirrored) tile enumeration against the default
    (CTA pairs, symmetric tiles); the switches are read once per process, hence the subprocesses."""
    import os, subprocess, sys
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, 'projects-2020-neural-svgd_b200')\n"
        "from nsvgd import engine\n"
        "rng = np.random.default_rng(4)\n"
        "X = torch.tensor((rng.standard_normal((2300, 700)) * 0.3 + 0.2).astype(np.float32)).cuda()\n"
        "S = torch.tensor((rng.standard_normal((2300, 700)) * 0.5).astype(np.float32)).cuda()\n"
        "phi, st = engine.phistar(X, X, S, 7.0, Sq=S, stats=True, path='large')\n"
        "np.save(sys.argv[1], np.concatenate([phi.cpu().numpy().ravel(), st.cpu().numpy().astype(np.float32)]))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for i, env in enumerate(({}, {"NSVGD_LARGE_CTAS": "1"}, {"NSVGD_LARGE_SYM": "0"}, {"NSVGD_LARGE_CTAS": "1", "NSVGD_LARGE_SYM": "0"})):
        out = f"/tmp/nsvgd_large_variant_{os.getpid()}_{i}.npy"
        subprocess.run([sys.executable, "-c", code, out], check=True, cwd=root, env={**os.environ, **env}, timeout=300)
        outs.append(np.load(out)); os.remove(out)
    for o in outs[1:]:
        assert not np.isnan(o).any()
        assert rel_fro(o[:-4], outs[0][:-4]) < 2e-6
        np.testing.assert_allclose(o[-4:], outs[0][-4:], rtol=2e-6)


@pytest.mark.parametrize("path,n,d,h", PATH_CASES)
def test_phi_paths_vs_oracle(path, n, d, h):
    """every phi path against the fp64 oracle on a cloud with non-zero mean (centring exercised),
    square and rectangular, with the fused KSD/l2 statistics."""
    from nsvgd import engine
    X, S = so.gaussian_diag_workload(n, d, seed=n + d, shift=0.5)
    S = (S * 0.1).astype(np.float32)
    m = 77
    Q = (np.random.default_rng(1).standard_normal((m, d)) * 0.9 + 0.4).astype(np.float32)
    Xd, Sd, Qd = dev(X), dev(S), dev(Q)
    ref, ref_aux = so.phistar(X, X, S, h, aux=True)
    # every path serves the [drift, repulsion] aux split (d > 128: the second GEMM of csrc/phi_large.cu then runs against
    # [S | Xc] instead of W = S - Xc/h^2)
    phi, aux, stats = engine.phistar(Xd, Xd, Sd, h, Sq=Sd, aux=True, stats=True, path=path)
    assert rel_fro(phi.cpu(), ref) < TOL, rel_fro(phi.cpu(), ref)
    assert max_row_rel(phi.cpu(), ref) < ROW_TOL, max_row_rel(phi.cpu(), ref)
    assert rel_fro(aux.cpu(), ref_aux) < TOL
    st = stats.cpu().numpy()
    sd_ref = so.stein_discrepancy_phistar(X, S, h)
    assert st[0] / n == pytest.approx(sd_ref, rel=2 * TOL, abs=2 * TOL * abs(so.l2_norm_squared(ref)) ** 0.5)
    assert st[1] / n == pytest.approx(so.l2_norm_squared(ref), rel=2 * TOL)
    got_rect = engine.phistar(Qd, Xd, Sd, h, path=path)
    assert rel_fro(got_rect.cpu(), so.phistar(Q, X, S, h)) < TOL
    # without the aux output the tensor-core path runs its COMBINED form (second GEMM against W = S - Xc/h^2);
    # the KSD statistic is then recovered from qc.(K W) + sum_j k g_ij / h^2: square and rectangular
    phi2, st2 = engine.phistar(Xd, Xd, Sd, h, Sq=Sd, stats=True, path=path)
    assert rel_fro(phi2.cpu(), ref) < TOL and max_row_rel(phi2.cpu(), ref) < ROW_TOL, max_row_rel(phi2.cpu(), ref)
    st2 = st2.cpu().numpy()
    assert st2[0] / n == pytest.approx(sd_ref, rel=2 * TOL, abs=2 * TOL * abs(so.l2_norm_squared(ref)) ** 0.5)
    assert st2[1] / n == pytest.approx(so.l2_norm_squared(ref), rel=2 * TOL)
    Sq = (np.random.default_rng(2).standard_normal((m, d)) * 0.3).astype(np.float32)
    phi3, st3 = engine.phistar(Qd, Xd, Sd, h, Sq=dev(Sq), stats=True, path=path)
    ref3 = so.phistar(Q, X, S, h)
    t_ref = float(np.sum(np.sum(Sq.astype(np.float64) * ref3, axis=1) + so.div_phistar(Q, X, S, h)))
    scale = float(np.sum(np.abs(np.sum(Sq.astype(np.float64) * ref3, axis=1)) + np.abs(so.div_phistar(Q, X, S, h))))
    assert rel_fro(phi3.cpu(), ref3) < TOL
    assert st3.cpu().numpy()[0] == pytest.approx(t_ref, abs=2 * TOL * scale)


def test_phi_far_from_origin_and_tight_cloud():
    """cancellation guard: cloud far from the origin, Gram-trick paths must centre."""
    from nsvgd import engine
    rng = np.random.default_rng(3)
    n, d = 1536, 128
    X = (rng.standard_normal((n, d)) * 0.05 + 30.0).astype(np.float32)
    S = rng.standard_normal((n, d)).astype(np.float32)
    h = float(so.median_heuristic(X[:400], "offdiag"))
    ref = so.phistar(X, X, S, h)
    for path in ("generic", "tc"):
        got = engine.phistar(dev(X), dev(X), dev(S), h, path=path)
        assert rel_fro(got.cpu(), ref) < TOL, (path, rel_fro(got.cpu(), ref))


def test_phi_nan_propagates():
    from nsvgd import engine
    X, S = so.gaussian_diag_workload(300, 4, seed=1)
    X[7, 1] = np.nan
    phi = engine.phistar(dev(X), dev(X), dev(S), 1.0)
    assert torch.isnan(phi).any()


def test_step_sgd_scaled_and_unscaled():
    from nsvgd import engine
    X, S = so.gaussian_diag_workload(400, 6, seed=9)
    S = (S * 0.05).astype(np.float32)
    for scaled in (False, True):
        Xd = dev(X).clone()
        aux, grads = engine.svgd_step_sgd(Xd, dev(S), 1.4, lr=0.05, scaled=scaled, lambda_reg=0.5, want_grads=True)
        newX, g_ref, _, extras = so.svgd_step_sgd(X, S, 0.05, bandwidth=1.4, scaled=scaled, lambda_reg=0.5)
        assert rel_fro(grads.cpu(), g_ref) < 2 * TOL
        assert rel_fro(Xd.cpu(), newX) < 1e-6
        a = aux.cpu().numpy()
        if scaled:
            assert a[2] == pytest.approx(extras["alpha"], rel=2 * TOL)
            assert a[0] == pytest.approx(extras["ksd"], rel=2 * TOL)
        assert a[3] == pytest.approx(float((g_ref ** 2).sum()), rel=4 * TOL)


def test_full_size_rows_sampled_against_oracle():
    """BASELINE size (n=65536, d=128): 48 sampled rows of phi against the fp64 oracle evaluated on
    those rows only (rectangular form), plus permutation equivariance on the device."""
    from nsvgd import engine
    n, d = 65536, 128
    X, S = so.gaussian_diag_workload(n, d, seed=0)
    S = (S * 0.02).astype(np.float32)
    Xd, Sd = dev(X), dev(S)
    h = 6.0
    phi = engine.phistar(Xd, Xd, Sd, h)
    rows = np.random.default_rng(0).choice(n, 48, replace=False)
    ref = so.phistar(X[rows], X, S, h)
    assert rel_fro(phi[rows].cpu(), ref) < TOL
    assert max_row_rel(phi[rows].cpu(), ref) < ROW_TOL, max_row_rel(phi[rows].cpu(), ref)
    perm = torch.randperm(n, device="cuda")
    phi_p = engine.phistar(Xd[perm].contiguous(), Xd[perm].contiguous(), Sd[perm].contiguous(), h)
    assert rel_fro(phi_p.cpu(), phi[perm].cpu()) < 2e-5


def _median_dataset(kind, n, d, seed):
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        return (rng.standard_normal((n, d)) * 1.3 + 0.5).astype(np.float32)
    if kind == "far":          # tight cloud far from the origin: centring + error radius must hold
        return (rng.standard_normal((n, d)) * 0.05 + 30.0).astype(np.float32)
    if kind == "cluster":      # well separated tight clusters: bimodal distance distribution
        c = rng.standard_normal((8, d)) * 6.0
        return (c[rng.integers(0, 8, n)] + rng.standard_normal((n, d)) * 0.02).astype(np.float32)
    if kind == "ties":         # every row duplicated many times: massive ties -> device-side fallback
        base = (rng.standard_normal((16, d))).astype(np.float32)
        return base[rng.integers(0, 16, n)]
    raise ValueError(kind)


@pytest.mark.parametrize("n,d,kind", [(8192, 128, "gauss"), (8192, 64, "gauss"), (9000, 100, "far"), (8192, 40, "cluster"),
                                      (8200, 128, "ties"), (12000, 128, "gauss"), (8192, 33, "far")])
@pytest.mark.parametrize("mode", so.MEDIAN_MODES)
def test_fast_median_bit_identical_to_simt_select(n, d, kind, mode, monkeypatch):
    """tensor-core filtered select == 3-pass SIMT select, bit for bit, on benign and adversarial clouds"""
    from nsvgd import engine
    X = dev(_median_dataset(kind, n, d, seed=n + d))
    monkeypatch.setenv("NSVGD_MEDIAN_FAST", "0")
    slow = engine.median_heuristic(X, mode).cpu().numpy()[0]
    monkeypatch.setenv("NSVGD_MEDIAN_FAST", "1")
    fast = engine.median_heuristic(X, mode).cpu().numpy()[0]
    assert np.float32(fast).tobytes() == np.float32(slow).tobytes(), (fast, slow)


@pytest.mark.parametrize("mode", ["literal", "offdiag"])
def test_fast_median_bit_exact_vs_oracle_n8192(mode):
    from nsvgd import engine
    X = _median_dataset("gauss", 8192, 128, seed=5)
    want = so.median_heuristic(X, mode)
    got = engine.median_heuristic(dev(X), mode).cpu().numpy()[0]
    assert np.float32(got).tobytes() == np.float32(want).tobytes(), (got, want)


def test_fast_median_nan():
    from nsvgd import engine
    X = _median_dataset("gauss", 8192, 128, seed=6)
    X[4000, 17] = np.nan
    assert np.isnan(engine.median_heuristic(dev(X), "offdiag").item())


@pytest.mark.parametrize("mode", ["literal", "offdiag"])
def test_fast_median_staged_two_ranks_on_one_gpu(mode):
    """the staged (row-sharded) form of the tensor-core filtered select: two simulated ranks on one
    GPU, reduce tensors and histograms summed by hand exactly as parallel.py all-reduces them."""
    from nsvgd import engine
    X = dev(_median_dataset("gauss", 8192, 128, seed=21))
    want = engine.median_heuristic(X, mode).cpu().numpy()[0]
    sels = [engine.MedianSelect(X, mode) for _ in range(2)]
    assert all(s.fast_supported for s in sels)
    for r, s in enumerate(sels):                      # each simulated rank samples its half ...
        s.fast_sample(r, 2)
    full0, _ = sels[0].fast_sample(0, 2); full1, _ = sels[1].fast_sample(1, 2)
    half = full0.numel() // 2                          # ... and the halves are exchanged (all-gather)
    full0[half:].copy_(full1[half:]); full1[:half].copy_(full0[:half])
    red = sum(s.fast_stage1(r, 2).clone() for r, s in enumerate(sels))
    for s in sels:
        s.reduce.copy_(red); s.fast_resolve()
    for p in range(3):
        hist = sum(s.fast_cand_hist(p).clone() for s in sels)
        for s in sels:
            s.hist.copy_(hist); s.fast_cand_advance(p)
    for s in sels:
        s.fast_commit()
    for p in range(3):
        hist = sum(s.fallback_hist(p, r, 2).clone() for r, s in enumerate(sels))
        for s in sels:
            s.hist_fb.copy_(hist); s.fallback_advance(p)
    for s in sels:
        got = s.finish().cpu().numpy()[0]
        assert np.float32(got).tobytes() == np.float32(want).tobytes()
    assert red[1].item() > 0 and red[2].item() == 0
