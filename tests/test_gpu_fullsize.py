"""GPU parity at BASELINE size, on the branch bench.py times.

bench.workload(65536, 128) is the benchmark's own cloud.  Checked here, through the C ABI:
  (a) the tensor-core filtered median == the 3-pass SIMT select, bit for bit, at n = 65536 (literal and offdiag);
  (b) one fused SVGD step (phi*, KSD, l2, alpha, update) at the literal-mode bandwidth (-> ONE fp16 Gram pass) and at the
      off-diagonal bandwidth (-> three bf16-split passes) against the fp64 oracle: 96 sampled rows at 1 x TOL row-wise,
      the statistics over all n rows;
  (c) adversarial clouds for the one-pass decision: rank-1, two clusters at +-u, every coordinate of the form
      (1 + 3 * 2^-e) * 2^k so that all bf16 (e = 9) or all fp16 (e = 12) rounding residuals share one sign -- the
      Cauchy-Schwarz worst case of a single-piece Gram.
Tolerance: 1e-4 relative (north_star), row-wise as well as in Frobenius norm."""
import os

import numpy as np
import pytest
import torch

from oracle import svgd_oracle as so

pytestmark = pytest.mark.gpu

TOL = 1e-4
N, D = 65536, 128


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


@pytest.fixture(scope="module")
def cloud():
    import bench
    X, S = bench.workload(N, D)
    return X, S


@pytest.fixture(scope="module")
def bandwidths(cloud):
    """fast (tensor-core filtered) and slow (SIMT 3-pass select) bandwidths of the benchmark cloud, both modes"""
    from nsvgd import engine
    Xd = dev(cloud[0])
    out = {}
    for mode in ("literal", "offdiag"):
        os.environ["NSVGD_MEDIAN_FAST"] = "1"
        fast = engine.median_heuristic(Xd, mode).cpu().numpy()[0]
        os.environ["NSVGD_MEDIAN_FAST"] = "0"
        try:
            slow = engine.median_heuristic(Xd, mode).cpu().numpy()[0]
        finally:
            os.environ["NSVGD_MEDIAN_FAST"] = "1"
        out[mode] = (fast, slow)
    return out


@pytest.mark.parametrize("mode", ["literal", "offdiag"])
def test_fullsize_fast_median_equals_simt_select(bandwidths, mode):
    fast, slow = bandwidths[mode]
    assert np.isfinite(fast) and fast > 0
    assert np.float32(fast).tobytes() == np.float32(slow).tobytes(), (mode, fast, slow)


@pytest.mark.parametrize("mode", ["literal", "offdiag"])
def test_fullsize_step_vs_fp64_oracle(cloud, bandwidths, mode):
    """the benchmarked call (nsvgd_step_sgd_f32, scaled) at the bandwidth the benchmark's median produces"""
    from nsvgd import engine
    X, S = cloud
    h = float(bandwidths[mode][0])
    rows = np.random.default_rng(7).choice(N, 96, replace=False)
    ref = so.phistar_stats_f64_blocked(X, S, h, block=512, rows=rows)
    lr = 1e-3
    Xd, Sd = dev(X), dev(S)
    aux, grads = engine.svgd_step_sgd(Xd, Sd, h, lr=lr, scaled=True, lambda_reg=0.5, want_grads=True)
    a = aux.cpu().numpy()
    assert a[0] == pytest.approx(ref["ksd"], rel=TOL), ("ksd", a[0], ref["ksd"])
    assert a[1] == pytest.approx(ref["l2"], rel=TOL), ("l2", a[1], ref["l2"])
    assert a[2] == pytest.approx(ref["alpha"], rel=TOL), ("alpha", a[2], ref["alpha"])
    g_ref = -ref["alpha"] * ref["phi"]
    g = grads[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    assert rel_fro(g, g_ref) < TOL, rel_fro(g, g_ref)
    assert max_row_rel(g, g_ref) < TOL, max_row_rel(g, g_ref)
    x_new = Xd[torch.as_tensor(rows, device="cuda")].cpu().numpy().astype(np.float64)
    assert np.abs(x_new - (X[rows].astype(np.float64) - lr * g_ref)).max() < 1e-6
    # phi* on its own (no statistics): same kernel, rectangular use of the sampled rows as queries
    phi_rect = engine.phistar(dev(X[rows]), dev(X), Sd, h).cpu().numpy()
    assert max_row_rel(phi_rect, ref["phi"]) < TOL


def _coherent_cloud(n, d, e, ratio, rng):
    """rank-1, two balanced clusters at +-u; u_c = (1 + 3 * 2^-e) * 2^k_c: bf16 rounds every coordinate the same way
    for e = 9, fp16 for e = 12.  Returns X and the bandwidth with max|xc|^2 / h^2 = ratio."""
    k = rng.integers(-3, 1, size=d)
    u = ((1.0 + 3.0 * 2.0 ** (-e)) * 2.0 ** k).astype(np.float32)
    sign = np.where(np.arange(n) % 2 == 0, 1.0, -1.0).astype(np.float32)
    X = sign[:, None] * u[None, :]
    h = float(np.sqrt(np.sum(u.astype(np.float64) ** 2) / ratio))
    return np.ascontiguousarray(X), h


@pytest.mark.parametrize("e", [9, 12])
@pytest.mark.parametrize("ratio", [0.099, 0.02, 0.5])
def test_adversarial_coherent_residuals(e, ratio):
    """every rounding residual of the single-piece Gram operand lines up: the error of a one-pass Gram reaches its
    Cauchy-Schwarz bound in the repulsion term.  The device-side decision must keep phi* within 1e-4 regardless."""
    from nsvgd import engine
    rng = np.random.default_rng(e)
    n, d = 4096, 128
    X, h = _coherent_cloud(n, d, e, ratio, rng)
    S = (0.01 * rng.standard_normal((n, d))).astype(np.float32)
    ref = so.phistar_stats_f64_blocked(X, S, h, block=1024)
    phi, stats = engine.phistar(dev(X), dev(X), dev(S), h, Sq=dev(S), stats=True, path="tc")
    phi = phi.cpu().numpy()
    assert rel_fro(phi, ref["phi"]) < TOL, rel_fro(phi, ref["phi"])
    assert max_row_rel(phi, ref["phi"]) < TOL, max_row_rel(phi, ref["phi"])
    st = stats.cpu().numpy()
    assert st[1] / n == pytest.approx(ref["l2"], rel=2 * TOL)
    assert st[0] / n == pytest.approx(ref["ksd"], rel=2 * TOL, abs=2 * TOL * abs(ref["l2"]) ** 0.5)


def test_fullsize_cfg5_bnn_step_pieces():
    """BASELINE configs[4] at full size (n = 16384, d = 4096, weights ~ 0.15 N(0, 1), the bench's `secondary` cloud), the
    d > 128 kernels of csrc/phi_large.cu and csrc/median_large.cu:
      (a) the tensor-core filtered median == one exact SIMT pass + stored distances, bit for bit (literal and offdiag);
      (b) phi* with the KSD / l2 statistics at the median-heuristic-sized bandwidth (self term = half of every row) against
          float64 on 384 query rows (torch on the GPU: the closed form of stein.py:180-222 with the RBF kernel), and the
          block form a row-sharded rank uses (queries = a row range of the particle array)."""
    from nsvgd import engine
    n, d = 16384, 4096
    rng = np.random.default_rng(5)
    X = torch.tensor((rng.standard_normal((n, d)) * 0.15).astype(np.float32)).cuda()
    S = torch.tensor((rng.standard_normal((n, d)) * 0.5).astype(np.float32)).cuda()
    for mode in ("literal", "offdiag"):
        fast = engine.median_heuristic(X, mode).item()
        os.environ["NSVGD_MEDIAN_LARGE"] = "0"
        try:
            slow = engine.median_heuristic(X, mode).item()
        finally:
            del os.environ["NSVGD_MEDIAN_LARGE"]
        assert np.float32(fast).tobytes() == np.float32(slow).tobytes(), (mode, fast, slow)
    h = float(np.sqrt(0.5 * 2 * d * 0.15 ** 2 / np.log(n + 1)))
    phi, stats = engine.phistar(X, X, S, h, Sq=S, stats=True)
    rows = torch.arange(1000, 1384, device="cuda")
    Xd, Sd = X.double(), S.double()
    Qd = Xd[rows]
    d2 = (Qd * Qd).sum(1)[:, None] + (Xd * Xd).sum(1)[None, :] - 2.0 * Qd @ Xd.T
    K = torch.exp(-d2.clamp_min(0) / (2 * h * h))
    K[torch.arange(len(rows)), rows] = 1.0
    ref = ((K @ Sd + (K.sum(1)[:, None] * Qd - K @ Xd) / (h * h)) / n).cpu().numpy()
    got = phi[rows].cpu().numpy()
    assert rel_fro(got, ref) < 0.5 * TOL, rel_fro(got, ref)
    assert max_row_rel(got, ref) < TOL, max_row_rel(got, ref)
    blk = engine.phistar(X[1000:1384], X, S, h).cpu().numpy()
    assert rel_fro(blk, ref) < 0.5 * TOL, rel_fro(blk, ref)
    assert np.isfinite(stats.cpu().numpy()).all() and stats[2].item() == n
    # l2 statistic = sum_i |phi_i|^2 over ALL rows, consistent with the phi the kernel returned
    assert stats[1].item() == pytest.approx(float((phi.double() ** 2).sum().item()), rel=1e-5)
