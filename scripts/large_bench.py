"""d > 128 path (csrc/phi_large.cu): on-GPU cross-check against the generic SIMT path at multi-tile / multi-chunk shapes and
the cfg5 timing (n = 16384, d = 4096).  Usage: python scripts/large_bench.py  (NSVGD_LARGE_CTAS=1|2 selects the variant)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "projects-2020-neural-svgd_b200"))
from nsvgd import engine

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
print("NSVGD_LARGE_CTAS =", os.environ.get("NSVGD_LARGE_CTAS", "2 (default)"), flush=True)
for n, m, d, h in ((4096, 4096, 512, 6.0), (9000, 9000, 300, 5.0), (5000, 700, 1030, 9.0)):
    X = torch.tensor((rng.standard_normal((n, d)) * 0.3 + 0.2).astype(np.float32), device=dev)
    S = torch.tensor((rng.standard_normal((n, d)) * 0.5).astype(np.float32), device=dev)
    if m == n:
        a, sa = engine.phistar(X, X, S, h, Sq=S, stats=True, path="large")
        b, sb = engine.phistar(X, X, S, h, Sq=S, stats=True, path="generic")
        st = (sa.cpu().numpy(), sb.cpu().numpy())
    else:
        Q = torch.tensor((rng.standard_normal((m, d)) * 0.3 + 0.1).astype(np.float32), device=dev)
        a = engine.phistar(Q, X, S, h, path="large")
        b = engine.phistar(Q, X, S, h, path="generic")
        st = None
    torch.cuda.synchronize()
    err = float((a - b).norm() / b.norm())
    rowerr = float(((a - b).norm(dim=1) / b.norm(dim=1)).max())
    print(f"n={n} m={m} d={d}: large vs generic rel_fro {err:.2e} max_row {rowerr:.2e} nan={bool(torch.isnan(a).any())}", st, flush=True)

n, d = 16384, 4096
rng = np.random.default_rng(5)
X = torch.tensor((rng.standard_normal((n, d)) * 0.15).astype(np.float32), device=dev)
S = torch.tensor((rng.standard_normal((n, d)) * 0.5).astype(np.float32), device=dev)
h = float(np.sqrt(0.5 * 2 * d * 0.15 ** 2 / np.log(n + 1)))
for _ in range(2):
    phi, stats = engine.phistar(X, X, S, h, Sq=S, stats=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    phi, stats = engine.phistar(X, X, S, h, Sq=S, stats=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"cfg5 n={n} d={d} h={h:.4f}: {ms:.3f} ms per phi* call, algorithmic {6.0*n*n*d/ms/1e9:.1f} TFLOP/s, executed "
      f"{12.0*n*n*d/ms/1e9:.1f} TFLOP/s; nan={bool(torch.isnan(phi).any())} stats={stats.cpu().numpy()}", flush=True)
# a slice of cfg5 against float64 (torch, closed form of stein.py:180-222 with the RBF kernel) and the generic path
Xd, Sd = X.double(), S.double()
Qd = Xd[:512]
d2 = (Qd * Qd).sum(1)[:, None] + (Xd * Xd).sum(1)[None, :] - 2.0 * Qd @ Xd.T
K = torch.exp(-d2.clamp_min(0) / (2 * h * h))
ref64 = (K @ Sd + (K.sum(1)[:, None] * Qd - K @ Xd) / (h * h)) / n
print("cfg5 slice (512 rows) large vs float64 rel_fro", float((phi[:512].double() - ref64).norm() / ref64.norm()),
      "max_row", float(((phi[:512].double() - ref64).norm(dim=1) / ref64.norm(dim=1)).max()), flush=True)
del Xd, Sd, d2, K
ref = engine.phistar(X[:512].contiguous(), X, S, h, path="generic")
got = engine.phistar(X[:512].contiguous(), X, S, h, path="large")
print("cfg5 slice (512 rows) large(rect) vs large(square) rel_fro", float((got - phi[:512]).norm() / got.norm()),
      "generic vs float64", float((ref.double() - ref64).norm() / ref64.norm()), flush=True)

# the exact median bandwidth at cfg5 (literal mode): one exact pass + two buffer passes vs three recomputing passes
from nsvgd import _lib
for name, env in (("tensor-core filter", {}), ("one exact pass + stored distances", {"NSVGD_MEDIAN_LARGE": "0"}),
                  ("three exact passes", {"NSVGD_MEDIAN_LARGE": "0", "NSVGD_MEDIAN_STORE": "0"})):
    for k in ("NSVGD_MEDIAN_LARGE", "NSVGD_MEDIAN_STORE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for mode in ("literal", "offdiag"):
        hh = engine.median_heuristic(X, mode)
        torch.cuda.synchronize()
        _lib.load().nsvgd_profile_enable(1)
        a.record()
        hh = engine.median_heuristic(X, mode)
        b.record(); torch.cuda.synchronize()
        ph = {nm: round(_lib.profile_read(tag)[0], 3) for nm, tag in (("prep", 8), ("sample", 4), ("bracket", 9), ("filter", 6), ("exact", 5),
                                                                    ("resolve", 10), ("fallback", 11))}
        _lib.load().nsvgd_profile_enable(0)
        print(f"cfg5 median_heuristic [{name}] {mode}: {a.elapsed_time(b):.2f} ms, h = {hh.item():.9g}  phases(ms) {ph}", flush=True)

