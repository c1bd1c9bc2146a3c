"""accuracy of the d > 128 path in the self-term-dominated regime vs torch float64 on the GPU"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "projects-2020-neural-svgd_b200"))
from nsvgd import engine
dev = torch.device("cuda:0")

def ref64(Q, X, S, h):
    Qd, Xd, Sd = Q.double(), X.double(), S.double()
    d2 = (Qd * Qd).sum(1)[:, None] + (Xd * Xd).sum(1)[None, :] - 2.0 * Qd @ Xd.T
    K = torch.exp(-d2.clamp_min(0) / (2 * h * h))
    return (K @ Sd + (K.sum(1)[:, None] * Qd - K @ Xd) / (h * h)) / X.shape[0]

print("env", {k: v for k, v in os.environ.items() if k.startswith("NSVGD_")}, flush=True)
for n, d, shift, hs in ((1200, 4096, 0.01, 1.0), (1200, 4096, 0.0, 1.0), (1280, 4096, 0.0, 1.0), (1200, 4096, 0.01, 2.0), (1200, 2048, 0.01, 1.0),
                        (4096, 4096, 0.0, 1.0)):
    rng = np.random.default_rng(11)
    X = torch.tensor((rng.standard_normal((n, d)) * 0.15 + shift).astype(np.float32), device=dev)
    S = torch.tensor((rng.standard_normal((n, d)) * 0.5).astype(np.float32), device=dev)
    h = hs * float(np.sqrt(0.5 * 2 * d * 0.15 ** 2 / np.log(n + 1)))
    r = ref64(X, X, S, h)
    for path in ("large", "generic"):
        phi = engine.phistar(X, X, S, h, path=path).double()
        e = float((phi - r).norm() / r.norm())
        rows = ((phi - r).norm(dim=1) / r.norm(dim=1))
        print(f"n={n} d={d} shift={shift} h={h:.3f} {path}: rel_fro {e:.3e} max_row {float(rows.max()):.3e} argmax {int(rows.argmax())}", flush=True)
