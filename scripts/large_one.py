"""one cfg5 phi* call (for ncu): python scripts/large_one.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "projects-2020-neural-svgd_b200"))
from nsvgd import engine
dev = torch.device("cuda:0")
n, d = 16384, 4096
rng = np.random.default_rng(5)
X = torch.tensor((rng.standard_normal((n, d)) * 0.15).astype(np.float32), device=dev)
S = torch.tensor((rng.standard_normal((n, d)) * 0.5).astype(np.float32), device=dev)
h = float(np.sqrt(0.5 * 2 * d * 0.15 ** 2 / np.log(n + 1)))
for _ in range(int(os.environ.get("REPS", "2"))):
    phi, stats = engine.phistar(X, X, S, h, Sq=S, stats=True)
torch.cuda.synchronize()
print(stats.cpu().numpy())
