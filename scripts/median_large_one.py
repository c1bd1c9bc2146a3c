"""one cfg5 median (for ncu): python scripts/median_large_one.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "projects-2020-neural-svgd_b200"))
from nsvgd import engine
dev = torch.device("cuda:0")
n, d = 16384, 4096
rng = np.random.default_rng(5)
X = torch.tensor((rng.standard_normal((n, d)) * 0.15).astype(np.float32), device=dev)
for _ in range(2):
    h = engine.median_heuristic(X, "offdiag")
torch.cuda.synchronize()
print(h.item())
