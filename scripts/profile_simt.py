"""small-d SIMT path (north star: "at small d the kernel is a SIMT FP32/MUFU-exp kernel"): phi* + KSD statistics at
d=2 (funnel) for ncu:  ncu --set full -k regex:phi_smalld -c 1 python scripts/profile_simt.py"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")]
from nsvgd import engine
from oracle import svgd_oracle as so
n, d = 65536, 2
rng = np.random.default_rng(0)
X = (rng.standard_normal((n, d)) * np.array([1.5, 1.0])).astype(np.float32)
S = so.funnel_score(X).astype(np.float32)
Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()
for _ in range(3):
    phi, st = engine.phistar(Xd, Xd, Sd, 1.0, Sq=Sd, stats=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    phi, st = engine.phistar(Xd, Xd, Sd, 1.0, Sq=Sd, stats=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
rows = rng.choice(n, 32, replace=False)
ref = so.phistar(X[rows], X, S, 1.0)
err = np.linalg.norm(phi[rows].cpu().numpy() - ref) / np.linalg.norm(ref)
print(f"simt d=2 n={n}: {ms:.3f} ms per phi* = {n * n / ms / 1e6:.0f} Gpairs/s, rel err {err:.1e}")
