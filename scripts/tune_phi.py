"""GPU: time the phi_tc kernel variants (pair warps x MMA interleave) at n=65536, d=128 and check them."""
import os, sys, itertools
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")]
from nsvgd import engine, _lib
from oracle import svgd_oracle as so

n, d = 65536, 128
X, S = so.gaussian_diag_workload(n, d, seed=0)
S = (S * 0.02).astype(np.float32)
Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()
rows = np.random.default_rng(0).choice(n, 16, replace=False)
h = 6.0
ref = so.phistar(X[rows], X, S, h)
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
for nw, il in itertools.product(("8", "16"), ("-",)):
    os.environ["NSVGD_TC_PAIR_WARPS"] = nw
    for _ in range(2):
        phi, st = engine.phistar(Xd, Xd, Sd, h, Sq=Sd, stats=True)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(5):
        flush.fill_(1)
        phi, st = engine.phistar(Xd, Xd, Sd, h, Sq=Sd, stats=True)
    torch.cuda.synchronize()
    ms, cnt = _lib.profile_read(0)
    op_ms, _ = _lib.profile_read(2)
    _lib.profile_enable(False)
    err = np.linalg.norm(phi[rows].cpu().numpy() - ref) / np.linalg.norm(ref)
    print(f"pair_warps={nw} interleave={il}: kernel {ms/cnt:.3f} ms, op {op_ms/cnt:.3f} ms, rel err {err:.2e}, ksd {st[0].item()/n:.6f}", flush=True)
