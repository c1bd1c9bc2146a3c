"""GPU diagnostic: phi* error of the tcgen05 path at n=65536, d=128 versus the number of j-splits
(sequential TMEM accumulations per CTA), split into drift and repulsion parts."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")]
from nsvgd import engine
from oracle import svgd_oracle as so

n, d = int(os.environ.get("N", 65536)), 128
X, S = so.gaussian_diag_workload(n, d, seed=0)
S = (S * 0.02).astype(np.float32)
Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()
rows = np.random.default_rng(0).choice(n, 48, replace=False)
Q = torch.tensor(X[rows]).cuda()
for h in (6.0, 12.0, 3.0):
    ref, ref_aux = so.phistar(X[rows], X, S, h, aux=True)
    def err(a, b):
        return np.linalg.norm(a - b) / np.linalg.norm(b)
    for ns in ("1", "2", "4", "8", "16", "32"):
        os.environ["NSVGD_TC_NSPLIT"] = ns
        phi, aux = engine.phistar(Q, Xd, Sd, h, aux=True, path="tc")
        a = aux.cpu().numpy().astype(np.float64)
        print(f"h={h} nsplit={ns}: phi {err(phi.cpu().numpy(), ref):.3e} drift {err(a[:,0], ref_aux[:,0]):.3e} "
              f"rep {err(a[:,1], ref_aux[:,1]):.3e}", flush=True)
    os.environ.pop("NSVGD_TC_NSPLIT")
    phi, aux = engine.phistar(Q, Xd, Sd, h, aux=True, path="generic")
    a = aux.cpu().numpy().astype(np.float64)
    print(f"h={h} generic : phi {err(phi.cpu().numpy(), ref):.3e} drift {err(a[:,0], ref_aux[:,0]):.3e} rep {err(a[:,1], ref_aux[:,1]):.3e}", flush=True)
