#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --tb=short -p no:cacheprovider -k "phi_paths and (large or generic)" 2>&1 | tail -12
timeout 300 python - <<'PY' 2>&1 | tail -8
import sys, time, numpy as np, torch
sys.path[:0] = ['.', 'projects-2020-neural-svgd_b200']
from nsvgd import engine
from oracle import svgd_oracle as so
n, d = 16384, 4096
rng = np.random.default_rng(0)
X = (rng.standard_normal((n, d)) * 0.15).astype(np.float32)          # nets.py:276 init scale
S = (rng.standard_normal((n, d)) * 0.5).astype(np.float32)
Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()
h = 6.0
for path in ("large", "generic"):
    phi, st = engine.phistar(Xd, Xd, Sd, h, Sq=Sd, stats=True, path=path)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        phi, st = engine.phistar(Xd, Xd, Sd, h, Sq=Sd, stats=True, path=path)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rows = rng.choice(n, 24, replace=False)
    ref = so.phistar(X[rows], X, S, h)
    err = np.linalg.norm(phi[rows].cpu().numpy() - ref) / np.linalg.norm(ref)
    print(f"cfg5 n={n} d={d} path={path}: {ms:.2f} ms  ({6.0*n*n*d/ms/1e9:.0f} algorithmic TFLOP/s)  rel err vs fp64 oracle (24 rows) {err:.2e}  stats {st.cpu().numpy()[:2]}")
PY
