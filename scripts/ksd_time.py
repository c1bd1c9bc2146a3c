import os, sys, numpy as np, torch
sys.path[:0] = ['.', 'projects-2020-neural-svgd_b200']
from nsvgd import engine
import bench
X, S = bench.workload(65536, 128)
Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()
for env in ("1", "0"):
    os.environ["NSVGD_KSD_TC"] = env
    for v in ("u", "v"):
        r = engine.ksd(Xd, Sd, 6.0, v); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = engine.ksd(Xd, Sd, 6.0, v); b.record(); torch.cuda.synchronize()
        print(f"cfg4 ksd_{v} NSVGD_KSD_TC={env}: {a.elapsed_time(b):.2f} ms value {r.item():.9g}", flush=True)
