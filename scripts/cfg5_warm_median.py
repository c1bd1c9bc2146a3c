import sys, numpy as np, torch
sys.path[:0]=[".","projects-2020-neural-svgd_b200"]
from nsvgd import engine
n,d=16384,4096
rng=np.random.default_rng(5)
X=torch.tensor((rng.standard_normal((n,d))*0.15).astype(np.float32)).cuda()
V=torch.tensor((rng.standard_normal((n,d))*1e-6).astype(np.float32)).cuda()
for mode in ("literal","offdiag"):
    hint=engine.new_median_hint("cuda")
    ts=[]
    for i in range(5):
        X+=V
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); h=engine.median_heuristic(X,mode,hint=hint); b.record(); torch.cuda.synchronize(); ts.append(round(a.elapsed_time(b),2))
    cold=engine.median_heuristic(X,mode)
    print(mode, "ms per call", ts, "stats", {k:v for k,v in engine.median_hint_stats(hint).items() if k in ("hits","misses","delta")}, "bit-identical", h.item()==cold.item(), flush=True)
