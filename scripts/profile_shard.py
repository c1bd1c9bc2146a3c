"""One GPU, the per-rank work of a row-sharded step: rank 0 of a VIRTUAL world of G ranks (collectives replaced
by local stand-ins: gather = the full array, all-reduce = identity).  Results are meaningless; kernel shapes and
launch sequence are those of a real G-rank step, so `ncu --metrics gpu__time_duration.sum` on this script ranks the
per-rank kernels without an N-GPU box.   python scripts/profile_shard.py G [steps]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")]
import torch.distributed as dist
from nsvgd import parallel
from nsvgd.parallel import ShardedSVGD, row_partition
from oracle.svgd_oracle import gaussian_diag_workload

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, d = 65536, 128
X, S = gaussian_diag_workload(n, d, seed=0)
S = (S * 0.02).astype(np.float32)
Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()


class Virtual(ShardedSVGD):
    def __init__(self):
        self.world, self.rank = G, 0
        from nsvgd import engine
        self.engine, self.n, self.d, self.group, self.median_mode = engine, n, d, None, "literal"
        self.r0, self.r1 = row_partition(n, G, 0)
        self.max_rows = row_partition(n, G, 0)[1]
        self._buf = {}

    def all_gather_rows(self, local, tag):
        return Xd if tag == "X" else Sd


# collectives -> local no-ops
dist.all_reduce = lambda t, **k: t
dist.all_gather_into_tensor = lambda out, inp, **k: out
sh = Virtual()
Xl, Sl = Xd[sh.r0:sh.r1].clone(), Sd[sh.r0:sh.r1].clone()
for _ in range(steps):
    sh.step(Xl, Sl, 1e-3, scaled=True, lambda_reg=0.5)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(steps):
    sh.step(Xl, Sl, 1e-3, scaled=True, lambda_reg=0.5)
ev1.record(); torch.cuda.synchronize()
print(f"virtual rank 0 of {G}: {ev0.elapsed_time(ev1) / steps:.3f} ms per step (no collectives)")
