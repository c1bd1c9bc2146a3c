"""Worker of tests/test_p2p_gpu.py: one of two processes that share cuda:0.  gloo carries only python objects
(the cudaIpc handles); everything that is checked runs through the peer-memory kernels of csrc/p2p.cuh."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")]
from nsvgd import engine                                    # noqa: E402
from nsvgd.parallel import ShardedSVGD, row_partition       # noqa: E402
from oracle.svgd_oracle import gaussian_diag_workload       # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(0)
dist.init_process_group("gloo", rank=rank, world_size=world)
ok = True

sh = ShardedSVGD(8192, 128, median_mode="literal", p2p="on")
comm = sh.comm
# 1. stand-alone all-reduce: integers, doubles in rank order, repeated (epoch / slot alternation)
for it in range(5):
    t = torch.arange(6000, dtype=torch.int64, device="cuda") * (rank + 1) + it
    comm.allreduce(t)
    exp = torch.arange(6000, dtype=torch.int64, device="cuda") * (world * (world + 1) // 2) + world * it
    f = torch.full((4,), 0.1 * (rank + 1) + it, dtype=torch.float64, device="cuda")
    comm.allreduce(f)
    fe = 0.0
    for r in range(world):
        fe += 0.1 * (r + 1) + it
    ok = ok and bool((t == exp).all()) and f[0].item() == fe
print(f"rank {rank} allreduce ok={ok}", flush=True)

# 2. row-sharded exact median with the fused exchanges == single-process median, bit for bit (tensor-core filtered
#    select at n=8192, SIMT select at n=3000)
for n, d, mode in ((8192, 128, "literal"), (3000, 7, "offdiag")):
    X, S = gaussian_diag_workload(n, d, seed=11)
    Xd = torch.tensor(X, device="cuda")
    s2 = ShardedSVGD(n, d, median_mode=mode, p2p="off")
    s2.comm = comm
    h = s2.median_heuristic(Xd)
    h_ref = engine.median_heuristic(Xd, mode)
    same = h.item() == h_ref.item()
    print(f"rank {rank} n={n} {mode}: h={h.item():.7f} bit-identical={same}", flush=True)
    ok = ok and same

# 3. phi* of the local rows + fused all-reduce of the KSD/l2 sums + update == the single-process step
n, d = 8192, 128
X, S = gaussian_diag_workload(n, d, seed=12)
S = (S * 0.02).astype(np.float32)
Xd, Sd = torch.tensor(X, device="cuda"), torch.tensor(S, device="cuda")
h = engine.median_heuristic(Xd, "literal")
Xs = Xd.clone()
aux_ref, _ = engine.svgd_step_sgd(Xs, Sd, h, 1e-2, scaled=True, lambda_reg=0.5)
r0, r1 = row_partition(n, world, rank)
Xl = Xd[r0:r1].clone()
phi, stats = engine.phistar(Xl, Xd, Sd, h, Sq=Sd[r0:r1].contiguous(), stats=True)
aux, _ = engine.apply_update_p2p(comm, Xl, phi, stats, 1e-2, True, 0.5)
err = (Xl - Xs[r0:r1]).norm().item() / Xs[r0:r1].norm().item()
a_rel = abs(aux[2].item() - aux_ref[2].item()) / abs(aux_ref[2].item())
print(f"rank {rank} step: X rel err={err:.2e} alpha rel err={a_rel:.2e}", flush=True)
ok = ok and err < 2e-6 and a_rel < 1e-5 and not comm.timed_out()
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0 if ok else 1)
