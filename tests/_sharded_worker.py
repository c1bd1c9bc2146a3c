"""Worker of tests/test_sharded_gpu.py (torchrun, NCCL, one rank per GPU): 2-rank NCCL check: the row-sharded step reproduces the single-GPU step (h bit-identical, X to 1e-6)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")]
from nsvgd import engine
from nsvgd.parallel import ShardedSVGD, row_partition
from oracle.svgd_oracle import gaussian_diag_workload

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
ok = True
# (2304, 192): d > 128 -- the rank's block is a row range of the gathered array (self pairs known), the median is the
# large-d filtered select run redundantly on every rank
for n, d, mode in ((8192, 128, "literal"), (3000, 7, "offdiag"), (9001, 100, "full"), (2304, 192, "literal"), (2500, 260, "offdiag")):
    X, S = gaussian_diag_workload(n, d, seed=3)
    S = (S * 0.02).astype(np.float32)
    Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()
    h_ref = engine.median_heuristic(Xd, mode)
    Xs = Xd.clone()
    engine.svgd_step_sgd(Xs, Sd, h_ref, 1e-2, scaled=True, lambda_reg=0.5)
    r0, r1 = row_partition(n, world, rank)
    Xl, Sl = Xd[r0:r1].clone(), Sd[r0:r1].clone()
    sh = ShardedSVGD(n, d, median_mode=mode)
    h, aux, _ = sh.step(Xl, Sl, 1e-2, scaled=True, lambda_reg=0.5)
    same_h = h.item() == h_ref.item()
    err = (Xl - Xs[r0:r1]).norm().item() / Xs[r0:r1].norm().item()
    p2p = sh.comm is not None
    tmo = sh.comm.timed_out() if p2p else False
    print(f"rank {rank} n={n} d={d} {mode}: h bit-identical={same_h} ({h.item():.7f}) X rel err={err:.2e} peer-memory={p2p} timeout={tmo}", flush=True)
    ok = ok and not tmo
    ok = ok and same_h and err < 2e-6
# the same step captured as ONE CUDA graph (collectives included), replayed three times
n, d, mode = 8192, 128, "literal"
X, S = gaussian_diag_workload(n, d, seed=5)
S = (S * 0.02).astype(np.float32)
Xd, Sd = torch.tensor(X).cuda(), torch.tensor(S).cuda()
Xs = Xd.clone()
hs = []
for _ in range(3):
    hs.append(engine.median_heuristic(Xs, mode))
    engine.svgd_step_sgd(Xs, Sd, hs[-1], 1e-2, scaled=True, lambda_reg=0.5)
r0, r1 = row_partition(n, world, rank)
Xl, Sl = Xd[r0:r1].clone(), Sd[r0:r1].clone()
sh = ShardedSVGD(n, d, median_mode=mode)
g = sh.capture(Xl, Sl, 1e-2, scaled=True, lambda_reg=0.5)
Xl.copy_(Xd[r0:r1])
same = True
for i in range(3):
    h, aux = g.replay()
    same = same and (h.item() == hs[i].item())
err = (Xl - Xs[r0:r1]).norm().item() / Xs[r0:r1].norm().item()
print(f"rank {rank} graphed x3: h bit-identical={same} X rel err={err:.2e} peer-memory={sh.comm is not None}", flush=True)
if sh.comm is not None:
    # stand-alone peer-memory all-reduce: integers and doubles (fixed rank order)
    t = torch.arange(5000, dtype=torch.int64, device="cuda") * (rank + 1)
    sh.comm.allreduce(t)
    exp = torch.arange(5000, dtype=torch.int64, device="cuda") * (world * (world + 1) // 2)
    f = torch.full((7,), 0.1 * (rank + 1), dtype=torch.float64, device="cuda")
    sh.comm.allreduce(f)
    fe = sum(0.1 * (r + 1) for r in range(world))
    good = bool((t == exp).all()) and abs(f[0].item() - fe) < 1e-15 and not sh.comm.timed_out()
    print(f"rank {rank} p2p allreduce int/double ok={good}", flush=True)
    ok = ok and good
ok = ok and same and err < 2e-6
torch.cuda.synchronize()
del g
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)      # destroy_process_group() hangs after NCCL graph replays (torch 2.11 / NCCL 2.28)
