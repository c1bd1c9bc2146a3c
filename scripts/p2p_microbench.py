"""torchrun microbenchmark: latency of the peer-memory all-reduce (csrc/p2p.cuh) vs NCCL for the small vectors the
row-sharded median exchanges, and of the 16 MB sample all-gather."""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")]
from nsvgd import engine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
comm = engine.PeerComm()
def timeit(fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for nw in (4, 2048, 4096, 8192):
    t = torch.ones(nw, dtype=torch.int64, device="cuda")
    us_p2p = timeit(lambda: comm.allreduce(t))
    us_nccl = timeit(lambda: dist.all_reduce(t))
    if rank == 0: print(f"allreduce {nw} words: p2p {us_p2p:.1f} us  nccl {us_nccl:.1f} us", flush=True)
full = torch.empty(1 << 22, dtype=torch.float32, device="cuda"); mine = full[rank * (1 << 22) // world:(rank + 1) * (1 << 22) // world]
us = timeit(lambda: dist.all_gather_into_tensor(full, mine), 50)
if rank == 0: print(f"all_gather 16 MB: {us:.1f} us", flush=True)
X = torch.empty(65536, 128, device="cuda"); Xl = X[rank * 65536 // world:(rank + 1) * 65536 // world]
us = timeit(lambda: dist.all_gather_into_tensor(X, Xl), 50)
if rank == 0: print(f"all_gather 33.5 MB: {us:.1f} us", flush=True)
torch.cuda.synchronize(); dist.barrier(); sys.stdout.flush(); os._exit(0)
