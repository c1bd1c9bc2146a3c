"""CPU, world_size 2, gloo: the host-side logic of the row-sharded step (nsvgd/parallel.py) —
row partition, padded all-gather, per-rank tile-row histograms + all-reduce between radix passes,
statistics all-reduce, local update — with an oracle-backed compute stand-in in place of the CUDA
engine (test infrastructure only; the product engine is CUDA and has no CPU path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import svgd_oracle as so
from oracle.select_sim import SelectSim


class OracleEngine:
    """same call surface as nsvgd.engine, NumPy oracle underneath"""

    class MedianSelect:
        def __init__(self, x, mode):
            self.sim = SelectSim(x.numpy(), mode)
            self.hist = torch.zeros(4096, dtype=torch.int64)

        def hist_pass(self, p, tile_row_start=0, tile_row_stride=1):
            self.hist.copy_(torch.from_numpy(self.sim.hist_pass(p, tile_row_start, tile_row_stride)))
            return self.hist

        def advance(self, p):
            self.sim.advance(p, self.hist.numpy())

        def finish(self):
            return torch.tensor([self.sim.finish()], dtype=torch.float32)

    @staticmethod
    def phistar(Q, X, S, h, Sq=None, stats=False):
        h = float(h)
        q, x, s = Q.numpy(), X.numpy(), S.numpy()
        phi = so.phistar(q, x, s, h)
        div = so.div_phistar(q, x, s, h)
        st = torch.tensor([float(np.sum(np.einsum("id,id->i", Sq.numpy().astype(np.float64), phi) + div)),
                           float(np.sum(phi * phi)), float(q.shape[0]), 0.0], dtype=torch.float64)
        return torch.from_numpy(phi.astype(np.float32)), st

    @staticmethod
    def apply_update(Xrows, phi, stats, lr, scaled, lambda_reg, want_grads=False):
        ksd, l2 = stats[0] / stats[2], stats[1] / stats[2]
        alpha = float(ksd / (2 * lambda_reg * l2)) if scaled else 1.0
        grads = -alpha * phi
        Xrows -= lr * grads
        return torch.tensor([float(ksd), float(l2), alpha, 0, 0, 0, 0, 0], dtype=torch.float64), (grads if want_grads else None)


def _worker(rank, world, port, n, d, mode, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "projects-2020-neural-svgd_b200"))
    from nsvgd.parallel import ShardedSVGD, row_partition
    X, S = so.gaussian_diag_workload(n, d, seed=11)
    S = (S * 0.05).astype(np.float32)
    r0, r1 = row_partition(n, world, rank)
    Xl, Sl = torch.tensor(X[r0:r1]), torch.tensor(S[r0:r1])
    sh = ShardedSVGD(n, d, median_mode=mode, engine=OracleEngine)
    h, aux, _ = sh.step(Xl, Sl, lr=0.05, scaled=True, lambda_reg=0.5)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), X=Xl.numpy(), h=h.numpy(), aux=aux.numpy(), r=np.array([r0, r1]))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,d,mode", [(150, 5, "literal"), (131, 3, "offdiag")])
def test_sharded_step_world2_equals_single(tmp_path, n, d, mode):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, d, mode, str(tmp_path)), nprocs=world, join=True)
    X, S = so.gaussian_diag_workload(n, d, seed=11)
    S = (S * 0.05).astype(np.float32)
    h_ref = so.median_heuristic(X, mode)
    newX, _, _, extras = so.svgd_step_sgd(X, S, 0.05, bandwidth=float(h_ref), scaled=True, lambda_reg=0.5)
    got = np.zeros_like(newX)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.float32(z["h"][0]).tobytes() == np.float32(h_ref).tobytes()      # identical h on every rank
        got[z["r"][0]: z["r"][1]] = z["X"]
        assert z["aux"][2] == pytest.approx(extras["alpha"], rel=1e-6)
    np.testing.assert_allclose(got, newX, rtol=2e-6, atol=2e-6)


def test_row_partition_covers_everything():
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "projects-2020-neural-svgd_b200"))
    from nsvgd.parallel import row_partition
    for n in (1, 7, 64, 65536, 1000003):
        for w in (1, 2, 3, 8):
            edges = [row_partition(n, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
