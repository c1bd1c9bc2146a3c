"""GPU: the peer-memory all-reduces that row-sharded runs fuse into their consumer kernels (csrc/p2p.cuh), exercised by
TWO PROCESSES that share one GPU: cudaIpc maps each process's block into the other, the flags / epochs / slot
alternation are the real protocol.  (The multi-GPU form is tests/_sharded_worker.py under torchrun + NCCL.)"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_processes_one_gpu_fused_allreduces():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PYTHONUNBUFFERED="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_p2p_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            pytest.fail("p2p worker timed out")
        outs.append(out)
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out[-3000:]}"
        assert "bit-identical=True" in out and "allreduce ok=True" in out
