"""GPU (>= 2 devices): the row-sharded SVGD step under torchrun + NCCL reproduces the single-GPU step -- bandwidth
bit-identical (tensor-core filtered select and SIMT select, all three median modes), particles to 1e-6, eager and as one
replayed CUDA graph, with the small all-reduces fused over NVLink peer memory.  Skipped on a single-GPU box (there the
two-process-one-GPU test tests/test_p2p_gpu.py covers the peer-memory protocol)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("p2p", ["1", "0", "1+replicated-sample"])
def test_sharded_step_equals_single_gpu_step(p2p):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, NSVGD_P2P=p2p[0], PYTHONUNBUFFERED="1")
    if "replicated" in p2p:      # the bracket from a reduced sample set computed identically on every rank (default for >= 4 ranks)
        env["NSVGD_MEDIAN_REPLICATED_SAMPLE"] = "1"
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_sharded_worker.py")],
                         env=env, capture_output=True, text=True, timeout=600)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-4000:]
    assert text.count("h bit-identical=True") >= 8 and "bit-identical=False" not in text, text[-4000:]
    if p2p[0] == "1":
        assert "peer-memory=True" in text
