"""GPU: tcgen05 bring-up probes — TMA(SWIZZLE_128B) -> smem descriptors -> tcgen05.mma (SS and TS)
-> TMEM -> tcgen05.ld, checked against numpy on bf16-rounded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _probe_lib():
    """the probes live in a TEST-ONLY library (libnsvgd_probe.so, include/nsvgd_probe.h), not in the product library"""
    import ctypes as C
    import os
    from nsvgd import _lib
    _lib.load()                               # libnsvgd.so first: the probe library links against it
    lib = C.CDLL(os.path.join(os.path.dirname(_lib.LIB_PATH), "libnsvgd_probe.so"))
    vp, sz = C.c_void_p, C.c_size_t
    lib.nsvgd_tc_probe_gram.argtypes = [vp, vp, vp, vp, vp, sz]
    lib.nsvgd_tc_probe_pv.argtypes = [vp, vp, vp, vp, vp, sz]
    return lib


def bf16_round(a):
    return torch.as_tensor(a, dtype=torch.float32).to(torch.bfloat16).to(torch.float32).numpy()


def test_probe_gram():
    from nsvgd import _lib
    lib = _probe_lib()
    rng = np.random.default_rng(0)
    A = rng.standard_normal((128, 128)).astype(np.float32)
    B = rng.standard_normal((64, 128)).astype(np.float32)
    Ad, Bd = torch.tensor(A).cuda(), torch.tensor(B).cuda()
    G = torch.full((128, 64), float("nan"), device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    _lib.check(lib.nsvgd_tc_probe_gram(_lib.stream_ptr(), Ad.data_ptr(), Bd.data_ptr(), G.data_ptr(), ws.data_ptr(), ws.numel()))
    torch.cuda.synchronize()
    ref = bf16_round(A).astype(np.float64) @ bf16_round(B).astype(np.float64).T
    got = G.cpu().numpy()
    err = np.abs(got - ref).max()
    assert err < 1e-3 * np.abs(ref).max(), (err, got[:2, :4], ref[:2, :4])


def test_probe_pv():
    from nsvgd import _lib
    lib = _probe_lib()
    rng = np.random.default_rng(1)
    P = rng.random((128, 64)).astype(np.float32)
    V = rng.standard_normal((64, 256)).astype(np.float32)
    Pd, Vd = torch.tensor(P).cuda(), torch.tensor(V).cuda()
    O = torch.full((128, 256), float("nan"), device="cuda")
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    _lib.check(lib.nsvgd_tc_probe_pv(_lib.stream_ptr(), Pd.data_ptr(), Vd.data_ptr(), O.data_ptr(), ws.data_ptr(), ws.numel()))
    torch.cuda.synchronize()
    ref = bf16_round(P).astype(np.float64) @ bf16_round(V).astype(np.float64)
    got = O.cpu().numpy()
    err = np.abs(got - ref).max()
    assert err < 1e-3 * np.abs(ref).max(), (err, got[:2, :4], ref[:2, :4])
