"""CPU: libnsvgd.so loads, exports every symbol include/nsvgd.h declares, validates arguments,
and refuses to compute without a CUDA device (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT

from nsvgd import _lib, engine, kernels, models, stein, utils


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nsvgd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsvgd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nsvgd.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.nsvgd_version() >= 100


def test_workspace_query_and_argument_validation():
    lib = _lib.load()
    out = C.c_size_t(0)
    assert lib.nsvgd_workspace_bytes(_lib.OPS["phi"], 65536, 65536, 128, 0, C.byref(out)) == _lib.OK
    assert out.value > 3 * 65536 * 128 * 4
    assert lib.nsvgd_workspace_bytes(99, 10, 10, 2, 0, C.byref(out)) == _lib.ERR_INVALID
    assert b"bad op" in lib.nsvgd_last_error()
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.nsvgd_phi_f32(None, None, None, 1, 1, None, None, 1, 1, 1, None, None, 1, None, None, 0, None, 0) == _lib.ERR_INVALID
    assert lib.nsvgd_median_h_f32(None, None, 10, 2, 2, 0, None, None, 0) == _lib.ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_INVALID)
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_UNSUPPORTED)


def test_p2p_entry_points_validate_the_communicator():
    """peer-memory all-reduces (csrc/p2p.cuh): bad rank/world/blocks are rejected before any CUDA call"""
    lib = _lib.load()
    assert lib.nsvgd_p2p_block_bytes() == 1024 + 2 * 8192 * 8
    blocks = (C.c_void_p * 2)(None, None)
    assert lib.nsvgd_p2p_allreduce(None, 0, 2, blocks, None, 4, 0) == _lib.ERR_INVALID        # null peer blocks
    assert b"communicator" in lib.nsvgd_last_error()
    assert lib.nsvgd_p2p_allreduce(None, 3, 2, blocks, None, 4, 0) == _lib.ERR_INVALID        # rank >= world
    assert lib.nsvgd_p2p_allreduce(None, 0, 99, blocks, None, 4, 0) == _lib.ERR_INVALID       # world > 16
    assert lib.nsvgd_p2p_median_advance(None, 0, 2, blocks, None, None, 0) == _lib.ERR_INVALID
    assert lib.nsvgd_p2p_apply_update_f32(None, 0, 2, blocks, None, None, 1, 1, 1, 1, None, 0.1, 1, 0.5, None, None) == _lib.ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    x = torch.randn(8, 2)
    with pytest.raises(_lib.NsvgdError, match="no CPU fallback"):
        engine.median_heuristic(x)
    with pytest.raises(_lib.NsvgdError):
        engine.phistar(x, x, x, 1.0)
    with pytest.raises(_lib.NsvgdError):
        engine.mmd(x, x)


def test_reference_shape_guards():
    k = kernels.get_rbf_kernel(1.0)
    with pytest.raises(ValueError, match="Shapes of particles"):
        k(torch.zeros(2), torch.zeros(3))                      # kernels.py:13-15
    with pytest.raises(ValueError, match="more than one"):
        k(torch.zeros(2, 2), torch.zeros(2, 2))                # kernels.py:16-18
    with pytest.raises(ValueError, match="Bandwidth"):
        kernels.get_rbf_kernel(torch.ones(2, 2, 2))            # kernels.py:33-35
    assert float(k(torch.tensor([0.0, 0.0]), torch.tensor([1.0, 1.0]))) == pytest.approx(float(torch.exp(torch.tensor(-1.0))))
    with pytest.raises(ValueError, match="One of target_logp"):
        models.KernelGradient()                                # models.py:645-646
    with pytest.raises(NotImplementedError):
        stein.get_phistar(lambda x, y: x @ y, lambda x: -x @ x, torch.zeros(3, 2))


def test_function_combinators():
    f = lambda x: x * 2.0
    assert float(utils.negative(f)(torch.tensor(3.0))) == -6.0       # utils.py:515-519
    assert float(utils.mul(f, 0.5)(torch.tensor(3.0))) == 3.0        # utils.py:542-544
    bf = utils.BatchedFunction(lambda xs: xs + 1.0)
    g = utils.mul(utils.negative(bf), 2.0)
    xs = torch.zeros(4, 3)
    assert torch.equal(utils.vmap(g)(xs), torch.full((4, 3), -2.0))
    assert torch.equal(g(xs[0]), torch.full((3,), -2.0))
    m = torch.arange(9.0).reshape(3, 3)
    assert utils.remove_diagonal(m).tolist() == [[1, 2], [3, 5], [6, 7]]   # utils.py:653-658
