"""ctypes binding of libnsvgd.so (the C ABI declared in include/nsvgd.h).

There is NO fallback: if the shared library has not been built, or a tensor is not a
CUDA float32 tensor, every compute call raises.  torch is used here only for device
memory, the current stream and (in parallel.py) torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnsvgd.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_WORKSPACE = range(5)
MEDIAN_MODES = {"literal": 0, "offdiag": 1, "full": 2}
PATHS = {"auto": 0, "simt": 1, "tc": 2, "generic": 3, "large": 4}
OPS = {"median": 0, "phi": 1, "ksd": 2, "mmd": 3, "step": 4}
KSD_VARIANTS = {"rect": 0, "u": 1, "v": 2}
SELECT_BINS = 2048
SELECT_PASSES = 3
SELECT_STATE_BYTES = 256

# every symbol include/nsvgd.h declares (tests/test_abi.py checks the library exports all of them)
SYMBOLS = [
    "nsvgd_version", "nsvgd_last_error", "nsvgd_launch_count", "nsvgd_workspace_bytes",
    "nsvgd_profile_enable", "nsvgd_profile_read",
    "nsvgd_median_h_f32", "nsvgd_median_prepare", "nsvgd_median_hist_pass", "nsvgd_median_advance",
    "nsvgd_median_finish", "nsvgd_median_fast_supported", "nsvgd_median_fast_workspace_bytes",
    "nsvgd_median_fast_sample", "nsvgd_median_fast_stage1", "nsvgd_median_fast_resolve", "nsvgd_median_fast_cand_hist",
    "nsvgd_median_fast_cand_advance", "nsvgd_median_fast_commit", "nsvgd_median_fallback_hist",
    "nsvgd_median_fallback_advance", "nsvgd_phi_f32", "nsvgd_step_sgd_f32", "nsvgd_apply_update_f32",
    "nsvgd_ksd_f32", "nsvgd_mmd_f32", "nsvgd_ksum_f32", "nsvgd_adam_update_f32",
    "nsvgd_loop_small_supported", "nsvgd_svgd_loop_small_f32", "nsvgd_median_large_supported", "nsvgd_ksd_tc_workspace_bytes", "nsvgd_ksd_tc_f32",
    "nsvgd_median_h_hinted_f32", "nsvgd_p2p_median_fast_run_hinted",
    "nsvgd_scale_cols_f32", "nsvgd_matmul_f32", "nsvgd_ksd_ard_f32", "nsvgd_ksd_linear_f32", "nsvgd_ksd_dh_f32", "nsvgd_phi_vjp_workspace_bytes", "nsvgd_phi_vjp_f32",
    "nsvgd_score_funnel_f32", "nsvgd_score_gaussian_f32", "nsvgd_score_logreg_f32",
    "nsvgd_p2p_block_bytes", "nsvgd_p2p_alloc", "nsvgd_p2p_open", "nsvgd_p2p_close", "nsvgd_p2p_free",
    "nsvgd_p2p_timeout_flag", "nsvgd_p2p_allreduce", "nsvgd_p2p_median_advance", "nsvgd_p2p_median_fast_resolve",
    "nsvgd_p2p_median_fast_advance2", "nsvgd_p2p_median_fast_run", "nsvgd_p2p_apply_update_f32",
]


class NsvgdError(RuntimeError):
    pass


_lib = None


def load():
    """Load libnsvgd.so; raise loudly when it is missing (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NsvgdError(
            f"{LIB_PATH} not found: build it with `make -C projects-2020-neural-svgd_b200` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). nsvgd has no fallback path.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, f32, i32, sz = C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_size_t
    lib.nsvgd_version.restype = i32
    lib.nsvgd_last_error.restype = C.c_char_p
    lib.nsvgd_launch_count.restype = C.c_uint64
    lib.nsvgd_workspace_bytes.argtypes = [i32, i64, i64, i64, i32, C.POINTER(sz)]
    lib.nsvgd_profile_enable.argtypes = [i32]
    lib.nsvgd_profile_read.argtypes = [i32, C.POINTER(C.c_double), C.POINTER(i32)]
    lib.nsvgd_median_h_f32.argtypes = [vp, vp, i64, i64, i64, i32, vp, vp, sz]
    lib.nsvgd_median_prepare.argtypes = [vp, vp, i64, i64, i64, i32, vp, C.POINTER(i64), C.POINTER(f32), vp]
    lib.nsvgd_median_hist_pass.argtypes = [vp, vp, i64, i64, i64, f32, i64, i64, vp, i32, vp]
    lib.nsvgd_median_advance.argtypes = [vp, vp, vp, i32]
    lib.nsvgd_median_finish.argtypes = [vp, vp, i64, vp]
    lib.nsvgd_median_fast_supported.argtypes = [i64, i64]
    lib.nsvgd_median_fast_workspace_bytes.argtypes = [i64, i64, C.POINTER(sz)]
    lib.nsvgd_median_fast_sample.argtypes = [vp, vp, i64, i64, i64, f32, i64, i64, vp, sz, C.POINTER(sz), C.POINTER(i64)]
    lib.nsvgd_median_fast_stage1.argtypes = [vp, vp, i64, i64, i64, f32, i32, i64, i64, vp, vp, vp, sz, vp]
    lib.nsvgd_median_fast_resolve.argtypes = [vp, vp, vp, sz, i64, i64, vp]
    lib.nsvgd_median_fast_cand_hist.argtypes = [vp, vp, sz, i64, i64, i32, vp]
    lib.nsvgd_median_fast_cand_advance.argtypes = [vp, vp, sz, i64, i64, vp, i32]
    lib.nsvgd_median_fast_commit.argtypes = [vp, vp, vp, sz, i64, i64]
    lib.nsvgd_median_fallback_hist.argtypes = [vp, vp, i64, i64, i64, f32, i64, i64, vp, vp, sz, i32, vp]
    lib.nsvgd_median_fallback_advance.argtypes = [vp, vp, vp, sz, i64, i64, vp, i32]
    lib.nsvgd_phi_f32.argtypes = [vp, vp, vp, i64, i64, vp, vp, i64, i64, i64, vp, vp, i64, vp, vp, i32, vp, sz]
    lib.nsvgd_step_sgd_f32.argtypes = [vp, vp, vp, i64, i64, i64, vp, f32, i32, f32, vp, vp, i32, vp, sz]
    lib.nsvgd_apply_update_f32.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, f32, i32, f32, vp, vp]
    lib.nsvgd_ksd_f32.argtypes = [vp, vp, vp, i64, i64, vp, vp, i64, i64, i64, vp, i32, vp, vp, sz]
    lib.nsvgd_mmd_f32.argtypes = [vp, vp, i64, i64, vp, i64, i64, i64, f32, i32, vp, vp, sz]
    lib.nsvgd_ksum_f32.argtypes = [vp, vp, i64, i64, vp, i64, i64, i64, f32, i32, vp, vp, sz]
    lib.nsvgd_adam_update_f32.argtypes = [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, i64]
    lib.nsvgd_median_large_supported.argtypes = [i64, i64]
    lib.nsvgd_ksd_tc_workspace_bytes.argtypes = [i64, i64, C.POINTER(sz)]
    lib.nsvgd_ksd_tc_f32.argtypes = [vp, vp, vp, i64, i64, i64, vp, i32, vp, vp, sz]
    lib.nsvgd_loop_small_supported.argtypes = [i64, i64]
    lib.nsvgd_svgd_loop_small_f32.argtypes = [vp, vp, i64, i64, i32, vp, i32, f32, i32, f32, f32, i32, f32, i64, i64,
                                              vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, f32, vp, vp, vp]
    lib.nsvgd_scale_cols_f32.argtypes = [vp, vp, i64, i64, i64, vp, i32, vp, i64]
    lib.nsvgd_matmul_f32.argtypes = [vp, vp, i64, i32, vp, i64, vp, i64, i64, i64, i64, i32]
    lib.nsvgd_ksd_ard_f32.argtypes = [vp, vp, vp, i64, i64, vp, vp, i64, i64, i64, vp, f32, i32, vp, vp, sz]
    lib.nsvgd_ksd_linear_f32.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, vp, sz]
    lib.nsvgd_ksd_dh_f32.argtypes = [vp, vp, vp, i64, i64, vp, vp, i64, i64, i64, vp, i32, vp, vp, vp, sz]
    lib.nsvgd_phi_vjp_workspace_bytes.argtypes = [i64, i64, i64, C.POINTER(sz)]
    lib.nsvgd_phi_vjp_f32.argtypes = [vp, vp, vp, i64, i64, vp, vp, i64, i64, i64, vp, vp, i64, vp, i64, vp, i64, vp, vp, sz]
    lib.nsvgd_score_funnel_f32.argtypes = [vp, vp, i64, i64, i64, vp, i64, vp]
    lib.nsvgd_score_gaussian_f32.argtypes = [vp, vp, i64, i64, i64, vp, vp, i32, f32, vp, i64, vp]
    lib.nsvgd_score_logreg_f32.argtypes = [vp, vp, i64, i64, i64, vp, vp, i64, i64, f32, f32, f32, vp, i64, vp]
    pvp = C.POINTER(vp)
    lib.nsvgd_p2p_alloc.argtypes = [pvp, vp]
    lib.nsvgd_p2p_open.argtypes = [vp, pvp]
    lib.nsvgd_p2p_close.argtypes = [vp]
    lib.nsvgd_p2p_free.argtypes = [vp]
    lib.nsvgd_p2p_timeout_flag.argtypes = [vp, C.POINTER(i32)]
    lib.nsvgd_p2p_allreduce.argtypes = [vp, i32, i32, pvp, vp, i64, i32]
    lib.nsvgd_p2p_median_advance.argtypes = [vp, i32, i32, pvp, vp, vp, i32]
    lib.nsvgd_p2p_median_fast_resolve.argtypes = [vp, i32, i32, pvp, vp, vp, sz, i64, i64, vp]
    lib.nsvgd_p2p_median_fast_advance2.argtypes = [vp, i32, i32, pvp, vp, vp, sz, i64, i64, vp, i32]
    lib.nsvgd_p2p_median_fast_run.argtypes = [vp, vp, i64, i64, i64, f32, i32, i32, i32, pvp, vp, vp, vp, sz]
    lib.nsvgd_p2p_median_fast_run_hinted.argtypes = [vp, vp, i64, i64, i64, f32, i32, i32, i32, pvp, vp, vp, vp, vp, sz]
    lib.nsvgd_median_h_hinted_f32.argtypes = [vp, vp, i64, i64, i64, i32, vp, vp, vp, sz]
    lib.nsvgd_p2p_apply_update_f32.argtypes = [vp, i32, i32, pvp, vp, vp, i64, i64, i64, i64, vp, f32, i32, f32, vp, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name == "nsvgd_p2p_block_bytes":
            fn.restype = sz
        elif name not in ("nsvgd_version", "nsvgd_last_error", "nsvgd_launch_count"):
            fn.restype = i32
    _lib = lib
    return lib


def check(status: int):
    if status == OK:
        return
    msg = load().nsvgd_last_error().decode("utf-8", "replace")
    if status == ERR_INVALID:
        raise ValueError(f"nsvgd: {msg}")
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(f"nsvgd: {msg}")
    raise NsvgdError(f"nsvgd (status {status}): {msg}")


def profile_enable(on: bool):
    check(load().nsvgd_profile_enable(int(on)))


def profile_read(tag: int):
    """(total milliseconds, launches) recorded for `tag` since the last read"""
    ms, cnt = C.c_double(0), C.c_int(0)
    check(load().nsvgd_profile_read(tag, C.byref(ms), C.byref(cnt)))
    return float(ms.value), int(cnt.value)


def launch_count() -> int:
    return int(load().nsvgd_launch_count())


def require_cuda_f32(t: torch.Tensor, name: str, ndim: int | None = 2) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise NsvgdError(f"{name} lives on {t.device}: nsvgd computes on CUDA only and has no CPU fallback")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32 (the reference computes in fp32, stein.py:26), got {t.dtype}")
    if ndim is not None and t.ndim != ndim:
        raise ValueError(f"{name} must have {ndim} dimensions, got shape {tuple(t.shape)}")
    if t.ndim == 2 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def ld(t: torch.Tensor) -> int:
    """leading dimension (elements) of a row-major 2-D tensor"""
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def stream_ptr(device=None) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


_ws_cache: dict = {}


def workspace(nbytes: int, device) -> torch.Tensor:
    """Per-device scratch buffer, grown geometrically; the library never allocates."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        size = max(int(nbytes * 1.25), 1 << 20)
        buf = torch.empty(size, dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def workspace_bytes(op: str, n: int, m: int, d: int, path: str = "auto") -> int:
    out = C.c_size_t(0)
    check(load().nsvgd_workspace_bytes(OPS[op], n, m, d, PATHS[path], C.byref(out)))
    return int(out.value)


def ptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())
