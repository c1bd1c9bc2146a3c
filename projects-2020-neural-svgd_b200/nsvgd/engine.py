"""Array-facing operations of the SVGD particle-interaction engine (thin wrappers over the C ABI).

Everything is enqueue-only on the current CUDA stream; the bandwidth stays on the device
(a 1-element float32 tensor) so median -> phi* -> update never synchronises with the host.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import check, ld, load, ptr, require_cuda_f32, stream_ptr, workspace, workspace_bytes

DEFAULT_MEDIAN_MODE = os.environ.get("NSVGD_MEDIAN_MODE", "literal")


def as_h_dev(h, device) -> torch.Tensor:
    """bandwidth as a 1-element float32 device tensor (no sync when it already is one)."""
    if isinstance(h, torch.Tensor):
        if h.numel() != 1:
            raise ValueError(f"Bandwidth needs to be a scalar on the SVGD hot path. Instead it has shape {tuple(h.shape)}")
        return h.to(device=device, dtype=torch.float32).reshape(1)
    return torch.full((1,), float(h), dtype=torch.float32, device=device)


def new_median_hint(device) -> torch.Tensor:
    """32 bytes of persistent device state for `median_heuristic(..., hint=)` (struct MedianHint, csrc/select.cuh):
    [lo, hi, delta (float32) | valid, seen, hits, misses, pad (uint32)] viewed as int32."""
    return torch.zeros(8, dtype=torch.int32, device=device)


def median_hint_stats(hint: torch.Tensor) -> dict:
    v = hint.cpu()
    f = v.view(torch.float32)
    return {"lo": float(f[0]), "hi": float(f[1]), "delta": float(f[2]), "valid": int(v[3]), "seen": int(v[4]) & 1,
            "cooldown": (int(v[4]) >> 1) & 0x7FFF, "penalty": (int(v[4]) >> 16) & 0xFF, "hits": int(v[5]), "misses": int(v[6])}


def median_heuristic(x: torch.Tensor, mode: str | None = None, hint: torch.Tensor | None = None) -> torch.Tensor:
    """kernels.median_heuristic (kernels.py:160-172) -> 1-element float32 device tensor.

    mode: "literal" (what kernels.py:169 executes), "offdiag" (docstring intent), "full".
    hint: `new_median_hint(device)` kept by the caller across the steps of an SVGD loop: the tensor-core filtered select then
    warm-starts its bracket from the previous call (bit-identical result; sampling + bracket select skipped on a hit)."""
    mode = mode or DEFAULT_MEDIAN_MODE
    if mode not in _lib.MEDIAN_MODES:
        raise ValueError(f"unknown median mode {mode!r}")
    if x.ndim == 1:
        x = x[:, None]
    x = require_cuda_f32(x, "x")
    n, d = x.shape
    h = torch.empty(1, dtype=torch.float32, device=x.device)
    ws = workspace(workspace_bytes("median", n, n, d), x.device)
    if hint is not None and os.environ.get("NSVGD_MEDIAN_HINT", "1") != "0":
        check(load().nsvgd_median_h_hinted_f32(stream_ptr(x.device), ptr(x), n, d, ld(x), _lib.MEDIAN_MODES[mode],
                                               ptr(h), ptr(hint), ptr(ws), ws.numel()))
    else:
        check(load().nsvgd_median_h_f32(stream_ptr(x.device), ptr(x), n, d, ld(x), _lib.MEDIAN_MODES[mode],
                                        ptr(h), ptr(ws), ws.numel()))
    return h


def phistar(Q, X, S, h, Sq=None, aux=False, stats=False, path="auto"):
    """phi*(q_i) for all rows of Q (stein.py:180-222).  Returns phi, or (phi, aux (m,2,d)), and
    additionally a float64[4] device tensor {sum_i(s_i.phi_i + div phi_i), sum_i|phi_i|^2, m, 0}
    when stats=True (requires Sq = scores at the queries)."""
    X = require_cuda_f32(X, "X"); S = require_cuda_f32(S, "S"); Q = require_cuda_f32(Q, "Q")
    if X.shape != S.shape:
        raise ValueError(f"particles and scores must have the same shape, got {tuple(X.shape)} and {tuple(S.shape)}")
    if Q.shape[1] != X.shape[1]:
        raise ValueError(f"Shapes of particles x and y need to match. Received d={Q.shape[1]} and d={X.shape[1]}")
    if ld(X) != ld(S):
        S = S.contiguous(); X = X.contiguous()
    m, d = Q.shape
    n = X.shape[0]
    if Sq is not None:
        Sq = require_cuda_f32(Sq, "Sq")
        if Sq.shape != Q.shape:
            raise ValueError("Sq must have the shape of Q")
        if ld(Sq) != ld(Q):
            Sq = Sq.contiguous(); Q = Q.contiguous()
    if stats and Sq is None:
        raise ValueError("stats=True needs the scores at the query points (Sq)")
    h_dev = as_h_dev(h, X.device)
    phi = torch.empty((m, d), dtype=torch.float32, device=X.device)
    aux_t = torch.empty((m, 2, d), dtype=torch.float32, device=X.device) if aux else None
    stats_t = torch.empty(4, dtype=torch.float64, device=X.device) if stats else None
    ws = workspace(workspace_bytes("phi", n, m, d, path), X.device)
    check(load().nsvgd_phi_f32(stream_ptr(X.device), ptr(Q), ptr(Sq), m, ld(Q), ptr(X), ptr(S), n, ld(X), d,
                               ptr(h_dev), ptr(phi), d, ptr(aux_t), ptr(stats_t), _lib.PATHS[path], ptr(ws), ws.numel()))
    out = (phi,)
    if aux:
        out += (aux_t,)
    if stats:
        out += (stats_t,)
    return out[0] if len(out) == 1 else out


def svgd_step_sgd(X, S, h, lr, scaled=False, lambda_reg=0.5, want_grads=False, path="auto"):
    """One fused SVGD step in place on X (models.py:143-167 with optax.sgd; 654-689):
    grads = -phi or -alpha phi, X <- X - lr grads.  Returns (step_aux float64[8] device tensor
    {ksd, l2, alpha, sum grads^2, sum updates^2, ...}, grads or None)."""
    X = require_cuda_f32(X, "X"); S = require_cuda_f32(S, "S")
    if not X.is_contiguous():
        raise ValueError("X must be contiguous for the in-place step")
    S = S.contiguous()
    n, d = X.shape
    h_dev = as_h_dev(h, X.device)
    grads = torch.empty_like(X) if want_grads else None
    step_aux = torch.empty(8, dtype=torch.float64, device=X.device)
    ws = workspace(workspace_bytes("step", n, n, d, path), X.device)
    check(load().nsvgd_step_sgd_f32(stream_ptr(X.device), ptr(X), ptr(S), n, d, d, ptr(h_dev), float(lr), int(bool(scaled)),
                                    float(lambda_reg), ptr(grads), ptr(step_aux), _lib.PATHS[path], ptr(ws), ws.numel()))
    return step_aux, grads


def apply_update(Xrows, phi, stats, lr, scaled, lambda_reg, want_grads=False):
    """second phase of a row-sharded step: stats are the (all-reduced) global sums."""
    m, d = Xrows.shape
    grads = torch.empty_like(phi) if want_grads else None
    step_aux = torch.empty(8, dtype=torch.float64, device=Xrows.device)
    check(load().nsvgd_apply_update_f32(stream_ptr(Xrows.device), ptr(Xrows), ptr(phi), m, d, ld(Xrows), ld(phi), ptr(stats),
                                        float(lr), int(bool(scaled)), float(lambda_reg), ptr(grads), ptr(step_aux)))
    return step_aux, grads


def apply_update_p2p(comm, Xrows, phi, stats, lr, scaled, lambda_reg, want_grads=False):
    """apply_update with the all-reduce of `stats` (float64[4], summed in place in rank order) over peer memory
    fused into the kernel that derives the step scaling."""
    m, d = Xrows.shape
    grads = torch.empty_like(phi) if want_grads else None
    step_aux = torch.empty(8, dtype=torch.float64, device=Xrows.device)
    check(load().nsvgd_p2p_apply_update_f32(stream_ptr(Xrows.device), comm.rank, comm.world, comm.blocks, ptr(Xrows), ptr(phi),
                                            m, d, ld(Xrows), ld(phi), ptr(stats), float(lr), int(bool(scaled)), float(lambda_reg),
                                            ptr(grads), ptr(step_aux)))
    return step_aux, grads


class PeerComm:
    """Peer-mapped blocks for the fused small all-reduces (csrc/p2p.cuh): one cudaMalloc'ed block per rank, shared
    through cudaIpc handles that are exchanged with one all-gather of the process group."""

    def __init__(self, group=None):
        """The handshake runs in phases that cannot diverge: EVERY rank runs the same sequence of collectives (gather the
        handles -- None on a local failure --, try the opens, gather an ok flag, barrier); failures are caught locally and
        reported through the gathered status, so either all ranks get a communicator or all ranks raise."""
        import torch.distributed as dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        lib = load()
        self.own, self.peers, self.blocks, err = None, [], None, None
        handle = (C.c_ubyte * 64)()
        try:
            own = C.c_void_p()
            check(lib.nsvgd_p2p_alloc(C.byref(own), handle))            # zero-filled and synchronised
            self.own = own.value
        except Exception as e:          # cudaIpc not permitted, out of memory, ...
            err = e
        # handles travel as python objects: works on every backend (nccl, gloo) and also orders every rank's zero-fill
        # before any peer can write a flag
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle) if err is None else None, group=group)
        if err is None and any(h is None for h in handles):
            err = RuntimeError("a peer could not allocate its block")
        if err is None:
            try:
                ptrs = (C.c_void_p * self.world)()
                for r in range(self.world):
                    if r == self.rank:
                        ptrs[r] = self.own
                        continue
                    pp = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                    check(lib.nsvgd_p2p_open(buf, C.byref(pp)))
                    ptrs[r] = pp.value
                    self.peers.append(pp.value)
                self.blocks = ptrs
            except Exception as e:      # no peer access between these devices, ...
                err = e
        oks = [None] * self.world
        dist.all_gather_object(oks, 1 if err is None else 0, group=group)
        dist.barrier(group=group)
        if min(oks) == 0:
            self.close()
            raise RuntimeError(f"peer-memory communicator unavailable on rank(s) {[r for r, o in enumerate(oks) if not o]}"
                               + (f": {err}" if err is not None else ""))

    def close(self):
        """unmap the peers' blocks and free the own one (idempotent; also run by the destructor)"""
        lib = load()
        for p in getattr(self, "peers", []):
            lib.nsvgd_p2p_close(p)
        self.peers = []
        if getattr(self, "own", None):
            lib.nsvgd_p2p_free(self.own)
        self.own, self.blocks = None, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def timed_out(self) -> bool:
        flag = C.c_int(0)
        check(load().nsvgd_p2p_timeout_flag(self.own, C.byref(flag)))
        return bool(flag.value)

    def allreduce(self, t: torch.Tensor) -> torch.Tensor:
        """stand-alone in-place sum (int64 or float64 tensors of <= 8192 elements)"""
        assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.int64, torch.float64) and t.numel() <= 8192
        check(load().nsvgd_p2p_allreduce(stream_ptr(t.device), self.rank, self.world, self.blocks, ptr(t), t.numel(),
                                         int(t.dtype == torch.float64)))
        return t


KSD_TC_MIN_N = 4096      # from here on the phi*-statistics route beats the fp32 SIMT pair tiles


def ksd(X, Sx, h, variant="v", Y=None, Sy=None) -> torch.Tensor:
    """KSD^2 closed form (stein.py:273-349) -> 0-dim float64 device tensor."""
    X = require_cuda_f32(X, "X").contiguous(); Sx = require_cuda_f32(Sx, "Sx").contiguous()
    n, d = X.shape
    m = n
    if variant == "rect":
        Y = require_cuda_f32(Y, "Y").contiguous(); Sy = require_cuda_f32(Sy, "Sy").contiguous()
        m = Y.shape[0]
    h_dev = as_h_dev(h, X.device)
    out = torch.empty(1, dtype=torch.float64, device=X.device)
    if variant in ("u", "v") and n >= KSD_TC_MIN_N and os.environ.get("NSVGD_KSD_TC", "1") != "0":
        # at scale the U / V statistics come out of the fused phi* kernels (tensor cores for d > 32): one phi* pass
        need = C.c_size_t(0)
        check(load().nsvgd_ksd_tc_workspace_bytes(n, d, C.byref(need)))
        ws = workspace(int(need.value), X.device)
        check(load().nsvgd_ksd_tc_f32(stream_ptr(X.device), ptr(X), ptr(Sx), n, d, d, ptr(h_dev), _lib.KSD_VARIANTS[variant],
                                      ptr(out), ptr(ws), ws.numel()))
        return out[0]
    ws = workspace(workspace_bytes("ksd", n, m, d), X.device)
    check(load().nsvgd_ksd_f32(stream_ptr(X.device), ptr(X), ptr(Sx), n, d, ptr(Y), ptr(Sy), m, d, d, ptr(h_dev),
                               _lib.KSD_VARIANTS[variant], ptr(out), ptr(ws), ws.numel()))
    return out[0]


def scale_cols(X, v, invert=False):
    """X * v or X / v per column on the device (the linear map of a vector bandwidth; csrc/vjp.cu)."""
    X = require_cuda_f32(X, "X")
    n, d = X.shape
    v = v.to(device=X.device, dtype=torch.float32).contiguous()
    out = torch.empty((n, d), dtype=torch.float32, device=X.device)
    check(load().nsvgd_scale_cols_f32(stream_ptr(X.device), ptr(X), n, d, ld(X), ptr(v), int(bool(invert)), ptr(out), d))
    return out


def matmul(A, B):
    """A @ B in fp32 through the library's SIMT GEMM (the d x d linear maps of a full-matrix Gaussian kernel)."""
    A = require_cuda_f32(A, "A"); B = require_cuda_f32(B.to(A.device), "B").contiguous()
    M, K = A.shape
    N = B.shape[1]
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    check(load().nsvgd_matmul_f32(stream_ptr(A.device), ptr(A), ld(A), 0, ptr(B), N, ptr(out), N, M, N, K, 0))
    return out


def ksd_ard(X, Sx, h_vec, variant="v", Y=None, Sy=None) -> torch.Tensor:
    """KSD^2 under a per-coordinate bandwidth vector h_a (closed form, csrc/pair_engine.cu) -> 0-dim float64 tensor."""
    X = require_cuda_f32(X, "X").contiguous(); Sx = require_cuda_f32(Sx, "Sx").contiguous()
    n, d = X.shape
    m = n
    if variant == "rect":
        Y = require_cuda_f32(Y, "Y").contiguous(); Sy = require_cuda_f32(Sy, "Sy").contiguous()
        m = Y.shape[0]
    w = (1.0 / (h_vec.to(device=X.device, dtype=torch.float32) ** 2)).contiguous()
    if w.shape != (d,):
        raise ValueError(f"Bandwidth has shape {tuple(h_vec.shape)}.")
    out = torch.empty(1, dtype=torch.float64, device=X.device)
    ws = workspace(workspace_bytes("ksd", n, m, d), X.device)
    check(load().nsvgd_ksd_ard_f32(stream_ptr(X.device), ptr(X), ptr(Sx), n, d, ptr(Y), ptr(Sy), m, d, d, ptr(w), float(w.sum().item()),
                                   _lib.KSD_VARIANTS[variant], ptr(out), ptr(ws), ws.numel()))
    return out[0]


def ksd_linear(X, S, h):
    """linear-time KSD estimator (stein.py:353-380) -> float64[2] device tensor {mean, std(ddof=1) / (n/2)}."""
    X = require_cuda_f32(X, "samples").contiguous(); S = require_cuda_f32(S, "scores").contiguous()
    n, d = X.shape
    h_dev = as_h_dev(h, X.device)
    out = torch.empty(2, dtype=torch.float64, device=X.device)
    ws = workspace(max(4 * (n // 2), 256), X.device)
    check(load().nsvgd_ksd_linear_f32(stream_ptr(X.device), ptr(X), ptr(S), n, d, d, ptr(h_dev), ptr(out), ptr(ws), ws.numel()))
    return out


def mmd(X, Y, h=1.0, funnel=False) -> torch.Tensor:
    """MMD^2 (metrics.py:88-105) -> 0-dim float64 device tensor."""
    X = require_cuda_f32(X, "xs").contiguous(); Y = require_cuda_f32(Y, "ys").contiguous()
    if X.shape[1] != Y.shape[1]:
        raise ValueError(f"Shapes of particles x and y need to match. Received d={X.shape[1]} and d={Y.shape[1]}")
    n, d = X.shape
    m = Y.shape[0]
    out = torch.empty(1, dtype=torch.float64, device=X.device)
    ws = workspace(workspace_bytes("mmd", n, m, d), X.device)
    check(load().nsvgd_mmd_f32(stream_ptr(X.device), ptr(X), n, d, ptr(Y), m, d, d, float(h), int(bool(funnel)),
                               ptr(out), ptr(ws), ws.numel()))
    return out[0]


def adam_update(X, grads, m1, m2, lr, t, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam(lr) applied to the particles in place (utils.py:711-715)."""
    check(load().nsvgd_adam_update_f32(stream_ptr(X.device), ptr(X), ptr(grads), ptr(m1), ptr(m2), X.numel(),
                                       float(lr), float(b1), float(b2), float(eps), int(t)))


# ---- device-side scores of the canonical targets (csrc/scores.cu) -----------------------------------
def score_funnel(X, want_logp=False):
    """grad log p (and log p) of Neal's funnel at every row of X (distributions.py:275-319)."""
    X = require_cuda_f32(X, "X")
    n, d = X.shape
    S = torch.empty((n, d), dtype=torch.float32, device=X.device)
    lp = torch.empty(n, dtype=torch.float32, device=X.device) if want_logp else None
    check(load().nsvgd_score_funnel_f32(stream_ptr(X.device), ptr(X), n, d, ld(X), ptr(S), d, ptr(lp)))
    return (S, lp) if want_logp else S


def score_gaussian(X, mean, prec, log_norm=0.0, want_logp=False):
    """grad log p (and log p) of N(mean, prec^-1); prec is a (d,d) precision matrix or its (d,) diagonal."""
    X = require_cuda_f32(X, "X")
    n, d = X.shape
    mean = require_cuda_f32(mean.to(X.device), "mean", ndim=1).contiguous()
    prec = prec.to(device=X.device, dtype=torch.float32).contiguous()
    if prec.shape not in ((d,), (d, d)) or mean.shape != (d,):
        raise ValueError(f"mean must be ({d},) and prec ({d},) or ({d},{d}); got {tuple(mean.shape)}, {tuple(prec.shape)}")
    S = torch.empty((n, d), dtype=torch.float32, device=X.device)
    lp = torch.empty(n, dtype=torch.float32, device=X.device) if want_logp else None
    check(load().nsvgd_score_gaussian_f32(stream_ptr(X.device), ptr(X), n, d, ld(X), ptr(mean), ptr(prec), int(prec.ndim == 1),
                                          float(log_norm), ptr(S), d, ptr(lp)))
    return (S, lp) if want_logp else S


def score_logreg(P, Xb, yb, num_datapoints, a0=1.0, b0=0.01, want_logp=False):
    """grad log p (and log p) of the Bayesian logistic-regression posterior at the particles P = [w | log alpha]
    for the minibatch (Xb, yb)  (experiments/bayesian_logistic_regression.py:73-114, 169-190)."""
    P = require_cuda_f32(P, "params")
    Xb = require_cuda_f32(Xb, "x").contiguous()
    yb = require_cuda_f32(yb.to(torch.float32), "y", ndim=1).contiguous()
    n, d = P.shape
    B, F = Xb.shape
    if yb.shape[0] != B:
        raise ValueError("x and y must have the same number of data points")
    S = torch.empty((n, d), dtype=torch.float32, device=P.device)
    lp = torch.empty(n, dtype=torch.float32, device=P.device) if want_logp else None
    check(load().nsvgd_score_logreg_f32(stream_ptr(P.device), ptr(P), n, d, ld(P), ptr(Xb), ptr(yb), B, F, float(num_datapoints),
                                        float(a0), float(b0), ptr(S), d, ptr(lp)))
    return (S, lp) if want_logp else S


# ---- select building blocks for the row-sharded median (parallel.py) -------------------------------
class MedianSelect:
    """Stateful wrapper over nsvgd_median_{prepare,hist_pass,advance,finish}."""

    def __init__(self, x: torch.Tensor, mode: str):
        x = require_cuda_f32(x, "x")
        self.x, self.mode = x, mode
        n, d = x.shape
        self.n = n
        self.state = torch.zeros(_lib.SELECT_STATE_BYTES, dtype=torch.uint8, device=x.device)
        # [select histogram (2 targets) | fallback-select histogram (2 targets)]: contiguous so that a row-sharded
        # run reduces both with ONE all-reduce per radix pass
        self.hist_all = torch.zeros(4 * _lib.SELECT_BINS, dtype=torch.int64, device=x.device)
        self.hist = self.hist_all[: 2 * _lib.SELECT_BINS]
        self.hist_fb = self.hist_all[2 * _lib.SELECT_BINS:]
        dz_max = d if n > d else n - 1
        self.Z = torch.empty((n, dz_max), dtype=torch.float32, device=x.device) if mode == "literal" else None
        dz, w = C.c_int64(0), C.c_float(0)
        check(load().nsvgd_median_prepare(stream_ptr(x.device), ptr(x), n, d, ld(x), _lib.MEDIAN_MODES[mode], ptr(self.Z),
                                          C.byref(dz), C.byref(w), ptr(self.state)))
        self.dz, self.w_last = int(dz.value), float(w.value)
        self.P = self.Z if mode == "literal" else x
        self.fast_supported = bool(load().nsvgd_median_fast_supported(n, self.dz)) and os.environ.get("NSVGD_MEDIAN_FAST", "1") != "0"
        # d > 128: the single-GPU filtered select (csrc/median_large.cu) beats the row-sharded 3-pass SIMT select up to 16+ ranks
        # (the single-GPU driver pads its own literal-mode Z to float4 rows; a caller's array must already be aligned)
        self.large_supported = bool(load().nsvgd_median_large_supported(n, self.dz)) and (
            mode == "literal" or (ld(x) % 4 == 0 and x.data_ptr() % 16 == 0 and self.dz % 4 == 0))
        self._fast_ws = None
        self.reduce = torch.zeros(4, dtype=torch.int64, device=x.device)

    # ---- tensor-core filtered select, staged for all-reduces in between (include/nsvgd.h) ----------
    def _fws(self):
        if self._fast_ws is None:
            need = C.c_size_t(0)
            check(load().nsvgd_median_fast_workspace_bytes(self.n, self.dz, C.byref(need)))
            self._fast_ws = torch.empty(int(need.value), dtype=torch.uint8, device=self.x.device)
        return self._fast_ws

    def fast_sample(self, part=0, nparts=1):
        """stage 0: prep + this rank's slice of the sampled distances.  Returns (full sample buffer, local
        slice) as float32 views into the workspace; all-gather the slices into the full buffer when sharded."""
        ws = self._fws()
        off, cnt = C.c_size_t(0), C.c_int64(0)
        check(load().nsvgd_median_fast_sample(stream_ptr(self.x.device), ptr(self.P), self.n, self.dz, ld(self.P), self.w_last,
                                              part, nparts, ptr(ws), ws.numel(), C.byref(off), C.byref(cnt)))
        full = ws[off.value: off.value + 4 * cnt.value].view(torch.float32)
        per = cnt.value // nparts
        return full, full[part * per: (part + 1) * per]

    def fast_stage1(self, block_row_start=0, block_row_stride=1) -> torch.Tensor:
        ws = self._fws()
        check(load().nsvgd_median_fast_stage1(stream_ptr(self.x.device), ptr(self.P), self.n, self.dz, ld(self.P), self.w_last,
                                              _lib.MEDIAN_MODES[self.mode], block_row_start, block_row_stride, ptr(self.state),
                                              ptr(self.hist), ptr(ws), ws.numel(), ptr(self.reduce)))
        return self.reduce

    def fast_resolve(self):
        ws = self._fws()
        check(load().nsvgd_median_fast_resolve(stream_ptr(self.x.device), ptr(self.state), ptr(ws), ws.numel(), self.n, self.dz,
                                               ptr(self.reduce)))

    def fast_cand_hist(self, p: int) -> torch.Tensor:
        ws = self._fws()
        check(load().nsvgd_median_fast_cand_hist(stream_ptr(self.x.device), ptr(ws), ws.numel(), self.n, self.dz, p, ptr(self.hist)))
        return self.hist

    def fast_cand_advance(self, p: int):
        ws = self._fws()
        check(load().nsvgd_median_fast_cand_advance(stream_ptr(self.x.device), ptr(ws), ws.numel(), self.n, self.dz, ptr(self.hist), p))

    def fast_commit(self):
        ws = self._fws()
        check(load().nsvgd_median_fast_commit(stream_ptr(self.x.device), ptr(self.state), ptr(ws), ws.numel(), self.n, self.dz))

    def fallback_hist(self, p: int, tile_row_start=0, tile_row_stride=1) -> torch.Tensor:
        ws = self._fws()
        check(load().nsvgd_median_fallback_hist(stream_ptr(self.x.device), ptr(self.P), self.n, self.dz, ld(self.P), self.w_last,
                                                tile_row_start, tile_row_stride, ptr(self.state), ptr(ws), ws.numel(), p, ptr(self.hist_fb)))
        return self.hist_fb

    def fallback_advance(self, p: int):
        ws = self._fws()
        check(load().nsvgd_median_fallback_advance(stream_ptr(self.x.device), ptr(self.state), ptr(ws), ws.numel(), self.n, self.dz,
                                                   ptr(self.hist_fb), p))

    # ---- peer-memory variants: the all-reduce happens inside the consumer kernel (include/nsvgd.h, csrc/p2p.cuh) ----
    def p2p_fast_resolve(self, comm):
        ws = self._fws()
        check(load().nsvgd_p2p_median_fast_resolve(stream_ptr(self.x.device), comm.rank, comm.world, comm.blocks, ptr(self.state),
                                                   ptr(ws), ws.numel(), self.n, self.dz, ptr(self.reduce)))

    def p2p_fast_advance2(self, comm, p: int):
        ws = self._fws()
        check(load().nsvgd_p2p_median_fast_advance2(stream_ptr(self.x.device), comm.rank, comm.world, comm.blocks, ptr(self.state),
                                                    ptr(ws), ws.numel(), self.n, self.dz, ptr(self.hist_all), p))

    def p2p_fast_run(self, comm, hint=None):
        """the whole filtered select of this rank in one C call (fused selects, every exchange inside its consumer kernel);
        hint: see engine.median_heuristic (identical content on every rank by construction)"""
        ws = self._fws()
        check(load().nsvgd_p2p_median_fast_run_hinted(stream_ptr(self.x.device), ptr(self.P), self.n, self.dz, ld(self.P),
                                                      self.w_last, _lib.MEDIAN_MODES[self.mode], comm.rank, comm.world, comm.blocks,
                                                      ptr(self.state), ptr(self.hist), ptr(hint), ptr(ws), ws.numel()))

    def p2p_advance(self, comm, p: int):
        check(load().nsvgd_p2p_median_advance(stream_ptr(self.x.device), comm.rank, comm.world, comm.blocks, ptr(self.state),
                                              ptr(self.hist), p))

    def hist_pass(self, p: int, tile_row_start=0, tile_row_stride=1) -> torch.Tensor:
        check(load().nsvgd_median_hist_pass(stream_ptr(self.x.device), ptr(self.P), self.n, self.dz, ld(self.P), self.w_last,
                                            tile_row_start, tile_row_stride, ptr(self.state), p, ptr(self.hist)))
        return self.hist

    def advance(self, p: int):
        check(load().nsvgd_median_advance(stream_ptr(self.x.device), ptr(self.state), ptr(self.hist), p))

    def finish(self) -> torch.Tensor:
        h = torch.empty(1, dtype=torch.float32, device=self.x.device)
        check(load().nsvgd_median_finish(stream_ptr(self.x.device), ptr(self.state), self.n, ptr(h)))
        return h
