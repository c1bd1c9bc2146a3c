"""Row-sharded SVGD step over torch.distributed (one process per GPU, NCCL over NVLink).

Rank g owns a contiguous block of particle rows (X, S, phi).  Per step:
  1. all-gather X and S                        (n/G * d * 4 bytes per rank each)
  2. median bandwidth: every rank histograms the tile rows  g, g+G, g+2G, ...  of the upper
     triangle of pairs, the uint64 histograms are all-reduced (sum) after each of the 3 radix
     passes, so every rank takes identical select decisions -> bit-identical h everywhere
  3. phi* for the local rows against all particles, with KSD/l2 partial sums
  4. all-reduce of the float64[4] statistics, then the local update x += lr * alpha * phi
The compute engine is injected (`engine=`) so that the host-side logic can be exercised on CPU
with gloo and an oracle-backed stand-in (tests/test_parallel_gloo.py); the default is the CUDA engine.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def row_partition(n: int, world: int, rank: int):
    """contiguous, balanced: the first n % world ranks get one extra row"""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class ShardedSVGD:
    def __init__(self, n: int, d: int, group=None, median_mode: str = "literal", engine=None, p2p="auto"):
        """p2p: "auto" -> the small all-reduces (select histograms, counts, KSD/l2 sums) run over NVLink peer memory
        inside their consumer kernels (csrc/p2p.cuh) when the CUDA engine is used and every rank could map its peers;
        "off" -> NCCL all-reduces; "on" -> peer memory or raise."""
        real_engine = engine is None
        if engine is None:
            from . import engine as _engine
            engine = _engine
        self.engine = engine
        self.n, self.d = n, d
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.median_mode = median_mode
        self.r0, self.r1 = row_partition(n, self.world, self.rank)
        self.max_rows = row_partition(n, self.world, 0)[1]
        self._buf = {}
        self._side = None
        self._hint = None          # warm start of the median bracket (engine.new_median_hint), created on first use
        self.comm = None
        import os
        if p2p == "auto" and os.environ.get("NSVGD_P2P", "1") == "0":
            p2p = "off"
        if p2p != "off" and real_engine and self.world > 1 and torch.cuda.is_available():
            ok, err = 1, None
            try:
                self.comm = engine.PeerComm(group)
            except Exception as e:          # cudaIpc not permitted in this container, no peer access, ...
                ok, err = 0, e
            oks = [None] * self.world
            dist.all_gather_object(oks, ok, group=group)                  # all ranks take the same path
            if min(oks) == 0:
                self.comm = None
                if p2p == "on":
                    raise RuntimeError(f"peer-memory communicator unavailable: {err}")

    def check_health(self):
        """Host-side check of the peer-memory watchdog (synchronises): a peer that did not arrive within the time limit
        has already turned h / alpha / the particles into NaN on the device (csrc/p2p.cuh); this raises on top of it."""
        if self.comm is not None and self.comm.timed_out():
            raise RuntimeError("a peer-memory all-reduce timed out: the results of the affected steps are NaN by construction")

    # ---- collectives -----------------------------------------------------------------------------
    def all_gather_rows(self, local: torch.Tensor, tag: str) -> torch.Tensor:
        """assemble the full (n,d) array from the per-rank row blocks (padded to equal size)"""
        assert local.shape == (self.r1 - self.r0, self.d)
        key = (tag, local.device, local.dtype)
        if key not in self._buf:
            self._buf[key] = (torch.empty((self.world * self.max_rows, self.d), dtype=local.dtype, device=local.device),
                              torch.zeros((self.max_rows, self.d), dtype=local.dtype, device=local.device))
        gathered, padded = self._buf[key]
        if self.n % self.world == 0:
            dist.all_gather_into_tensor(gathered, local.contiguous(), group=self.group)
            return gathered
        padded[: local.shape[0]].copy_(local)
        dist.all_gather_into_tensor(gathered, padded, group=self.group)
        parts = []
        for g in range(self.world):
            a, b = row_partition(self.n, self.world, g)
            parts.append(gathered[g * self.max_rows: g * self.max_rows + (b - a)])
        return torch.cat(parts, dim=0)

    def median_heuristic(self, X_all: torch.Tensor) -> torch.Tensor:
        sel = self.engine.MedianSelect(X_all, self.median_mode)
        sum_ = dict(op=dist.ReduceOp.SUM, group=self.group)
        if getattr(sel, "fast_supported", False) and self.comm is not None and hasattr(sel, "p2p_fast_run") \
                and (1 << 22) % self.world == 0 and ((1 << 22) // self.world) % 32 == 0:
            # one C call: sample slice, fused bracket select, filter, exact list, resolve, fused candidate select; every
            # exchange (4 histogram sums, the counts) happens inside its consumer kernel over NVLink peer memory
            # warm-started bracket: the previous step's order statistics (identical on every rank) replace the sampling and
            # bracket-select phases -- three of the serial exchange rounds of the step (NSVGD_MEDIAN_HINT=0 switches it off)
            if self._hint is None and os.environ.get("NSVGD_MEDIAN_HINT", "1") != "0":
                self._hint = self.engine.new_median_hint(X_all.device)
            sel.p2p_fast_run(self.comm, self._hint)
            return sel.finish()
        if getattr(sel, "fast_supported", False):
            # tensor-core filtered select: the bracket is identical on every rank (deterministic sample of
            # the gathered particles); counts and candidate histograms are all-reduced between the stages.
            nparts = self.world if (1 << 22) % self.world == 0 and ((1 << 22) // self.world) % 32 == 0 else 1
            if dist.get_backend(self.group) != "nccl":
                nparts = 1          # device-tensor all-gather needs NCCL; every rank then samples everything
            full, mine = sel.fast_sample(self.rank if nparts > 1 else 0, nparts)
            if nparts > 1:
                dist.all_gather_into_tensor(full, mine, group=self.group)
            red = sel.fast_stage1(self.rank, self.world)
            if self.comm is not None:
                sel.p2p_fast_resolve(self.comm)                      # all-reduce(red) fused into the resolve kernel
            else:
                dist.all_reduce(red, **sum_)
                sel.fast_resolve()
            # per radix pass ONE exchange of [candidate histogram | fallback histogram]: the conditional SIMT
            # fallback (kernels return at once unless the device-side flag is set) rides along with the
            # candidate select instead of costing three more collectives
            for p in range(3):
                sel.fast_cand_hist(p)
                sel.fallback_hist(p, self.rank, self.world)
                if self.comm is not None:
                    sel.p2p_fast_advance2(self.comm, p)              # all-reduce + both decisions in one kernel
                else:
                    dist.all_reduce(sel.hist_all, **sum_)
                    sel.fast_cand_advance(p)
                    sel.fallback_advance(p)
            sel.fast_commit()
            return sel.finish()
        if getattr(sel, "large_supported", False):
            # d > 128: every rank runs the tensor-core filtered select on the gathered array (an exact order statistic:
            # identical on every rank, no exchange); 9 ms at cfg5 against 191 / world ms for the sharded SIMT passes
            return self.engine.median_heuristic(X_all, self.median_mode)
        for p in range(3):
            hist = sel.hist_pass(p, tile_row_start=self.rank, tile_row_stride=self.world)
            if self.comm is not None:
                sel.p2p_advance(self.comm, p)
            else:
                dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)
                sel.advance(p)
        return sel.finish()

    # ---- one step --------------------------------------------------------------------------------
    def step(self, X_local, S_local, lr, scaled=True, lambda_reg=0.5, bandwidth=None, want_grads=False):
        """Updates X_local in place.  Returns (h, step_aux, grads_local | None)."""
        X_all = self.all_gather_rows(X_local, "X")
        if bandwidth or not X_local.is_cuda:
            S_all = self.all_gather_rows(S_local, "S")
            h = bandwidth if bandwidth else self.median_heuristic(X_all)  # truthiness as in models.py:657
        else:
            # the scores are not needed before phi*: their all-gather rides on a side stream under the bandwidth
            # selection (fork / join through events, so the pattern is also what a CUDA-graph capture records)
            cur = torch.cuda.current_stream(X_local.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=X_local.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                S_all = self.all_gather_rows(S_local, "S")
            h = self.median_heuristic(X_all)
            cur.wait_stream(self._side)
        # the queries as a row range OF the gathered array: the d > 128 path then reuses the particles' operand pieces and
        # knows which pairs are self pairs (csrc/phi_large.cu)
        Xq = X_all[self.r0:self.r1] if X_all.is_cuda and X_all.shape[0] == self.n else X_local
        phi, stats = self.engine.phistar(Xq, X_all, S_all, h, Sq=S_local, stats=True)
        if self.comm is not None:
            step_aux, grads = self.engine.apply_update_p2p(self.comm, X_local, phi, stats, lr, scaled, lambda_reg,
                                                           want_grads=want_grads)
        else:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
            step_aux, grads = self.engine.apply_update(X_local, phi, stats, lr, scaled, lambda_reg, want_grads=want_grads)
        return h, step_aux, grads

    # ---- the same step as ONE CUDA graph --------------------------------------------------------
    def capture(self, X_local, S_local, lr, scaled=True, lambda_reg=0.5, bandwidth=None):
        """Capture `step` (collectives included: NCCL is graph-capturable, every C-ABI entry is enqueue-only)
        into a CUDA graph.  A row-sharded step is ~45 kernels + ~12 small collectives issued from Python;
        at 8 GPUs the device work per step (~2 ms) is shorter than the time the host needs to issue it, so
        the graph is what removes the host from the critical path.  X_local / S_local are the static
        input/output buffers (X_local is updated in place by every replay).
        Returns a GraphedStep with .replay() and the static outputs .h, .step_aux."""
        self.step(X_local, S_local, lr, scaled, lambda_reg, bandwidth)          # warm-up: communicator, caches
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            h, step_aux, _ = self.step(X_local, S_local, lr, scaled, lambda_reg, bandwidth)
        return GraphedStep(graph, h, step_aux)


class GraphedStep:
    def __init__(self, graph, h, step_aux):
        self.graph, self.h, self.step_aux = graph, h, step_aux

    def replay(self):
        self.graph.replay()
        return self.h, self.step_aux
