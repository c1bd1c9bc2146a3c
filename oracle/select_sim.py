"""NumPy simulation of the GPU radix-select state machine (csrc/pair_engine.cu: select_init /
pair_tile_kernel<MODE_HIST> / select_advance / select_finish).  TEST INFRASTRUCTURE ONLY.

Used (a) to check the 3-pass 11/11/10-bit select against a sort-based median on the CPU and
(b) as the oracle-backed compute stand-in of the gloo world_size-2 test of nsvgd/parallel.py,
where each rank histograms only its own tile rows and the histograms are all-reduced.
"""
from __future__ import annotations

import numpy as np

from . import svgd_oracle as so

BINS = 2048
TILE = 64
SHIFTS = (21, 10, 0)
NBINS = (2048, 2048, 1024)


class SelectSim:
    def __init__(self, x, mode):
        x = np.asarray(x, dtype=np.float32)
        self.n = x.shape[0]
        self.mode = mode
        z, w = so.literal_z(x) if mode == "literal" else (x, 1.0)
        self.z, self.w = z, w
        r_lo, r_hi = so.median_ranks(self.n, mode)
        self.is_zero = [r_lo < 0, r_hi < 0]
        self.rank = [max(r_lo, 0), max(r_hi, 0)]
        self.prefix = [0, 0]
        self.mask = 0
        self.nan = False
        # pair distances with their tile row, computed once (the GPU recomputes them every pass)
        iu, ju = np.triu_indices(self.n, 1)
        self._tile_row = iu // TILE
        self._bits = so.pair_sqdist_f32(z, w).view(np.uint32)
        self._isnan = np.isnan(self._bits.view(np.float32))

    def hist_pass(self, p, tile_row_start=0, tile_row_stride=1):
        mine = (self._tile_row >= tile_row_start) & ((self._tile_row - tile_row_start) % tile_row_stride == 0)
        if (self._isnan & mine).any():
            self.nan = True
        bits = self._bits[mine & ~self._isnan]
        hist = np.zeros(2 * BINS, dtype=np.int64)
        b = (bits >> np.uint32(SHIFTS[p])) & np.uint32(NBINS[p] - 1)
        for t in range(2):
            sel = (bits & np.uint32(self.mask)) == np.uint32(self.prefix[t])
            hist[t * BINS: t * BINS + NBINS[p]] = np.bincount(b[sel], minlength=NBINS[p])[: NBINS[p]]
        return hist

    def advance(self, p, hist):
        for t in range(2):
            hh = hist[t * BINS: t * BINS + NBINS[p]]
            cum = np.cumsum(hh)
            b = int(np.searchsorted(cum, self.rank[t], side="right"))
            if b < NBINS[p]:
                self.rank[t] -= int(cum[b] - hh[b])
                self.prefix[t] |= b << SHIFTS[p]
        self.mask |= (NBINS[p] - 1) << SHIFTS[p]

    def finish(self):
        if self.nan:
            return np.float32(np.nan)
        v = [np.float32(0) if self.is_zero[t] else np.array([self.prefix[t]], dtype=np.uint32).view(np.float32)[0]
             for t in range(2)]
        med = np.float32((v[0] + v[1]) * np.float32(0.5))
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.float32(np.sqrt(np.float32(np.float32(0.5) * med) / so.log_np1_f32(self.n)))


def median_heuristic_select(x, mode, world=1):
    """run the select with `world` simulated ranks (histograms summed between passes)"""
    sims = [SelectSim(x, mode) for _ in range(world)]
    for p in range(3):
        hist = sum(s.hist_pass(p, r, world) for r, s in enumerate(sims))
        for s in sims:
            s.nan = any(q.nan for q in sims)
            s.advance(p, hist)
    return sims[0].finish()
