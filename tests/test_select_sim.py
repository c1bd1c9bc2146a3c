"""CPU: the 3-pass radix-select state machine (as implemented on the GPU) equals the sort-based
median of the oracle, for every mode, including simulated multi-rank histogram partitioning."""
import numpy as np
import pytest

from oracle import svgd_oracle as so
from oracle.select_sim import median_heuristic_select


@pytest.mark.parametrize("n,d", [(2, 1), (5, 3), (64, 2), (65, 7), (130, 4), (200, 33), (40, 60)])
@pytest.mark.parametrize("mode", so.MEDIAN_MODES)
@pytest.mark.parametrize("world", [1, 2, 3])
def test_select_equals_sort(n, d, mode, world):
    rng = np.random.default_rng(n + 7 * d)
    X = (rng.standard_normal((n, d)) * 2.0 - 0.3).astype(np.float32)
    if n > 4:
        X[3] = X[1]
    want = so.median_heuristic(X, mode)
    got = median_heuristic_select(X, mode, world)
    assert np.float32(got).tobytes() == np.float32(want).tobytes()


def test_select_nan():
    X = np.random.default_rng(0).standard_normal((70, 3)).astype(np.float32)
    X[66, 0] = np.nan
    assert np.isnan(median_heuristic_select(X, "full", 2))
