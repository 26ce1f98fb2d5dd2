"""The production random streams (csrc/rtb_shading.cuh): every path sample owns the stream
pcg_seed(sample index, seed); its float draws are the raw top 24 bits of a 64-bit LCG whose increment all
streams share (a cut in instructions on an issue-bound kernel).  The image gates would only see a bad
generator after the fact; these tests look at it directly: moments, serial correlation, 2-D / 3-D
equidistribution inside a stream (consecutive draws drive one BSDF sample) and ACROSS neighbouring
streams (neighbouring samples of a pixel must not be correlated), with chi-square gates."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

P_MIN = 1e-4   # a correct generator fails one of the ~20 chi-square gates below with probability ~2e-3


@pytest.fixture(scope="module")
def draws(hostcheck):
    hostcheck.hc_rng_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]

    def get(first, n_streams, k, seed):
        out = np.zeros((n_streams, k), np.float32)
        hostcheck.hc_rng_draws(first, n_streams, k, seed, out.ctypes.data)
        return out
    return get


def chi2_p(counts):
    counts = np.asarray(counts, np.float64).ravel()
    e = counts.sum() / counts.size
    x = ((counts - e) ** 2 / e).sum()
    return stats.chi2.sf(x, counts.size - 1)


def test_range_moments_and_serial_correlation(draws):
    u = draws(0, 4096, 256, 1).astype(np.float64)
    assert u.min() >= 0 and u.max() < 1
    n = u.size
    assert abs(u.mean() - 0.5) < 5 * np.sqrt(1 / 12 / n)
    assert abs(u.var() - 1 / 12) < 5 * np.sqrt(1 / 180 / n)
    for lag in (1, 2, 3, 7, 16):       # inside a stream
        r = np.corrcoef(u[:, :-lag].ravel(), u[:, lag:].ravel())[0, 1]
        assert abs(r) < 5 / np.sqrt(n), (lag, r)
    for d in (1, 2, 5):                # same draw index, neighbouring streams
        r = np.corrcoef(u[:-d].ravel(), u[d:].ravel())[0, 1]
        assert abs(r) < 5 / np.sqrt(n), (d, r)


@pytest.mark.parametrize("seed,first", [(1, 0), (12345, 10 ** 9), (2 ** 63 + 5, 2 ** 40)])
def test_equidistribution_inside_a_stream(draws, seed, first):
    u = draws(first, 2048, 384, seed)
    # 1-D, 256 cells
    assert chi2_p(np.bincount((u.ravel() * 256).astype(int), minlength=256)) > P_MIN
    # consecutive pairs and triples: what one BSDF / light sample consumes
    a, b, c = u[:, 0::3].ravel(), u[:, 1::3].ravel(), u[:, 2::3].ravel()
    cell = (a * 64).astype(int) * 64 + (b * 64).astype(int)
    assert chi2_p(np.bincount(cell, minlength=64 * 64)) > P_MIN
    cell = ((a * 16).astype(int) * 16 + (b * 16).astype(int)) * 16 + (c * 16).astype(int)
    assert chi2_p(np.bincount(cell, minlength=16 ** 3)) > P_MIN
    # the low bits of the 24 that a draw keeps (an LCG's weak end)
    low = (u.ravel().astype(np.float64) * 16777216).astype(np.int64) & 255
    assert chi2_p(np.bincount(low, minlength=256)) > P_MIN


@pytest.mark.parametrize("k", [0, 1, 5])
def test_neighbouring_streams_are_not_correlated(draws, k):
    """Samples s and s+1 of a pixel have neighbouring stream indices: their k-th draws (the lens / pixel
    jitter at k = 0, 1) must fill the unit square evenly."""
    u = draws(7_000_000, 400_000, 6, 99)
    a, b = u[0::2, k], u[1::2, k]
    cell = (a * 64).astype(int) * 64 + (b * 64).astype(int)
    assert chi2_p(np.bincount(cell, minlength=64 * 64)) > P_MIN
    # and a stream's draws against its neighbour's NEXT draw (all streams walk one LCG orbit at different offsets)
    if k + 1 < u.shape[1]:
        a, b = u[0::2, k], u[1::2, k + 1]
        cell = (a * 64).astype(int) * 64 + (b * 64).astype(int)
        assert chi2_p(np.bincount(cell, minlength=64 * 64)) > P_MIN


def test_streams_of_different_seeds_differ(draws):
    a, b = draws(0, 1024, 8, 1), draws(0, 1024, 8, 2)
    assert (a != b).mean() > 0.99
    assert np.array_equal(a, draws(0, 1024, 8, 1))     # and a stream depends on (index, seed) only
    assert np.array_equal(draws(100, 4, 8, 1), draws(0, 1024, 8, 1)[100:104])
