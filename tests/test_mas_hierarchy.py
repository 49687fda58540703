"""Host logic of the solver's multilevel preconditioner (optcuts_b200/csrc/ocb_mas.cu): the row order and the
node hierarchy built at pattern time.  No GPU: ocb_precond_hierarchy does no CUDA work."""
import numpy as np
import pytest

from optcuts_b200 import _capi


def _check(xy, grid):
    n = len(xy)
    vert_of, levels, lloc = _capi.precond_hierarchy(xy, grid)
    assert sorted(vert_of.tolist()) == list(range(n))                  # a permutation
    rows_per = -(-n // grid)
    sizes = [len(l) - 1 for l in levels]
    # level 1: leaves of <= 8 consecutive rows that never straddle a CTA boundary
    cb = levels[0]
    assert cb[0] == 0 and cb[-1] == n and np.all(np.diff(cb) >= 1) and np.all(np.diff(cb) <= 8)
    assert np.all(cb[:-1] // rows_per == (cb[1:] - 1) // rows_per)
    # upper levels: groups of <= 8 consecutive nodes covering the level below
    for l in range(1, len(levels)):
        cb = levels[l]
        assert cb[0] == 0 and cb[-1] == sizes[l - 1] and np.all(np.diff(cb) >= 1) and np.all(np.diff(cb) <= 8)
    # the last level is the coarse one: the first level small enough for the exact dense inverse (the cap grows with the
    # system: n/8 DOFs, between 600 and 3072), or the level with one node per CTA if none is
    cap = min(3072, max(600, n // 8))
    assert 6 * sizes[-1] <= 3072 and lloc == len(levels)
    assert 6 * sizes[-1] <= cap or sizes[-1] == -(-n // rows_per)
    assert all(6 * s > cap for s in sizes[:-1])
    # no node of any level straddles a CTA boundary: map every node to its row range
    lo, hi = levels[0][:-1].copy(), levels[0][1:].copy()
    for l in range(1, len(levels)):
        cb = levels[l]
        lo, hi = lo[cb[:-1]], hi[cb[1:] - 1]
        assert np.all(lo // rows_per == (hi - 1) // rows_per)
    return vert_of, levels, lloc


@pytest.mark.parametrize("n,grid", [(16, 1), (100, 1), (5358, 16), (5358, 1), (20000, 148), (1000, 148), (131, 16), (80000, 148), (300000, 148)])
def test_hierarchy_invariants(n, grid):
    rng = np.random.default_rng(n)
    _check(rng.uniform(-1, 1, (n, 2)), grid)


def test_leaves_are_spatially_compact():
    """Recursive coordinate bisection: a leaf of a uniform point cloud is far smaller than the cloud."""
    rng = np.random.default_rng(1)
    xy = rng.uniform(0, 1, (8000, 2))
    vert_of, levels, _ = _check(xy, 16)
    cb = levels[0]
    ext = [np.ptp(xy[vert_of[cb[k]:cb[k + 1]]], axis=0).max() for k in range(len(cb) - 1)]
    assert np.median(ext) < 0.08        # ~ sqrt(8/8000) * a small factor


def test_degenerate_points():
    """Coincident / non-finite coordinates must not break the bisection."""
    xy = np.zeros((300, 2))
    xy[::7] = np.nan
    _check(xy, 4)
