"""CPU: the host side of the direct safety net (csrc/ocb_direct.cu): the breadth-first level order and the blocks of the
block-tridiagonal Cholesky.  The factorisation is only correct if EVERY coupling of the matrix joins equal or adjacent blocks; this
is checked on the bimba mesh's vertex graph (the reference's golden state), on a graph with several components and isolated
vertices (fixed rows are isolated in the solver's pattern), and on a path / a star (degenerate level structures)."""
import numpy as np
import pytest
from conftest import GOLDEN  # noqa: F401

from optcuts_b200 import _capi


def csr_of_edges(n, edges, diagonal=True):
    nb = [set() for _ in range(n)]
    for a, b in edges:
        nb[a].add(b); nb[b].add(a)
    if diagonal:
        for v in range(n):
            nb[v].add(v)
    ptr = np.zeros(n + 1, np.int32)
    idx = []
    for v in range(n):
        row = list(nb[v])
        row.reverse()                                   # rows need not be sorted
        idx.extend(row)
        ptr[v + 1] = len(idx)
    return ptr, np.asarray(idx, np.int32)


def check(n, ptr, idx, target):
    pos, blk, beg = _capi.direct_level_blocks(ptr, idx, target)
    assert sorted(pos.tolist()) == list(range(n)), "pos is not a permutation"
    assert beg[0] == 0 and beg[-1] == n and np.all(np.diff(beg) > 0)
    for v in range(n):
        assert beg[blk[v]] <= pos[v] < beg[blk[v] + 1]
        for q in range(ptr[v], ptr[v + 1]):
            assert abs(int(blk[v]) - int(blk[idx[q]])) <= 1, "coupling (%d, %d) joins blocks %d and %d" % (v, idx[q], blk[v], blk[idx[q]])
    return beg


def test_mesh_graph_is_block_tridiagonal(golden):
    F = golden["s1_F"]
    n = int(F.max()) + 1
    edges = [(int(t[i]), int(t[(i + 1) % 3])) for t in F for i in range(3)]
    ptr, idx = csr_of_edges(n, edges)
    for target in (16, 192, 100000):
        beg = check(n, ptr, idx, target)
        sizes = np.diff(beg)
        assert np.all(sizes[:-1] >= min(target, n)) or len(sizes) == 1          # blocks are whole levels merged up to the target
    # a 5k-vertex disk-like mesh: the levels stay narrow, so the chain is long and the blocks small
    beg = check(n, ptr, idx, 192)
    assert len(beg) - 1 >= 8 and np.diff(beg).max() < 1500


def test_components_isolated_vertices_and_degenerate_graphs():
    rng = np.random.default_rng(0)
    # two grids + isolated vertices, vertex ids shuffled
    def grid(m, off):
        return [(off + i * m + j, off + i * m + j + 1) for i in range(m) for j in range(m - 1)] + [(off + i * m + j, off + (i + 1) * m + j) for i in range(m - 1) for j in range(m)]
    n = 20 * 20 + 7 * 7 + 5
    perm = rng.permutation(n)
    edges = [(int(perm[a]), int(perm[b])) for a, b in grid(20, 0) + grid(7, 400)]
    ptr, idx = csr_of_edges(n, edges)
    check(n, ptr, idx, 24)
    check(n, ptr, idx, 1)
    # a path (levels of one vertex) and a star (one huge level)
    ptr, idx = csr_of_edges(50, [(i, i + 1) for i in range(49)])
    beg = check(50, ptr, idx, 8)
    assert np.all(np.diff(beg)[:-1] == 8)
    ptr, idx = csr_of_edges(40, [(0, i) for i in range(1, 40)])
    check(40, ptr, idx, 4)


def test_bad_arguments_are_statuses():
    ptr, idx = csr_of_edges(4, [(0, 1), (2, 3)])
    bad = idx.copy(); bad[0] = 99
    with pytest.raises(_capi.OcbError):
        _capi.direct_level_blocks(ptr, bad, 4)
    with pytest.raises(_capi.OcbError):
        _capi.direct_level_blocks(ptr, idx, 0)
