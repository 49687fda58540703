"""GPU: run-to-run reproducibility.  The reference is byte-reproducible (every parallel_for writes disjoint
slots, all reductions are serial); so is this path: the gradient and the matrix are assembled by vertex / block-row
gathers in a fixed order, the preconditioner's Galerkin products are gathers too, and the PCG reductions add in a fixed
order -- no floating-point atomics anywhere.  Five repetitions (and a second context) must give the same BITS."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _assemble(ctx, s):
    s.upload(ctx)
    g, sq = ctx.gradient(s.p0)
    ctx.set_pattern_from_elements()
    ctx.hessian_assemble(s.p0)
    ia, ja, a = ctx.download_csr()
    return g, sq, ia, ja, a


def test_gradient_and_matrix_are_bitwise_reproducible(ctx, state):
    ref = _assemble(ctx, state)
    for _ in range(4):
        cur = _assemble(ctx, state)
        for x, y in zip(ref, cur):
            assert np.array_equal(x, y)


def test_search_direction_is_bitwise_reproducible(ctx, state):
    import optcuts_b200 as ob
    out = []
    for rep in range(5):
        c = ctx if rep < 4 else ob.Context(0)          # the last repetition in a fresh context
        try:
            _assemble(c, state)
            c.factorize()
            p, info = c.solve(None, 1e-12, 0)
            out.append((p.copy(), info["iters"], info["rel_res"]))
        finally:
            if c is not ctx:
                c.close()
    for p, it, rr in out[1:]:
        assert it == out[0][1] and rr == out[0][2]
        assert np.array_equal(p, out[0][0])


def test_newton_trajectory_is_bitwise_reproducible(ctx, state1):
    """three free-running Newton iterations, twice: same energies, same UVs, same CG iteration counts"""
    runs = []
    for _ in range(2):
        state1.upload(ctx)
        ctx.set_pattern_from_elements()
        tr = []
        for _ in range(3):
            r = ctx.newton_step(state1.p0, 0.0)
            tr.append((r["E_new"], r["alpha"], r["pcg_iters"], r["sqn_g"]))
        V, Va = ctx.get_uv(want_air=True)
        runs.append((tr, V.copy(), Va.copy()))
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])


def test_divgrad_scores_are_bitwise_reproducible_and_match_the_port(ctx, port, state100):
    s = state100
    s.upload(ctx, with_air=False)
    d0 = ctx.divgrad_scores()
    for _ in range(3):
        assert np.array_equal(ctx.divgrad_scores(), d0)
    ref = port.divgrad(s.F, s.UV, s.rest8, s.surfaceArea)
    # same operations in the same order as the reference's two serial loops (SymDirichletEnergy.cpp:108-149)
    assert np.array_equal(d0, ref), "max abs diff %g" % np.max(np.abs(d0 - ref))
