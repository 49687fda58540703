"""Kernel (6): batched local-stencil Newton solves + arg-max (TriMesh::computeLocalLDec -> computeLocalEdDec_* ->
nested dense Optimizer, TriMesh.cpp:2105-2794)."""
import numpy as np
import pytest
from stencils import one_ring, port_local_solve
from test_oracle_vs_reference import random_mesh


def _stencils_from_state(s, count, rng):
    out = []
    interior = np.setdiff1d(np.arange(s.nV), s.air["bnd"])
    for v in rng.choice(interior, size=count, replace=False):
        out.append(one_ring(s.F, s.V_rest, s.UV, [int(v)]))
    edges = s.F[rng.choice(s.nF, size=count // 2, replace=False)][:, :2]
    for a, b in edges:                       # two free vertices: the rings of an edge's endpoints
        out.append(one_ring(s.F, s.V_rest, s.UV, [int(a), int(b)]))
    return out


def test_port_local_solve_matches_reference(ref, port, state100):
    """the oracle restatement of the nested Optimizer == the reference's own dense-mode Optimizer"""
    rng = np.random.default_rng(0)
    for Vr, UV, F, free in _stencils_from_state(state100, 12, rng):
        r = ref.local_solve(Vr, F, UV, free)
        p = port_local_solve(port, Vr, F, UV, free)
        assert abs(p["E_init"] - r["E_init"]) <= 1e-13 * r["E_init"]
        assert p["iters"] == r["iters"]
        assert abs(p["E_final"] - r["E_final"]) <= 1e-11 * r["E_final"]
        assert np.max(np.abs(p["UV"] - r["UV"])) <= 1e-9 * np.max(np.abs(r["UV"]))


@pytest.mark.gpu
def test_eval_stencils_matches_oracle(ctx, port, state):
    rng = np.random.default_rng(1)
    st = _stencils_from_state(state, 64, rng)
    scale = rng.uniform(0.5, 1.5, len(st))
    offs = rng.uniform(-1e-4, 0.0, len(st))
    out = ctx.eval_stencils(st, 100, 1e-6, scale, offs)
    assert np.all(out["status"] == 0)
    score = np.zeros(len(st))
    for k, (Vr, UV, F, free) in enumerate(st):
        p = port_local_solve(port, Vr, F, UV, free)
        assert abs(out["E_init"][k] - p["E_init"]) <= 1e-13 * p["E_init"], k
        assert out["iters"][k] == p["iters"], k
        assert abs(out["E_final"][k] - p["E_final"]) <= 1e-9 * p["E_final"], k         # north_star: 1e-9 relative
        assert np.max(np.abs(out["UV"][k] - p["UV"])) <= 1e-8 * np.max(np.abs(p["UV"])), k
        assert np.array_equal(out["UV"][k][~free], UV[~free])                             # fixed vertices untouched
        score[k] = scale[k] * (p["E_init"] - p["E_final"]) + offs[k]
    assert np.max(np.abs(out["score"] - score)) <= 1e-9 * np.max(np.abs(score))
    assert out["argmax"] == int(np.argmax(score))                                         # identical decision


@pytest.mark.gpu
def test_eval_stencils_edge_cases(ctx):
    assert ctx.eval_stencils([])["argmax"] == -1                                         # empty batch
    V_rest, F, UV = random_mesh(5, n=4)
    iso = (np.column_stack([V_rest[:, :2], np.zeros(len(UV))]), V_rest[:, :2].copy(), F, np.isin(np.arange(len(UV)), [5]))
    out = ctx.eval_stencils([iso, iso])
    assert np.all(out["iters"] == 1) and np.allclose(out["E_init"], 4.0) and np.allclose(out["E_final"], 4.0)   # isometry: converged at once
    assert out["argmax"] == 0                                                              # tie -> first maximum
    big = random_mesh(6, n=12)                                                             # 242 triangles: over the per-stencil limit
    out = ctx.eval_stencils([(big[0], big[2], big[1], np.ones(len(big[2]), bool))])
    assert out["status"][0] == -2 and out["argmax"] == -1 or out["score"][0] == -np.inf
    flipped = (iso[0], iso[1] * np.array([1.0, -1.0]), F, iso[3])                          # inverted input
    assert ctx.eval_stencils([flipped])["status"][0] == -4
