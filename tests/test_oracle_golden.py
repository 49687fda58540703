"""Pins the C restatement (oracle/port) against golden vectors recorded from the UNMODIFIED reference
(tests/golden/make_golden.py: the reference program run on bimba_i_f10000, BASELINE.json configs[1],
plus one replayed Newton iteration through the real OptCuts::Optimizer at iterations 1 and 100).
CPU only."""
import numpy as np
from conftest import relerr


def test_rest_features_bit_exact(port, state):
    r8, sc, rc = port.rest_features(state.V_rest, state.F)
    assert rc == 0
    assert np.array_equal(r8, state.rest8)                      # TriMesh.cpp:343-398
    assert abs(sc["surfaceArea"] - state.surfaceArea) <= 1e-13 * state.surfaceArea
    assert abs(sc["avgEdgeLen"] - state.avgEdgeLen) <= 1e-13 * state.avgEdgeLen
    assert abs(sc["virtualRadius"] - state.virtualRadius) <= 1e-13 * state.virtualRadius


def test_air_rest_features_with_clamp(port, state):
    a = state.air
    r8, sc, rc = port.rest_features(a["V"], a["F"], a["areaThres_AM"])
    assert np.array_equal(r8, a["rest8"])                        # incl. TriMesh.cpp:373-383 surrogate


def test_energy_per_elem_bit_exact(port, state1):
    per = port.energy_per_elem(state1.F, state1.UV, state1.rest8, state1.surfaceArea)
    assert np.array_equal(per, state1.r("energy_per_elem"))      # SymDirichletEnergy.cpp:24-46


def test_total_energy(port, state):
    a = state.air
    esd = port.energy(state.F, state.UV, state.rest8, state.surfaceArea)
    escaf = port.energy(a["F"], a["V"], a["rest8"], 1.0, uniform=True) * state.w_scaf / a["F"].shape[0]
    assert abs(esd - float(state.r("E_sd_last"))) <= 1e-13 * esd
    assert abs(escaf - float(state.r("E_scaf_last"))) <= 1e-13 * escaf
    assert abs(state.p0 * esd + escaf - float(state.r("E_last"))) <= 1e-13 * float(state.r("E_last"))


def _port_gradient(port, s):
    a = s.air
    g = np.zeros(2 * (s.nV + a["V"].shape[0] - a["nBnd"]))
    g[:2 * s.nV] = s.p0 * port.gradient(s.F, s.UV, s.rest8, s.surfaceArea, fixed=s.fixed)
    ga = port.gradient(a["F"], a["V"], a["rest8"], 1.0, uniform=True, fixed=a["fixed"]) * (s.w_scaf / a["F"].shape[0])
    l2g = a["localVI2Global"]
    np.add.at(g, 2 * l2g, ga[0::2])
    np.add.at(g, 2 * l2g + 1, ga[1::2])
    return g


def test_gradient(port, state):
    g = _port_gradient(port, state)
    assert relerr(g, state.r("gradient")) < 1e-14                # Optimizer.cpp:783-797, Scaffold.cpp:210-229
    assert abs(g @ g - float(state.r("sqn_g"))) <= 1e-13 * float(state.r("sqn_g"))


def test_divgrad(port, state):
    d = port.divgrad(state.F, state.UV, state.rest8, state.surfaceArea)
    assert relerr(d, state.r("divgrad")) < 1e-12                 # SymDirichletEnergy.cpp:108-149


def test_seam_sparsity(port, state100):
    s = state100
    for soup, key in ((False, "seam_sparsity"), (True, "seam_sparsity_soup")):
        v = port.seam_sparsity(s.cohE, s.r("boundaryEdge"), s.r("edgeLen"), s.UV, s.avgEdgeLen, 0.0, soup)
        assert abs(v - float(s.r(key))) <= 1e-13 * max(1.0, abs(float(s.r(key))))


def _merged_adjacency(s):
    from optcuts_b200.trimesh import adjacency_from_faces, merge_adjacency
    a = s.air
    nVtot = s.nV + a["V"].shape[0] - a["nBnd"]
    return merge_adjacency(adjacency_from_faces(s.F, s.nV), s.nV, a["localVI2Global"][a["F"]], nVtot), nVtot


def test_pattern_and_matrix_vs_reference_csr(port, state1):
    """LinSysSolver::set_pattern + update_a on the port's triplets == the reference's ia/ja/a."""
    s = state1
    (ptr, idx), nVtot = _merged_adjacency(s)
    ia, ja = port.set_pattern(ptr, idx, s.fixed)
    # the golden ia/ja were read back after EigenLibSolver::set_pattern made them 0-based (EigenLibSolver.cpp:21-39)
    assert np.array_equal(ia, s.r("ia") + 1) and np.array_equal(ja, s.r("ja") + 1)
    a = s.air
    I, J, V = port.hessian_triplets(s.F, s.UV, s.rest8, s.surfaceArea, fixed=s.fixed)
    Ia, Ja, Va = port.hessian_triplets(a["F"], a["V"], a["rest8"], 1.0, uniform=True, fixed=a["fixed"])
    l2g = a["localVI2Global"]
    Ia, Ja = l2g[Ia // 2] * 2 + Ia % 2, l2g[Ja // 2] * 2 + Ja % 2
    vals, miss = port.update_a(ia, ja, np.concatenate([I, Ia]), np.concatenate([J, Ja]),
                               np.concatenate([s.p0 * V, s.w_scaf / a["F"].shape[0] * Va]))
    assert miss == 0
    ref_a = s.r("a")
    # projected element blocks agree with Eigen's eigen-solver to rounding, relative to the row's diagonal scale
    assert np.max(np.abs(vals - ref_a)) <= 1e-11 * np.max(np.abs(ref_a))
    assert relerr(vals, ref_a) < 1e-11


def test_ldlt_search_direction(port, state1):
    s = state1
    x, rc = port.ldlt_solve(s.r("ia") + 1, s.r("ja") + 1, s.r("a"), -s.r("gradient"))
    assert rc == 0
    p = s.r("searchDir")
    assert np.linalg.norm(x - p) <= 1e-9 * np.linalg.norm(p)     # SimplicialLDLT (AMD) vs skyline LDL^T (RCM)


def test_newton_step_reproduces_reference_iteration(port, state):
    s = state
    air = dict(s.air, V=s.air["V"].copy(order="F"))
    UV1, UVa1, p, r = port.newton_step(s.F, s.UV, s.rest8, s.surfaceArea, s.fixed, s.p0, float(s.r("targetGRes")),
                                       air=air, w_scaf=s.w_scaf)
    assert abs(r["sqn_g"] - float(s.r("sqn_g"))) <= 1e-12 * float(s.r("sqn_g"))
    assert abs(r["alpha"] - float(s.r("alpha"))) <= 1e-9 * float(s.r("alpha"))
    assert abs(r["E_new"] - float(s.r("E_new"))) <= 1e-9 * float(s.r("E_new"))     # north_star tolerance
    assert abs(r["E_sd_new"] - float(s.r("E_sd_new"))) <= 1e-9 * float(s.r("E_sd_new"))
    assert abs(r["E_scaf_new"] - float(s.r("E_scaf_new"))) <= 1e-9 * float(s.r("E_scaf_new"))
    assert abs(r["lastEDec"] - float(s.r("lastEDec"))) <= 1e-7 * abs(float(s.r("lastEDec")))
    assert np.linalg.norm(p - s.r("searchDir")) <= 1e-8 * np.linalg.norm(s.r("searchDir"))
    assert np.max(np.abs(UV1 - s.next_uv())) <= 1e-9 * np.max(np.abs(s.next_uv()))


def test_trace_fixture_end_state():
    """The committed traces end where SURVEY.md §8c says the reference ends (523 / 170 iterations)."""
    import os
    from conftest import GOLDEN
    for name, iters, esd, ese in (("bimba_cfg1", 523, 4.09986, 3.88828), ("bimba_cfg2", 170, 4.28086, 2.65232)):
        lines = open(os.path.join(GOLDEN, name + "_trace.txt")).read().strip().split("\n")
        assert len(lines) == iters
        info = open(os.path.join(GOLDEN, name + "_info.txt")).read().split("\n")
        assert int(info[1].split()[0]) == iters
        e = [float(v) for v in info[3].split()]
        assert abs(e[0] - esd) < 1e-5 and abs(e[1] - ese) < 1e-5
