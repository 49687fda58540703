"""Pins the C restatement (oracle/port) against the REAL reference classes (oracle/_ref/liboptcuts_ref.so,
compiled from the unmodified reference by oracle/Makefile) on seeded random inputs and edge cases.
Skipped when oracle/_ref is not built.  CPU only."""
import numpy as np
import pytest
from conftest import relerr


def random_mesh(seed, n=9, jitter=0.25, flat=False):
    """n x n grid triangulated, jittered; UV = a smooth, orientation-preserving distortion of the rest shape."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(n, dtype=float), np.arange(n, dtype=float), indexing="ij")
    P = np.stack([xs.ravel(), ys.ravel()], axis=1)
    P += jitter * rng.uniform(-1, 1, P.shape)
    z = np.zeros(len(P)) if flat else 0.3 * np.sin(P[:, 0]) * np.cos(0.7 * P[:, 1])
    V_rest = np.column_stack([P, z])
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    F = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    A = np.array([[1.3, 0.2], [-0.1, 0.8]])
    UV = P @ A.T + 0.05 * np.sin(P[:, ::-1])
    return V_rest, F, UV


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_element_functions(ref, port, seed):
    V_rest, F, UV = random_mesh(seed)
    m = ref.RefMesh(V_rest, F, UV)
    rest8, sc = m.features()
    r8, psc, rc = port.rest_features(V_rest, F)
    assert rc == 0 and np.array_equal(r8, rest8)
    surf = sc["surfaceArea"]
    assert abs(psc["surfaceArea"] - surf) <= 1e-14 * surf
    for uniform in (False, True):
        assert np.array_equal(port.energy_per_elem(F, UV, rest8, surf, uniform), m.energy_per_elem(uniform))
        assert relerr(port.gradient(F, UV, rest8, surf, uniform), m.gradient(uniform)) < 1e-14
        I, J, V = m.hessian_triplets(uniform)
        pI, pJ, pV = port.hessian_triplets(F, UV, rest8, surf, uniform)
        assert np.array_equal(I, pI) and np.array_equal(J, pJ)
        assert np.max(np.abs(V - pV)) <= 1e-12 * np.max(np.abs(V))
    assert relerr(port.divgrad(F, UV, rest8, surf), m.divgrad()) < 1e-12
    rng = np.random.default_rng(seed + 100)
    for scale in (0.05, 1.0, 20.0):
        p = scale * rng.standard_normal(2 * len(UV))
        assert port.init_step_size(F, UV, p, 1.0) == m.init_step_size(p, 1.0)     # SymDirichletEnergy.cpp:551-610
    m.close()


def test_make_pd6_random_symmetric(ref, port):
    rng = np.random.default_rng(7)
    for k in range(200):
        B = rng.standard_normal((6, 6))
        M = B + B.T
        if k % 4 == 0:
            M = B @ B.T                                       # PSD: must come back untouched (IglUtils.hpp:74-76)
        R, P = ref.make_pd6(M), port.make_pd6(M)
        if k % 4 == 0:
            assert np.array_equal(P, M) and np.array_equal(R, M)
        assert np.max(np.abs(R - P)) <= 1e-13 * np.max(np.abs(M))
        assert np.min(np.linalg.eigvalsh(0.5 * (P + P.T))) > -1e-12 * np.max(np.abs(M))


def test_pattern_update_solve(ref, port):
    V_rest, F, UV = random_mesh(3, n=12)
    m = ref.RefMesh(V_rest, F, UV)
    rest8, sc = m.features()
    ptr, idx = m.adjacency()
    s = ref.RefSolver()
    for fixed in ([0], [0, 5, 77], []):
        s.set_pattern(ptr, idx, fixed)
        ia, ja, _ = s.csr()
        pia, pja = port.set_pattern(ptr, idx, fixed)
        # LinSysSolver.hpp:37-135 builds 1-based ia/ja; EigenLibSolver::set_pattern (EigenLibSolver.cpp:21-39)
        # then shifts them to 0-based in place, which is what the read-back sees
        assert np.array_equal(ia + 1, pia) and np.array_equal(ja + 1, pja)
    s.set_pattern(ptr, idx, [0])
    I, J, V = m.hessian_triplets(False)
    s.update_a(I, J, V)
    ia, ja, a = s.csr()
    pia, pja = port.set_pattern(ptr, idx, [0])
    pa, miss = port.update_a(pia, pja, I, J, V)
    assert miss == 0 and np.array_equal(pa, a)                           # LinSysSolver.hpp:138-159
    assert s.factorize()
    rhs = np.random.default_rng(0).standard_normal(len(pia) - 1)
    x = s.solve(rhs)
    px, rc = port.ldlt_solve(pia, pja, pa, rhs)
    assert rc == 0 and np.linalg.norm(px - x) <= 1e-10 * np.linalg.norm(x)
    s.close(); m.close()


def test_newton_iterations_no_scaffold(ref, port):
    """Five free-running reference Newton iterations (no scaffold) reproduced by the port."""
    V_rest, F, UV = random_mesh(4, n=10)
    m = ref.RefMesh(V_rest, F, UV)
    rest8, sc = m.features()
    opt = ref.RefOptimizer(m, 0.9, scaffolding=False, mute=True)
    tg = opt.scalars()["targetGRes"]
    uv = UV.copy()
    for it in range(5):
        opt.solve(1)
        uv, _, p, r = port.newton_step(F, uv, rest8, sc["surfaceArea"], [0], 0.9, tg)
        rs = opt.scalars()
        assert abs(r["E_new"] - rs["lastEnergyVal"]) <= 1e-9 * rs["lastEnergyVal"]
        assert np.max(np.abs(uv - opt.uv())) <= 1e-8 * np.max(np.abs(uv))
    opt.close(); m.close()


def test_scaffold_step_matches(ref, port, state1):
    """Reference Scaffold + one Newton iteration from the golden Tutte state == port.newton_step."""
    s = state1
    m = ref.RefMesh(s.V_rest, s.F, s.UV)
    sc = ref.build_scaffold(m)
    assert np.array_equal(sc["V"], s.air["V"]) and np.array_equal(sc["F"], s.air["F"])      # Triangle is deterministic
    assert np.array_equal(sc["rest8"], s.air["rest8"])
    m.close()
