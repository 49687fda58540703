"""The PRODUCT's device-inline element math (optcuts_b200/csrc/ocb_element.cuh) compiled for the host
(tests/harness/element_host.cpp) and checked against the oracle, so formula regressions are caught
without a GPU.  The harness is test-only; it is never shipped or imported by the package."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from conftest import ROOT, relerr
from test_oracle_vs_reference import random_mesh

_d, _i = C.POINTER(C.c_double), C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = tmp_path_factory.mktemp("harness") / "libelement_host.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", str(out),
                           os.path.join(ROOT, "tests", "harness", "element_host.cpp")])
    L = C.CDLL(str(out))
    L.host_step_bound.restype = C.c_double
    return L


def _args(F, UV, rest8):
    F, UV, rest8 = np.asfortranarray(F, np.int32), np.asfortranarray(UV, np.float64), np.ascontiguousarray(rest8)
    return F, UV, rest8, (UV.shape[0], F.shape[0], F.ctypes.data_as(_i), UV.ctypes.data_as(_d), rest8.ctypes.data_as(_d))


def test_energy_and_gradient_bit_exact(host, port, state):
    F, UV, rest8, a = _args(state.F, state.UV, state.rest8)
    out = np.zeros(state.nF)
    host.host_energy(*a, C.c_double(state.surfaceArea), 0, out.ctypes.data_as(_d))
    assert np.array_equal(out, port.energy_per_elem(F, UV, rest8, state.surfaceArea))
    gc = np.zeros((state.nF, 6))
    host.host_gradient_corners(*a, C.c_double(state.surfaceArea), 0, gc.ctypes.data_as(_d))
    g = np.zeros(2 * state.nV)
    for k in range(3):
        np.add.at(g, 2 * F[:, k], gc[:, 2 * k])
        np.add.at(g, 2 * F[:, k] + 1, gc[:, 2 * k + 1])
    g[2 * state.fixed] = 0
    g[2 * state.fixed + 1] = 0
    assert relerr(g, port.gradient(F, UV, rest8, state.surfaceArea, fixed=state.fixed)) < 1e-14


def test_gather_gradient_is_bit_identical_to_the_reference(host, state):
    """The vertex-gather assembly (grad_gather_kernel) sums, per vertex, the corner gradients in ascending triangle order
    and combines  energyParam0 * g_mesh + w_scaf/|Fa| * g_air  like Optimizer::computeGradient + Scaffold::augmentGradient:
    restated on the host with the product's own element function it reproduces the gradient recorded from the reference
    BIT FOR BIT (mesh + air mesh, both golden states)."""
    F, UV, rest8, a = _args(state.F, state.UV, state.rest8)
    host.host_corner_vs_triangle.restype = C.c_long
    assert host.host_corner_vs_triangle(*a, C.c_double(state.surfaceArea), 0) == 0
    gm = np.zeros(2 * state.nV)
    host.host_gather_gradient(*a, C.c_double(state.surfaceArea), 0, gm.ctypes.data_as(_d))
    gm[2 * state.fixed] = 0
    gm[2 * state.fixed + 1] = 0
    air = state.air
    Fa, Va, r8a, aa = _args(air["F"], air["V"], air["rest8"])
    ga = np.zeros(2 * Va.shape[0])
    host.host_gather_gradient(*aa, C.c_double(1.0), 1, ga.ctypes.data_as(_d))
    w = state.w_scaf / Fa.shape[0]
    nB, nVa = air["nBnd"], Va.shape[0]
    g = np.zeros(2 * (state.nV + nVa - nB))
    g[:2 * state.nV] = state.p0 * gm
    l2g = air["localVI2Global"]
    for i in range(nB):
        g[2 * l2g[i]] += w * ga[2 * i]
        g[2 * l2g[i] + 1] += w * ga[2 * i + 1]
    g[2 * state.nV:] = w * ga[2 * nB:]
    ref = state.r("gradient")
    assert g.shape == ref.shape
    assert np.array_equal(g, ref), "max abs diff %g" % np.max(np.abs(g - ref))


def test_divgrad_gather_is_bit_identical_to_the_port(host, port):
    """divgrad_gather_kernel's arithmetic (per vertex: mean of the incident corner gradients, then squared deviations, both in
    ascending triangle order) against the oracle's restatement of computeLocalGradient + computeDivGradPerVert."""
    V_rest, F, UV = random_mesh(5, n=12)
    rest8, sc, _ = port.rest_features(V_rest, F)
    F, UV, rest8, a = _args(F, UV, rest8)
    out = np.zeros(UV.shape[0])
    host.host_divgrad(*a, C.c_double(sc["surfaceArea"]), out.ctypes.data_as(_d))
    assert np.array_equal(out, port.divgrad(F, UV, rest8, sc["surfaceArea"]))


def test_hessian_projection_matches_makePD(host, port, state):
    """closed-form 4x4 projection in the translation-free basis == eigen-clamp of the 6x6"""
    F, UV, rest8, a = _args(state.F, state.UV, state.rest8)
    for project in (0, 1):
        H = np.zeros((state.nF, 36))
        cl = np.zeros(state.nF, np.int32)
        host.host_hessian_blocks(*a, C.c_double(state.surfaceArea), 0, project, H.ctypes.data_as(_d), cl.ctypes.data_as(_i))
        Hp = port.hessian_blocks(F, UV, rest8, state.surfaceArea, False, bool(project)).reshape(state.nF, 36)
        scale = np.abs(Hp).max(axis=1)
        assert np.max(np.abs(H - Hp).max(axis=1) / scale) < 5e-13
    if state.tag == "s1_":
        assert (cl > 0).sum() > 100          # the Tutte start has many indefinite element Hessians


def test_projection_fast_path_covers_the_indefinite_elements(host, state):
    """the rank-1 (smallest eigenpair) projection must handle practically every indefinite element: a fallback to
    the general eigen-solve stalls the 31 other lanes of its warp"""
    F, UV, rest8, a = _args(state.F, state.UV, state.rest8)
    H = np.zeros((state.nF, 36))
    cl = np.zeros(state.nF, np.int32)
    host.host_general_path.restype = C.c_long
    host.host_general_path(1)
    host.host_hessian_blocks(*a, C.c_double(state.surfaceArea), 0, 1, H.ctypes.data_as(_d), cl.ctypes.data_as(_i))
    general = host.host_general_path(1)
    assert general <= max(1, (cl > 0).sum() // 1000), (general, (cl > 0).sum())


def test_projection_on_synthetic_spectra(host):
    """sd_project_psd against a numpy eigen-clamp on translation-free 6x6 matrices with 0..4 negative eigenvalues,
    clustered / repeated / tiny eigenvalues and bad scaling: the rank-1 fast path (exactly one negative eigenvalue) and
    the general fallback must both reproduce makePD (IglUtils.hpp:71-90)"""
    rng = np.random.default_rng(7)
    s2, s6 = np.sqrt(0.5), 1.0 / np.sqrt(6.0)
    W = np.kron(np.array([[s2, s6], [-s2, s6], [0.0, -2.0 * s6]]), np.eye(2))      # 6x4 basis of the translation-free space
    mats, nneg = [], []
    spectra = [
        lambda: rng.uniform(0.1, 10.0, 4) * np.array([-1, 1, 1, 1]),                # the common case
        lambda: rng.uniform(0.1, 10.0, 4) * np.array([-1, -1, 1, 1]),
        lambda: rng.uniform(0.1, 10.0, 4) * np.array([-1, -1, -1, 1]),
        lambda: -rng.uniform(0.1, 10.0, 4),
        lambda: rng.uniform(0.1, 10.0, 4),
        lambda: np.array([-1e-9, 1.0, 1.0, 2.0]) * rng.uniform(0.5, 2.0),           # tiny negative, repeated positive
        lambda: np.array([-3.0, 1e-12, 1.0, 1e4]) * rng.uniform(0.5, 2.0),          # second eigenvalue ~ 0: hard for a naive RQI
        lambda: np.array([-1e6, 1e-3, 1.0, 1e6]) * rng.uniform(0.5, 2.0),           # 12 decades
        lambda: np.array([-2.0, -2.0, 3.0, 3.0]),                                   # repeated negative
        lambda: np.array([-1.0, 1.0, 1.0, 1.0]) * 10.0 ** rng.uniform(-8, 8),       # scaling
    ]
    for make in spectra:
        for _ in range(200):
            lam = make()
            Q, _ = np.linalg.qr(rng.standard_normal((4, 4)))
            M = (Q * lam) @ Q.T
            mats.append(W @ (0.5 * (M + M.T)) @ W.T)
            nneg.append(int((lam < 0).sum()))
    H = np.ascontiguousarray(np.array(mats).reshape(-1, 36))
    want = []
    for A in mats:
        w, V = np.linalg.eigh(A)
        want.append((V * np.maximum(w, 0.0)) @ V.T if w[0] < 0 else A)          # null space of W W^T stays null either way
    cl = np.zeros(len(mats), np.int32)
    host.host_general_path.restype = C.c_long
    host.host_general_path(1)
    host.host_project_psd(len(mats), H.ctypes.data_as(_d), cl.ctypes.data_as(_i))
    general = host.host_general_path(1)
    got = H.reshape(-1, 6, 6)
    scale = np.array([np.abs(A).max() for A in mats])
    err = np.array([np.abs(g - w_).max() for g, w_ in zip(got, want)]) / scale
    assert err.max() < 5e-12, (err.max(), int(err.argmax()), nneg[int(err.argmax())])
    for g, sc in zip(got, scale):
        assert np.linalg.eigvalsh(g)[0] > -1e-10 * sc
    nneg = np.array(nneg)
    assert np.array_equal(cl[nneg != 1] > 0, nneg[nneg != 1] > 0)
    # every matrix with 2+ negative eigenvalues must have gone through the general path; single-negative ones mostly not
    assert general >= int((nneg >= 2).sum())
    assert general <= int((nneg >= 2).sum()) + int(0.35 * (nneg == 1).sum())


def test_isometric_known_answer(host):
    """SymDirichletEnergy::checkEnergyVal (SymDirichletEnergy.cpp:612-645): isometry => E_t = 4 w, zero gradient,
    PSD Hessian with a 3-dimensional null space (2 translations + rotation)."""
    V_rest, F, _ = random_mesh(11, n=5, flat=True)
    UV = V_rest[:, :2].copy()
    from oracle import portapi
    rest8, sc, _ = portapi.rest_features(V_rest, F)
    F, UV, rest8, a = _args(F, UV, rest8)
    nF = F.shape[0]
    out = np.zeros(nF)
    host.host_energy(*a, C.c_double(sc["surfaceArea"]), 0, out.ctypes.data_as(_d))
    assert np.max(np.abs(out - 4.0 * rest8[0] / sc["surfaceArea"])) < 1e-14
    gc = np.zeros((nF, 6))
    host.host_gradient_corners(*a, C.c_double(sc["surfaceArea"]), 1, gc.ctypes.data_as(_d))
    assert np.max(np.abs(gc)) < 1e-12
    H = np.zeros((nF, 36))
    host.host_hessian_blocks(*a, C.c_double(1.0), 1, 1, H.ctypes.data_as(_d), None)
    for t in range(nF):
        ev = np.linalg.eigvalsh(H[t].reshape(6, 6))
        assert ev[0] > -1e-10 and np.sum(np.abs(ev) < 1e-9 * ev[-1]) == 3


def test_finite_difference_gradient_and_hessian(host, port):
    """Energy::checkGradient / checkHessian logic (Energy.cpp:42-147) on the unprojected blocks."""
    V_rest, F, UV = random_mesh(12, n=4)
    rest8, sc, _ = port.rest_features(V_rest, F)
    F, UV, rest8, a = _args(F, UV, rest8)
    nF, nV = F.shape[0], UV.shape[0]
    surf = sc["surfaceArea"]
    gc = np.zeros((nF, 6)); H = np.zeros((nF, 36))
    host.host_gradient_corners(*a, C.c_double(surf), 0, gc.ctypes.data_as(_d))
    host.host_hessian_blocks(*a, C.c_double(surf), 0, 0, H.ctypes.data_as(_d), None)
    h = 1e-6
    for t in (0, 7, nF - 1):
        for k in range(3):
            for c in range(2):
                def at(delta):
                    U2 = UV.copy(order="F"); U2[F[t, k], c] += delta
                    e = port.energy_per_elem(F, U2, rest8, surf)[t]
                    g2 = np.zeros((nF, 6))
                    F2, U2, r2, a2 = _args(F, U2, rest8)
                    host.host_gradient_corners(*a2, C.c_double(surf), 0, g2.ctypes.data_as(_d))
                    return e, g2[t]
                ep, gp = at(h); em, gm = at(-h)
                assert abs((ep - em) / (2 * h) - gc[t, 2 * k + c]) < 1e-6 * max(1.0, abs(gc[t]).max())
                assert np.max(np.abs((gp - gm) / (2 * h) - H[t].reshape(6, 6)[2 * k + c])) < 1e-5 * max(1.0, np.abs(H[t]).max())


def test_step_bound_bit_exact(host, port, state):
    F, UV, rest8, a = _args(state.F, state.UV, state.rest8)
    p = np.ascontiguousarray(state.r("searchDir")[:2 * state.nV])
    got = host.host_step_bound(state.nV, state.nF, F.ctypes.data_as(_i), UV.ctypes.data_as(_d), p.ctypes.data_as(_d), C.c_double(1.0))
    assert got == port.init_step_size(F, UV, p, 1.0)
