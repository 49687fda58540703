"""TEST INFRASTRUCTURE — ctypes binding of oracle/liboracle_port.so (our plain-C restatement)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_port.so")
_lib = None
_d = C.POINTER(C.c_double)
_i = C.POINTER(C.c_int32)


class NewtonResult(C.Structure):
    _fields_ = [("sqn_g", C.c_double), ("alpha", C.c_double), ("E_new", C.c_double), ("E_scaf_new", C.c_double),
                ("E_sd_new", C.c_double), ("lastEDec", C.c_double), ("E_last", C.c_double),
                ("converged", C.c_int), ("stopped", C.c_int), ("n_halvings", C.c_int)]


def build():
    src = os.path.join(_HERE, "port", "sd_port.c")
    if (not os.path.exists(LIB_PATH)) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB_PATH, src, "-lm"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.port_energy.restype = C.c_double
        L.port_init_step_size.restype = C.c_double
        L.port_seam_sparsity.restype = C.c_double
        L.port_hessian_triplets.restype = C.c_int64
        L.port_set_pattern.restype = C.c_int64
        _lib = L
    return _lib


def _pd(a):
    return None if a is None else a.ctypes.data_as(_d)


def _pi(a):
    return None if a is None else a.ctypes.data_as(_i)


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _n(a):
    return np.asfortranarray(a, dtype=np.int32)


def rest_features(V_rest, F, thres=0.0):
    V_rest, F = _f(V_rest), _n(F)
    if V_rest.shape[1] == 2:
        V_rest = _f(np.hstack([V_rest, np.zeros((V_rest.shape[0], 1))]))
    nV, nF = V_rest.shape[0], F.shape[0]
    rest8, sc = np.zeros((8, nF)), np.zeros(3)
    rc = lib().port_rest_features(nV, nF, _pd(V_rest), _pi(F), C.c_double(thres), _pd(rest8), _pd(sc))
    return rest8, dict(surfaceArea=sc[0], avgEdgeLen=sc[1], virtualRadius=sc[2]), rc


def energy_per_elem(F, UV, rest8, surf, uniform=False):
    F, UV, rest8 = _n(F), _f(UV), np.ascontiguousarray(rest8)
    out = np.zeros(F.shape[0])
    lib().port_energy_per_elem(UV.shape[0], F.shape[0], _pi(F), _pd(UV), _pd(rest8), C.c_double(surf), int(uniform), _pd(out))
    return out


def energy(F, UV, rest8, surf, uniform=False):
    F, UV, rest8 = _n(F), _f(UV), np.ascontiguousarray(rest8)
    return lib().port_energy(UV.shape[0], F.shape[0], _pi(F), _pd(UV), _pd(rest8), C.c_double(surf), int(uniform))


def gradient(F, UV, rest8, surf, uniform=False, fixed=(0,)):
    F, UV, rest8, fx = _n(F), _f(UV), np.ascontiguousarray(rest8), _n(np.asarray(fixed, np.int32))
    g = np.zeros(2 * UV.shape[0])
    lib().port_gradient(UV.shape[0], F.shape[0], _pi(F), _pd(UV), _pd(rest8), C.c_double(surf), int(uniform), _pi(fx), len(fx), _pd(g))
    return g


def make_pd6(M):
    A = np.ascontiguousarray(M, dtype=np.float64).copy()
    lib().port_make_pd6(_pd(A))
    return A


def hessian_blocks(F, UV, rest8, surf, uniform=False, project=True):
    F, UV, rest8 = _n(F), _f(UV), np.ascontiguousarray(rest8)
    out = np.zeros((F.shape[0], 6, 6))
    lib().port_hessian_blocks(UV.shape[0], F.shape[0], _pi(F), _pd(UV), _pd(rest8), C.c_double(surf), int(uniform), int(project), _pd(out))
    return out


def hessian_triplets(F, UV, rest8, surf, uniform=False, fixed=(0,)):
    F, UV, rest8, fx = _n(F), _f(UV), np.ascontiguousarray(rest8), _n(np.asarray(fixed, np.int32))
    args = (UV.shape[0], F.shape[0], _pi(F), _pd(UV), _pd(rest8), C.c_double(surf), int(uniform), _pi(fx), len(fx))
    n = lib().port_hessian_triplets(*args, None, None, None)
    V, I, J = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().port_hessian_triplets(*args, _pd(V), _pi(I), _pi(J))
    return I, J, V


def init_step_size(F, UV, searchDir, stepSize=1.0):
    F, UV, p = _n(F), _f(UV), np.ascontiguousarray(searchDir, dtype=np.float64)
    return lib().port_init_step_size(UV.shape[0], F.shape[0], _pi(F), _pd(UV), _pd(p), C.c_double(stepSize))


def set_pattern(adjPtr, adjIdx, fixed):
    adjPtr, adjIdx, fx = _n(adjPtr), _n(adjIdx), _n(np.asarray(fixed, np.int32))
    nV = len(adjPtr) - 1
    nnz = lib().port_set_pattern(nV, _pi(adjPtr), _pi(adjIdx), _pi(fx), len(fx), None, None)
    ia, ja = np.zeros(2 * nV + 1, np.int32), np.zeros(nnz, np.int32)
    lib().port_set_pattern(nV, _pi(adjPtr), _pi(adjIdx), _pi(fx), len(fx), _pi(ia), _pi(ja))
    return ia, ja


def update_a(ia, ja, I, J, S):
    ia, ja, I, J, S = _n(ia), _n(ja), _n(I), _n(J), _f(S)
    a = np.zeros(len(ja))
    miss = lib().port_update_a(len(ia) - 1, _pi(ia), _pi(ja), C.c_int64(len(S)), _pi(I), _pi(J), _pd(S), _pd(a))
    return a, miss


def ldlt_solve(ia, ja, a, rhs):
    ia, ja, a, rhs = _n(ia), _n(ja), _f(a), _f(rhs)
    x = np.zeros_like(rhs)
    rc = lib().port_ldlt_solve(len(ia) - 1, _pi(ia), _pi(ja), _pd(a), _pd(rhs), _pd(x))
    return x, rc


def seam_sparsity(cohE, boundaryEdge, edgeLen, UV, avgEdgeLen, initSeamLen=0.0, triSoup=False):
    cohE, b, e, UV = _n(np.asarray(cohE).reshape(-1, 4)), _n(boundaryEdge), _f(edgeLen), _f(UV)
    coh = np.where(cohE < 0, 0, cohE).astype(np.int32, order="F")
    return lib().port_seam_sparsity(UV.shape[0], coh.shape[0], _pi(coh), _pi(b), _pd(e), _pd(UV), C.c_double(avgEdgeLen),
                                    C.c_double(initSeamLen), int(triSoup))


def divgrad(F, UV, rest8, surf):
    F, UV, rest8 = _n(F), _f(UV), np.ascontiguousarray(rest8)
    out = np.zeros(UV.shape[0])
    lib().port_divgrad(UV.shape[0], F.shape[0], _pi(F), _pd(UV), _pd(rest8), C.c_double(surf), _pd(out))
    return out


def newton_step(F, UV, rest8, surf, fixed, energyParam0, targetGRes, air=None, w_scaf=0.0, allowEDecRelTol=True):
    """One Optimizer::solve(1) geometry step.  Returns (UV_new, UVa_new, searchDir, result dict)."""
    F, UV, rest8, fx = _n(F), _f(UV).copy(order="F"), np.ascontiguousarray(rest8), _n(np.asarray(fixed, np.int32))
    nV, nF = UV.shape[0], F.shape[0]
    if air is not None:
        Fa, UVa, r8a = _n(air["F"]), _f(air["V"]).copy(order="F"), np.ascontiguousarray(air["rest8"])
        l2g, fxa = _n(air["localVI2Global"]), _n(np.asarray(air["fixed"], np.int32))
        nVa, nFa, nB = UVa.shape[0], Fa.shape[0], int(air["nBnd"])
    else:
        Fa = UVa = r8a = l2g = None
        fxa = np.zeros(0, np.int32)
        nVa = nFa = nB = 0
    n = 2 * (nV + nVa - nB)
    p = np.zeros(n)
    r = NewtonResult()
    lib().port_newton_step(nV, nF, _pi(F), _pd(UV), _pd(rest8), C.c_double(surf), _pi(fx), len(fx),
                           nVa, nFa, _pi(Fa), _pd(UVa), _pd(r8a), _pi(l2g), nB, _pi(fxa), len(fxa),
                           C.c_double(energyParam0), C.c_double(w_scaf), C.c_double(targetGRes), int(allowEDecRelTol),
                           _pd(p), C.byref(r))
    return UV, UVa, p, {k: getattr(r, k) for k, _ in NewtonResult._fields_}
