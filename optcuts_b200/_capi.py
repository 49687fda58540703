"""ctypes binding of liboptcuts_b200.so — one Python method per C-ABI entry point.

Arrays cross the boundary as host numpy buffers in the layouts the header states (Eigen
column-major matrices, interleaved system vectors).  No torch types in any signature.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_d = C.POINTER(C.c_double)
_i = C.POINTER(C.c_int32)
_u8 = C.POINTER(C.c_uint8)
_i64 = C.POINTER(C.c_int64)

STATUS = {0: "OCB_OK", -1: "OCB_ERR_CUDA", -2: "OCB_ERR_ARG", -3: "OCB_ERR_STATE", -4: "OCB_ERR_INVERTED",
          -5: "OCB_ERR_NOT_CONVERGED", -6: "OCB_ERR_BREAKDOWN"}
OCB_ERR_INVERTED = -4
OCB_ERR_NOT_CONVERGED = -5
OCB_ERR_BREAKDOWN = -6


class OcbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (STATUS.get(code, "?"), code, msg))
        self.code = code


class LineSearchResult(C.Structure):
    _fields_ = [("alpha", C.c_double), ("E_new", C.c_double), ("E_scaf_new", C.c_double), ("E_sd_new", C.c_double),
                ("E_last", C.c_double), ("lastEDec", C.c_double), ("n_halvings", C.c_int), ("stopped", C.c_int)]


class NewtonResult(C.Structure):
    _fields_ = [("sqn_g", C.c_double), ("targetGRes", C.c_double), ("alpha", C.c_double), ("E_new", C.c_double),
                ("E_scaf_new", C.c_double), ("E_sd_new", C.c_double), ("lastEDec", C.c_double),
                ("pcg_rel_res", C.c_double), ("converged", C.c_int), ("stopped", C.c_int),
                ("n_halvings", C.c_int), ("pcg_iters", C.c_int),
                ("alpha_init", C.c_double), ("E_last", C.c_double), ("ms_solve", C.c_double), ("ms_line_search", C.c_double),
                ("pcg_status", C.c_int), ("reserved", C.c_int)]


class StencilStepBatch(C.Structure):
    _fields_ = [("nStencil", C.c_int), ("vert_ptr", _i), ("tri_ptr", _i), ("n_mesh_vert", _i), ("n_mesh_tri", _i), ("V_rest", _d), ("UV", _d),
                ("F", _i), ("is_free", _u8), ("area_thres", _d), ("target_gres", _d), ("w_scaf", C.c_double)]


class StencilBatch(C.Structure):
    _fields_ = [("nStencil", C.c_int), ("vert_ptr", _i), ("tri_ptr", _i), ("V_rest", _d), ("UV", _d),
                ("F", _i), ("is_free", _u8), ("score_scale", _d), ("score_offset", _d)]


def lib_path():
    return os.environ.get("OPTCUTS_B200_LIB", os.path.join(_HERE, "lib", "liboptcuts_b200.so"))


# every symbol include/optcuts_b200.h declares: (name, restype, argtypes)
_SIGS = [
    ("ocb_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    ("ocb_destroy", None, [C.c_void_p]),
    ("ocb_last_error", C.c_char_p, [C.c_void_p]),
    ("ocb_version", C.c_char_p, []),
    ("ocb_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("ocb_use_own_stream", C.c_int, [C.c_void_p]),
    ("ocb_synchronize", C.c_int, [C.c_void_p]),
    ("ocb_set_option", C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    ("ocb_timer_start", C.c_int, [C.c_void_p]),
    ("ocb_timer_stop_ms", C.c_int, [C.c_void_p, _d]),
    ("ocb_launch_count", C.c_int64, [C.c_void_p]),
    ("ocb_profile_enable", C.c_int, [C.c_void_p, C.c_int]),
    ("ocb_profile_get", C.c_int, [C.c_void_p, _d, _i64]),
    ("ocb_profile_name", C.c_char_p, [C.c_int]),
    ("ocb_profile_classes", C.c_int, []),
    ("ocb_save_uv", C.c_int, [C.c_void_p]),
    ("ocb_restore_uv", C.c_int, [C.c_void_p]),
    ("ocb_rest_features", C.c_int, [C.c_void_p, C.c_int, C.c_int, _d, _i, C.c_double, _d, _d]),
    ("ocb_set_mesh", C.c_int, [C.c_void_p, C.c_int, C.c_int, _i, _d, C.c_double, _i, C.c_int]),
    ("ocb_set_air", C.c_int, [C.c_void_p, C.c_int, C.c_int, _i, _d, _i, C.c_int, _i, C.c_int, C.c_double]),
    ("ocb_set_uv", C.c_int, [C.c_void_p, _d, _d]),
    ("ocb_get_uv", C.c_int, [C.c_void_p, _d, _d]),
    ("ocb_get_sizes", C.c_int, [C.c_void_p, _i64]),
    ("ocb_energy", C.c_int, [C.c_void_p, C.c_double, _d, _d, _d]),
    ("ocb_energy_per_elem", C.c_int, [C.c_void_p, C.c_int, _d]),
    ("ocb_energy_by_elem", C.c_int, [C.c_void_p, C.c_int, C.c_int, _d]),
    ("ocb_hessian_dense", C.c_int, [C.c_void_p, C.c_int, _d]),
    ("ocb_gradient", C.c_int, [C.c_void_p, C.c_double, _d, _d]),
    ("ocb_set_pattern", C.c_int, [C.c_void_p, C.c_int, _i, _i, _i, C.c_int]),
    ("ocb_set_pattern_from_elements", C.c_int, [C.c_void_p]),
    ("ocb_hessian_assemble", C.c_int, [C.c_void_p, C.c_double]),
    ("ocb_hessian_triplets", C.c_int, [C.c_void_p, C.c_int, _d, _i, _i, _i64]),
    ("ocb_hessian_blocks", C.c_int, [C.c_void_p, C.c_int, _d]),
    ("ocb_update_values_triplets", C.c_int, [C.c_void_p, C.c_int64, _i, _i, _d]),
    ("ocb_download_csr", C.c_int, [C.c_void_p, _i, _i, _d]),
    ("ocb_matrix_version", C.c_longlong, [C.c_void_p]),
    ("ocb_multiply", C.c_int, [C.c_void_p, _d, _d]),
    ("ocb_factorize", C.c_int, [C.c_void_p]),
    ("ocb_solve", C.c_int, [C.c_void_p, _d, _d, C.c_double, C.c_int, C.POINTER(C.c_int), _d]),
    ("ocb_direct_level_blocks", C.c_int, [C.c_int, _i, _i, C.c_int, _i, _i, _i]),
    ("ocb_get_search_dir", C.c_int, [C.c_void_p, _d]),
    ("ocb_set_search_dir", C.c_int, [C.c_void_p, _d]),
    ("ocb_step_bound", C.c_int, [C.c_void_p, _d, _d]),
    ("ocb_line_search", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(LineSearchResult)]),
    ("ocb_step_forward", C.c_int, [C.c_void_p, C.c_double]),
    ("ocb_newton_step", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(NewtonResult)]),
    ("ocb_newton_step_ex", C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(NewtonResult)]),
    ("ocb_gradient_info", C.c_int, [C.c_void_p, _d]),
    ("ocb_seam_energy", C.c_int, [C.c_void_p, C.c_int, _i, _d, _i, C.c_double, C.c_double, C.c_double, C.c_int, _d]),
    ("ocb_divgrad_scores", C.c_int, [C.c_void_p, _d]),
    ("ocb_eval_stencils", C.c_int, [C.c_void_p, C.POINTER(StencilBatch), C.c_int, C.c_double, _d, _d, _d, _i, _d, _i, C.POINTER(C.c_int)]),
    ("ocb_stencil_newton_step", C.c_int, [C.c_void_p, C.POINTER(StencilStepBatch), _d, _d, _i]),
    ("ocb_precond_info", C.c_int, [C.c_void_p, _i]),
    ("ocb_set_coordinate_hint", C.c_int, [C.c_void_p, C.c_int, _d]),
    ("ocb_precond_hierarchy", C.c_int, [C.c_void_p, C.c_int, _d, C.c_int, _i, _i, _i, C.c_int]),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGS]


def precond_hierarchy(xy, grid):
    """Host-only: row order and multilevel hierarchy the solver would build for points xy (n x 2) and `grid` CTAs.
    Returns (vert_of, [child_beg per level], local_levels)."""
    L = load_library()
    xy = _f64(np.asarray(xy, dtype=np.float64), order="C")
    n = xy.shape[0]
    h = C.c_void_p()
    if L.ocb_create(C.byref(h), 0) != 0:
        raise OcbError(-1, "ocb_create failed")
    try:
        vert_of, info = np.zeros(n, np.int32), np.zeros(16, np.int32)
        cap = 2 * n + 64 * 16
        cb = np.zeros(cap, np.int32)
        w = L.ocb_precond_hierarchy(h, n, _pd(xy), int(grid), _pi(vert_of), _pi(info), _pi(cb), cap)
        if w < 0:
            raise OcbError(w, L.ocb_last_error(h).decode())
        levels, o = [], 0
        for l in range(int(info[1])):
            m = int(info[4 + l]) + 1
            levels.append(cb[o:o + m].copy())
            o += m
        return vert_of, levels, int(info[2])
    finally:
        L.ocb_destroy(h)


def direct_level_blocks(row_ptr, col_idx, target=192):
    """Host-only: level order and blocks of the direct safety net for a symmetric vertex pattern (CSR).  Returns (pos, blk_of, blk_beg)."""
    L = load_library()
    row_ptr, col_idx = _i32(np.asarray(row_ptr), "C"), _i32(np.asarray(col_idx), "C")
    n = len(row_ptr) - 1
    pos, blk_of, blk_beg = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n + 1, np.int32)
    nb = L.ocb_direct_level_blocks(n, _pi(row_ptr), _pi(col_idx), int(target), _pi(pos), _pi(blk_of), _pi(blk_beg))
    if nb < 0:
        raise OcbError(nb, "ocb_direct_level_blocks")
    return pos, blk_of, blk_beg[:nb + 1].copy()


def load_library():
    """Load the CUDA library or fail loudly — there is no fallback implementation."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise OcbError(-3, "CUDA library %s is missing: run `python -m optcuts_b200.build` (nvcc, sm_100a). "
                               "optcuts_b200 has no CPU fallback." % path)
        L = C.CDLL(path)
        for name, res, args in _SIGS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def _f64(a, order="F"):
    return np.require(a, dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def _i32(a, order="F"):
    return np.require(a, dtype=np.int32, requirements=["F" if order == "F" else "C", "A"])


def _pd(a):
    return None if a is None else a.ctypes.data_as(_d)


def _pi(a):
    return None if a is None else a.ctypes.data_as(_i)


class Context:
    """One ocb_ctx: device buffers, stream, matrix.  Methods are the C entry points, 1:1."""

    def __init__(self, device=0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.ocb_create(C.byref(h), int(device))
        if rc != 0:
            raise OcbError(rc, "ocb_create failed")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.ocb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise OcbError(rc, self._L.ocb_last_error(self._h).decode())
        return rc

    # -- plumbing
    def version(self):
        return self._L.ocb_version().decode()

    def set_stream(self, cuda_stream_ptr):
        self._chk(self._L.ocb_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def use_own_stream(self):
        self._chk(self._L.ocb_use_own_stream(self._h))

    def set_option(self, key, value):
        self._chk(self._L.ocb_set_option(self._h, key.encode(), float(value)))

    def synchronize(self):
        self._chk(self._L.ocb_synchronize(self._h))

    def timer_start(self):
        self._chk(self._L.ocb_timer_start(self._h))

    def timer_stop_ms(self):
        ms = C.c_double()
        self._chk(self._L.ocb_timer_stop_ms(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self._L.ocb_launch_count(self._h))

    def profile_enable(self, on=True):
        self._chk(self._L.ocb_profile_enable(self._h, int(on)))

    def profile_get(self):
        n = self._L.ocb_profile_classes()
        ms, cnt = np.zeros(n), np.zeros(n, np.int64)
        self._chk(self._L.ocb_profile_get(self._h, _pd(ms), cnt.ctypes.data_as(_i64)))
        return {self._L.ocb_profile_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}

    def save_uv(self):
        self._chk(self._L.ocb_save_uv(self._h))

    def restore_uv(self):
        self._chk(self._L.ocb_restore_uv(self._h))

    def sizes(self):
        s = np.zeros(8, np.int64)
        self._chk(self._L.ocb_get_sizes(self._h, s.ctypes.data_as(_i64)))
        return dict(zip(("nV", "nF", "nVa", "nFa", "nBnd", "nSys", "nnz_upper", "nnz_blocks"), (int(v) for v in s)))

    # -- a1
    def rest_features(self, V_rest, F, areaThres_AM=0.0):
        V_rest, F = _f64(V_rest), _i32(F)
        nV, nF = V_rest.shape[0], F.shape[0]
        rest8 = np.zeros((8, nF))
        sc = np.zeros(3)
        self._chk(self._L.ocb_rest_features(self._h, nV, nF, _pd(V_rest), _pi(F), float(areaThres_AM), _pd(rest8), _pd(sc)))
        return rest8, dict(surfaceArea=sc[0], avgEdgeLen=sc[1], virtualRadius=sc[2])

    # -- problem data
    def set_mesh(self, nV, F, rest8, surfaceArea, fixed):
        F, rest8, fixed = _i32(F), _f64(rest8, "C"), _i32(np.asarray(fixed, dtype=np.int32).ravel())
        self._chk(self._L.ocb_set_mesh(self._h, int(nV), F.shape[0], _pi(F), _pd(rest8), float(surfaceArea), _pi(fixed), len(fixed)))

    def set_air(self, Fa, rest8, localVI2Global, nBnd, fixedAir, w_scaf_over_Fa):
        if Fa is None or len(Fa) == 0:
            self._chk(self._L.ocb_set_air(self._h, 0, 0, None, None, None, 0, None, 0, 0.0))
            return
        Fa, rest8, l2g = _i32(Fa), _f64(rest8, "C"), _i32(np.asarray(localVI2Global).ravel())
        fx = _i32(np.asarray(fixedAir, dtype=np.int32).ravel())
        self._chk(self._L.ocb_set_air(self._h, len(l2g), Fa.shape[0], _pi(Fa), _pd(rest8), _pi(l2g), int(nBnd), _pi(fx), len(fx), float(w_scaf_over_Fa)))

    def set_uv(self, V=None, Va=None):
        V = None if V is None else _f64(V)
        Va = None if Va is None else _f64(Va)
        self._chk(self._L.ocb_set_uv(self._h, _pd(V), _pd(Va)))

    def get_uv(self, want_air=False):
        s = self.sizes()
        V = np.zeros((s["nV"], 2), order="F")
        Va = np.zeros((s["nVa"], 2), order="F") if (want_air and s["nVa"] > 0) else None
        self._chk(self._L.ocb_get_uv(self._h, _pd(V), _pd(Va)))
        return (V, Va) if want_air else V

    # -- energy / gradient
    def energy(self, energyParam0=1.0, check_inversion=True):
        et, es, ec = C.c_double(), C.c_double(), C.c_double()
        self._chk(self._L.ocb_energy(self._h, float(energyParam0), C.byref(et), C.byref(es), C.byref(ec)),
                  allow=() if check_inversion else (OCB_ERR_INVERTED,))
        return et.value, es.value, ec.value

    def energy_per_elem(self, uniformWeight=False):
        out = np.zeros(self.sizes()["nF"])
        self._chk(self._L.ocb_energy_per_elem(self._h, int(uniformWeight), _pd(out)))
        return out

    def energy_by_elem(self, triI, uniformWeight=False):
        e = C.c_double()
        self._chk(self._L.ocb_energy_by_elem(self._h, int(triI), int(uniformWeight), C.byref(e)))
        return e.value

    def hessian_dense(self, uniformWeight=False):
        n = 2 * self.sizes()["nV"]
        H = np.zeros((n, n))
        self._chk(self._L.ocb_hessian_dense(self._h, int(uniformWeight), _pd(H)))
        return H

    def gradient(self, energyParam0=1.0, download=True):
        g = np.zeros(self.sizes()["nSys"]) if download else None
        sq = C.c_double()
        self._chk(self._L.ocb_gradient(self._h, float(energyParam0), _pd(g), C.byref(sq)))
        return g, sq.value

    # -- pattern / matrix
    def set_pattern(self, adj_ptr, adj_idx, fixed):
        adj_ptr, adj_idx = _i32(np.asarray(adj_ptr).ravel()), _i32(np.asarray(adj_idx).ravel())
        fixed = _i32(np.asarray(fixed, dtype=np.int32).ravel())
        self._chk(self._L.ocb_set_pattern(self._h, len(adj_ptr) - 1, _pi(adj_ptr), _pi(adj_idx), _pi(fixed), len(fixed)))

    def set_pattern_from_elements(self):
        self._chk(self._L.ocb_set_pattern_from_elements(self._h))

    def hessian_assemble(self, energyParam0=1.0):
        self._chk(self._L.ocb_hessian_assemble(self._h, float(energyParam0)))

    def hessian_blocks(self, uniformWeight=False):
        out = np.zeros((self.sizes()["nF"], 6, 6))
        self._chk(self._L.ocb_hessian_blocks(self._h, int(uniformWeight), _pd(out)))
        return out

    def hessian_triplets(self, uniformWeight=False):
        n = C.c_int64()
        self._chk(self._L.ocb_hessian_triplets(self._h, int(uniformWeight), None, None, None, C.byref(n)))
        V, I, J = np.zeros(n.value), np.zeros(n.value, np.int32), np.zeros(n.value, np.int32)
        self._chk(self._L.ocb_hessian_triplets(self._h, int(uniformWeight), _pd(V), _pi(I), _pi(J), C.byref(n)))
        return I, J, V

    def update_values_triplets(self, I, J, S):
        I, J, S = _i32(np.asarray(I).ravel()), _i32(np.asarray(J).ravel()), _f64(np.asarray(S).ravel())
        self._chk(self._L.ocb_update_values_triplets(self._h, len(S), _pi(I), _pi(J), _pd(S)))

    def download_csr(self):
        s = self.sizes()
        ia, ja, a = np.zeros(s["nSys"] + 1, np.int32), np.zeros(s["nnz_upper"], np.int32), np.zeros(s["nnz_upper"])
        self._chk(self._L.ocb_download_csr(self._h, _pi(ia), _pi(ja), _pd(a)))
        return ia, ja, a

    def multiply(self, x):
        x = _f64(np.asarray(x).ravel())
        y = np.zeros_like(x)
        self._chk(self._L.ocb_multiply(self._h, _pd(x), _pd(y)))
        return y

    # -- solve
    def factorize(self):
        self._chk(self._L.ocb_factorize(self._h))

    def set_coordinate_hint(self, xy):
        """Positions (n x 2) of the first n vertices of the next set_pattern of a solver-only context."""
        xy = _f64(np.asarray(xy, dtype=np.float64).reshape(-1, 2))
        self._chk(self._L.ocb_set_coordinate_hint(self._h, xy.shape[0], _pd(xy)))

    def precond_info(self):
        """The solver's multilevel preconditioner hierarchy (built with the pattern)."""
        info = np.zeros(16, np.int32)
        self._chk(self._L.ocb_precond_info(self._h, _pi(info)))
        L = int(info[1])
        return dict(enabled=bool(info[0]), levels=L, local_levels=int(info[2]), grid=int(info[3]), nodes=[int(v) for v in info[4:4 + L]],
                    fallbacks=int(info[15]), direct_solves=int(info[14]))

    def solve(self, rhs=None, rel_tol=1e-12, max_it=0, download=True, allow_not_converged=False):
        rhs = None if rhs is None else _f64(np.asarray(rhs).ravel())
        x = np.zeros(self.sizes()["nSys"]) if download else None
        it, rr = C.c_int(), C.c_double()
        rc = self._chk(self._L.ocb_solve(self._h, _pd(rhs), _pd(x), float(rel_tol), int(max_it), C.byref(it), C.byref(rr)),
                       allow=(OCB_ERR_NOT_CONVERGED,) if allow_not_converged else ())
        return x, dict(iters=it.value, rel_res=rr.value, status=rc)

    def get_search_dir(self):
        p = np.zeros(self.sizes()["nSys"])
        self._chk(self._L.ocb_get_search_dir(self._h, _pd(p)))
        return p

    def set_search_dir(self, p):
        p = _f64(np.asarray(p).ravel())
        self._chk(self._L.ocb_set_search_dir(self._h, _pd(p)))

    # -- line search
    def step_bound(self, searchDir=None, alpha=1.0):
        d = None if searchDir is None else _f64(np.asarray(searchDir).ravel())
        a = C.c_double(alpha)
        self._chk(self._L.ocb_step_bound(self._h, _pd(d), C.byref(a)))
        return a.value

    def line_search(self, energyParam0, E_last, alpha0, allowEDecRelTol=True):
        r = LineSearchResult()
        self._chk(self._L.ocb_line_search(self._h, float(energyParam0), float(E_last), float(alpha0), int(allowEDecRelTol), C.byref(r)))
        return {k: getattr(r, k) for k, _ in LineSearchResult._fields_}

    def step_forward(self, alpha):
        self._chk(self._L.ocb_step_forward(self._h, float(alpha)))

    def newton_step(self, energyParam0, targetGRes, pcg_rel_tol=1e-12, pcg_max_it=0, allowEDecRelTol=True, flags=0):
        r = NewtonResult()
        self._chk(self._L.ocb_newton_step_ex(self._h, float(energyParam0), float(targetGRes), float(pcg_rel_tol),
                                             int(pcg_max_it), int(allowEDecRelTol), int(flags), C.byref(r)), allow=(OCB_ERR_NOT_CONVERGED,))
        return {k: getattr(r, k) for k, _ in NewtonResult._fields_}

    # -- seam / candidate filter
    def seam_energy(self, cohE, edgeLen, boundaryEdge, initSeamLen, virtualRadius, avgEdgeLen, triSoup=False):
        cohE = _i32(np.asarray(cohE).reshape(-1, 4))
        n = cohE.shape[0]
        e = C.c_double()
        self._chk(self._L.ocb_seam_energy(self._h, n, _pi(cohE), _pd(_f64(np.asarray(edgeLen).ravel())),
                                          _pi(_i32(np.asarray(boundaryEdge).ravel())), float(initSeamLen),
                                          float(virtualRadius), float(avgEdgeLen), int(triSoup), C.byref(e)))
        return e.value

    def eval_stencils(self, stencils, maxIter=100, relGL2Tol=1e-6, score_scale=None, score_offset=None):
        """stencils: list of (V_rest (nv,3), UV (nv,2), F (nt,3) local ids, is_free (nv,) bool).  Returns a dict of arrays."""
        nS = len(stencils)
        vp = np.zeros(nS + 1, np.int32); tp = np.zeros(nS + 1, np.int32)
        for k, (Vr, UV, F, fr) in enumerate(stencils):
            vp[k + 1] = vp[k] + len(Vr); tp[k + 1] = tp[k] + len(F)
        Vr = np.ascontiguousarray(np.concatenate([np.asarray(s[0], np.float64).reshape(-1, 3) for s in stencils]) if nS else np.zeros((0, 3)))
        UV = np.ascontiguousarray(np.concatenate([np.asarray(s[1], np.float64).reshape(-1, 2) for s in stencils]) if nS else np.zeros((0, 2)))
        F = np.ascontiguousarray(np.concatenate([np.asarray(s[2], np.int32).reshape(-1, 3) for s in stencils]) if nS else np.zeros((0, 3), np.int32))
        fr = np.ascontiguousarray(np.concatenate([np.asarray(s[3]).astype(np.uint8).ravel() for s in stencils]) if nS else np.zeros(0, np.uint8))
        sc = None if score_scale is None else _f64(np.asarray(score_scale, np.float64).ravel())
        so = None if score_offset is None else _f64(np.asarray(score_offset, np.float64).ravel())
        b = StencilBatch(nS, _pi(vp), _pi(tp), _pd(Vr), _pd(UV), _pi(F), fr.ctypes.data_as(_u8), _pd(sc), _pd(so))
        E0, E1, UVo = np.zeros(nS), np.zeros(nS), np.zeros_like(UV)
        it, st, score = np.zeros(nS, np.int32), np.zeros(nS, np.int32), np.zeros(nS)
        arg = C.c_int(-1)
        self._chk(self._L.ocb_eval_stencils(self._h, C.byref(b), int(maxIter), float(relGL2Tol), _pd(E0), _pd(E1), _pd(UVo), _pi(it),
                                            _pd(score), _pi(st), C.byref(arg)))
        return dict(E_init=E0, E_final=E1, UV=[UVo[vp[k]:vp[k + 1]] for k in range(nS)], iters=it, status=st, score=score, argmax=arg.value)

    def stencil_newton_step(self, stencils, w_scaf=0.01):
        """stencils: list of dicts {V_rest (nv,3), UV (nv,2), F (nt,3) local ids, is_free (nv,), n_mesh_vert, n_mesh_tri, area_thres,
        target_gres}: ONE Newton iteration of every nested (bijective) optimizer.  Returns dict(UV=[...], out6, result)."""
        nS = len(stencils)
        vp = np.zeros(nS + 1, np.int32); tp = np.zeros(nS + 1, np.int32)
        for k, s in enumerate(stencils):
            vp[k + 1] = vp[k] + len(s["UV"]); tp[k + 1] = tp[k] + len(s["F"])
        cat = lambda key, dt, w: np.ascontiguousarray(np.concatenate([np.asarray(s[key], dt).reshape(-1, w) for s in stencils]))
        Vr, UV, F = cat("V_rest", np.float64, 3), cat("UV", np.float64, 2), cat("F", np.int32, 3)
        fr = np.ascontiguousarray(np.concatenate([np.asarray(s["is_free"]).astype(np.uint8).ravel() for s in stencils]))
        nvm = np.array([s["n_mesh_vert"] for s in stencils], np.int32); ntm = np.array([s["n_mesh_tri"] for s in stencils], np.int32)
        th = np.array([s["area_thres"] for s in stencils], np.float64); tg = np.array([s["target_gres"] for s in stencils], np.float64)
        b = StencilStepBatch(nS, _pi(vp), _pi(tp), _pi(nvm), _pi(ntm), _pd(Vr), _pd(UV), _pi(F), fr.ctypes.data_as(_u8), _pd(th), _pd(tg), float(w_scaf))
        UVo, out6, res = np.zeros_like(UV), np.zeros((nS, 6)), np.zeros(nS, np.int32)
        self._chk(self._L.ocb_stencil_newton_step(self._h, C.byref(b), _pd(UVo), _pd(out6), _pi(res)))
        return dict(UV=[UVo[vp[k]:vp[k + 1]] for k in range(nS)], out6=out6, result=res)

    def divgrad_scores(self):
        out = np.zeros(self.sizes()["nV"])
        self._chk(self._L.ocb_divgrad_scores(self._h, _pd(out)))
        return out
