"""TEST INFRASTRUCTURE — ctypes binding of oracle/_ref/liboptcuts_ref.so (the real reference).

Every function forwards to the unmodified reference class named in oracle/ref_capi.cpp.
Arrays cross as column-major (Fortran order) exactly like Eigen stores them.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboptcuts_ref.so")
_lib = None

_d = C.POINTER(C.c_double)
_i = C.POINTER(C.c_int32)
_l = C.POINTER(C.c_long)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_mesh_create.restype = C.c_void_p
        L.ref_mesh_create.argtypes = [C.c_int, C.c_int, _d, _i, _d, C.c_int, _i, C.c_double]
        for name in ("ref_mesh_destroy", "ref_solver_destroy", "ref_opt_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.ref_mesh_set_uv.argtypes = [C.c_void_p, _d]
        L.ref_mesh_get_uv.argtypes = [C.c_void_p, _d]
        L.ref_mesh_nV.argtypes = [C.c_void_p]
        L.ref_mesh_nF.argtypes = [C.c_void_p]
        L.ref_mesh_features.argtypes = [C.c_void_p, _d, _d]
        L.ref_mesh_coh_features.argtypes = [C.c_void_p, _i, _d]
        L.ref_mesh_adjacency.argtypes = [C.c_void_p, _i, _i]
        L.ref_mesh_adjacency.restype = C.c_long
        L.ref_mesh_check_inversion.argtypes = [C.c_void_p]
        L.ref_mesh_seam_sparsity.argtypes = [C.c_void_p, C.c_int]
        L.ref_mesh_seam_sparsity.restype = C.c_double
        L.ref_sd_energy.argtypes = [C.c_void_p, C.c_int]
        L.ref_sd_energy.restype = C.c_double
        L.ref_sd_energy_per_elem.argtypes = [C.c_void_p, C.c_int, _d]
        L.ref_sd_gradient.argtypes = [C.c_void_p, C.c_int, _d]
        L.ref_sd_hessian_triplets.argtypes = [C.c_void_p, C.c_int, _d, _i, _i]
        L.ref_sd_hessian_triplets.restype = C.c_long
        L.ref_sd_hessian_dense.argtypes = [C.c_void_p, C.c_int, _d]
        L.ref_sd_init_step_size.argtypes = [C.c_void_p, _d, C.c_double]
        L.ref_sd_init_step_size.restype = C.c_double
        L.ref_sd_divgrad.argtypes = [C.c_void_p, _d]
        L.ref_make_pd6.argtypes = [_d]
        L.ref_solver_create.restype = C.c_void_p
        L.ref_solver_set_pattern.argtypes = [C.c_void_p, C.c_int, _i, _i, C.c_int, _i]
        L.ref_solver_update_a.argtypes = [C.c_void_p, C.c_long, _i, _i, _d]
        L.ref_solver_num_rows.argtypes = [C.c_void_p]
        L.ref_solver_nnz.argtypes = [C.c_void_p]
        L.ref_solver_nnz.restype = C.c_long
        L.ref_solver_get_csr.argtypes = [C.c_void_p, _i, _i, _d]
        L.ref_solver_factorize.argtypes = [C.c_void_p]
        L.ref_solver_solve.argtypes = [C.c_void_p, _d, _d]
        L.ref_opt_create.restype = C.c_void_p
        L.ref_opt_create.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.ref_opt_solve.argtypes = [C.c_void_p, C.c_int]
        L.ref_opt_set_energy_param.argtypes = [C.c_void_p, C.c_double]
        L.ref_opt_scalars.argtypes = [C.c_void_p, _d]
        L.ref_opt_sizes.argtypes = [C.c_void_p, _l]
        L.ref_opt_get_uv.argtypes = [C.c_void_p, _d]
        L.ref_opt_get_air.argtypes = [C.c_void_p, _d, _i, _i, _d, _d, _i]
        L.ref_opt_get_gradient.argtypes = [C.c_void_p, _d]
        L.ref_opt_get_search_dir.argtypes = [C.c_void_p, _d]
        L.ref_opt_get_csr.argtypes = [C.c_void_p, _i, _i, _d]
        L.ref_opt_get_triplets.argtypes = [C.c_void_p, _i, _i, _d]
        L.ref_opt_recompute_gradient.argtypes = [C.c_void_p]
        L.ref_opt_recompute_energy.argtypes = [C.c_void_p]
        L.ref_opt_recompute_energy.restype = C.c_double
        L.ref_scaffold_create.restype = C.c_void_p
        L.ref_scaffold_create.argtypes = [C.c_void_p]
        L.ref_scaffold_destroy.argtypes = [C.c_void_p]
        L.ref_scaffold_destroy.restype = None
        L.ref_scaffold_sizes.argtypes = [C.c_void_p, _l]
        L.ref_scaffold_get.argtypes = [C.c_void_p, _d, _i, _i, _d, _d, _i, _d]
        L.ref_mesh_reset_fixed.argtypes = [C.c_void_p, C.c_int, _i]
        L.ref_local_solve.argtypes = [C.c_void_p, C.c_double, C.c_int, _d, _d]
        L.ref_timers_get.argtypes = [_d, _d]
        L.ref_set_output_folder.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def _pd(a):
    return a.ctypes.data_as(_d)


def _pi(a):
    return a.ctypes.data_as(_i)


def _f64(a):
    return np.asfortranarray(a, dtype=np.float64)


def _i32(a):
    return np.asfortranarray(a, dtype=np.int32)


class RefMesh:
    """OptCuts::TriMesh built with separateTri=false (TriMesh.cpp:41-199); vertex 0 is fixed."""

    def __init__(self, V_rest, F, UV, cohE=None, areaThres_AM=0.0):
        L = lib()
        self.V_rest, self.F, UV = _f64(V_rest), _i32(F), _f64(UV)
        self.nV, self.nF = self.V_rest.shape[0], self.F.shape[0]
        coh = _i32(cohE) if cohE is not None and len(cohE) else np.zeros((0, 4), np.int32, order="F")
        self.nCoh = coh.shape[0]
        self.h = L.ref_mesh_create(self.nV, self.nF, _pd(self.V_rest), _pi(self.F), _pd(UV),
                                   self.nCoh, _pi(coh), float(areaThres_AM))

    def close(self):
        if self.h:
            lib().ref_mesh_destroy(self.h)
            self.h = None

    def set_uv(self, UV):
        UV = _f64(UV)
        lib().ref_mesh_set_uv(self.h, _pd(UV))

    def get_uv(self):
        out = np.zeros((self.nV, 2), order="F")
        lib().ref_mesh_get_uv(self.h, _pd(out))
        return out

    def features(self):
        rest8 = np.zeros((8, self.nF))
        sc = np.zeros(3)
        lib().ref_mesh_features(self.h, _pd(rest8), _pd(sc))
        return rest8, dict(surfaceArea=sc[0], avgEdgeLen=sc[1], virtualRadius=sc[2])

    def coh_features(self):
        b = np.zeros(self.nCoh, np.int32)
        e = np.zeros(self.nCoh)
        lib().ref_mesh_coh_features(self.h, _pi(b), _pd(e))
        return b, e

    def adjacency(self):
        L = lib()
        tot = L.ref_mesh_adjacency(self.h, None, None)
        ptr = np.zeros(self.nV + 1, np.int32)
        idx = np.zeros(tot, np.int32)
        L.ref_mesh_adjacency(self.h, _pi(ptr), _pi(idx))
        return ptr, idx

    def check_inversion(self):
        return bool(lib().ref_mesh_check_inversion(self.h))

    def seam_sparsity(self, triSoup=False):
        return lib().ref_mesh_seam_sparsity(self.h, int(triSoup))

    # --- SymDirichletEnergy on this mesh
    def energy(self, uniform=False):
        return lib().ref_sd_energy(self.h, int(uniform))

    def energy_per_elem(self, uniform=False):
        out = np.zeros(self.nF)
        lib().ref_sd_energy_per_elem(self.h, int(uniform), _pd(out))
        return out

    def gradient(self, uniform=False):
        out = np.zeros(2 * self.nV)
        lib().ref_sd_gradient(self.h, int(uniform), _pd(out))
        return out

    def hessian_triplets(self, uniform=False):
        L = lib()
        n = L.ref_sd_hessian_triplets(self.h, int(uniform), None, None, None)
        V, I, J = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
        L.ref_sd_hessian_triplets(self.h, int(uniform), _pd(V), _pi(I), _pi(J))
        return I, J, V

    def hessian_dense(self, uniform=False):
        n = 2 * self.nV
        H = np.zeros((n, n), order="F")
        lib().ref_sd_hessian_dense(self.h, int(uniform), _pd(H))
        return H

    def init_step_size(self, searchDir, stepSize0=1.0):
        p = np.ascontiguousarray(searchDir, dtype=np.float64)
        return lib().ref_sd_init_step_size(self.h, _pd(p), float(stepSize0))

    def divgrad(self):
        out = np.zeros(self.nV)
        lib().ref_sd_divgrad(self.h, _pd(out))
        return out


def local_solve(V_rest, F, UV, is_free, relGL2Tol=1e-6, maxIter=100):
    """The nested dense Optimizer of TriMesh::computeLocalEdDec_* on one local stencil (no air mesh)."""
    m = RefMesh(V_rest, F, UV)
    fixed = _i32(np.nonzero(~np.asarray(is_free, bool))[0])
    lib().ref_mesh_reset_fixed(m.h, len(fixed), _pi(fixed))
    out = np.zeros(3)
    UVo = np.zeros((m.nV, 2), order="F")
    lib().ref_local_solve(m.h, float(relGL2Tol), int(maxIter), _pd(out), _pd(UVo))
    m.close()
    return dict(E_init=out[0], E_final=out[1], iters=int(out[2]), UV=UVo)


def build_scaffold(mesh):
    """OptCuts::Scaffold(mesh) (Scaffold.cpp:27-208): returns the air mesh as plain arrays."""
    L = lib()
    h = L.ref_scaffold_create(mesh.h)
    sz = np.zeros(5, dtype=np.int64)
    L.ref_scaffold_sizes(h, sz.ctypes.data_as(_l))
    nVa, nFa, nB, nFx = int(sz[0]), int(sz[1]), int(sz[2]), int(sz[3])
    Va = np.zeros((nVa, 2), order="F")
    Fa = np.zeros((nFa, 3), np.int32, order="F")
    bnd = np.zeros(nB, np.int32)
    rest8 = np.zeros((8, nFa))
    sc = np.zeros(3)
    fx = np.zeros(nFx, np.int32)
    thr = C.c_double()
    L.ref_scaffold_get(h, _pd(Va), _pi(Fa), _pi(bnd), _pd(rest8), _pd(sc), _pi(fx), C.byref(thr))
    L.ref_scaffold_destroy(h)
    return dict(V=Va, F=Fa, bnd=bnd, rest8=rest8, fixed=fx, areaThres_AM=thr.value, wholeMeshSize=int(sz[4]),
                surfaceArea=sc[0], avgEdgeLen=sc[1])


def make_pd6(M):
    A = np.array(M, dtype=np.float64, order="F").copy(order="F")
    lib().ref_make_pd6(_pd(A))
    return A


class RefSolver:
    """OptCuts::EigenLibSolver through the LinSysSolver surface (LinSysSolver.hpp:37-170)."""

    def __init__(self):
        self.h = lib().ref_solver_create()

    def close(self):
        if self.h:
            lib().ref_solver_destroy(self.h)
            self.h = None

    def set_pattern(self, adjPtr, adjIdx, fixed):
        adjPtr, adjIdx, fixed = _i32(adjPtr), _i32(adjIdx), _i32(fixed)
        lib().ref_solver_set_pattern(self.h, len(adjPtr) - 1, _pi(adjPtr), _pi(adjIdx), len(fixed), _pi(fixed))

    def update_a(self, I, J, S):
        I, J, S = _i32(I), _i32(J), _f64(S)
        lib().ref_solver_update_a(self.h, len(S), _pi(I), _pi(J), _pd(S))

    def csr(self):
        L = lib()
        n, nnz = L.ref_solver_num_rows(self.h), L.ref_solver_nnz(self.h)
        ia, ja, a = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        L.ref_solver_get_csr(self.h, _pi(ia), _pi(ja), _pd(a))
        return ia, ja, a

    def factorize(self):
        return bool(lib().ref_solver_factorize(self.h))

    def solve(self, rhs):
        rhs = _f64(rhs)
        x = np.zeros_like(rhs)
        lib().ref_solver_solve(self.h, _pd(rhs), _pd(x))
        return x


class RefOptimizer:
    """OptCuts::Optimizer (sparse path, propagateFracture=0) on a RefMesh; precompute() is run."""

    def __init__(self, mesh, energyParam0, scaffolding=True, mute=True):
        self.mesh = mesh
        self.h = lib().ref_opt_create(mesh.h, float(energyParam0), int(scaffolding), int(mute))
        self.scaffolding = scaffolding

    def close(self):
        if self.h:
            lib().ref_opt_destroy(self.h)
            self.h = None

    def solve(self, maxIter=1):
        return lib().ref_opt_solve(self.h, int(maxIter))

    def scalars(self):
        s = np.zeros(8)
        lib().ref_opt_scalars(self.h, _pd(s))
        keys = ("lastEnergyVal", "energyVal_scaffold", "energyVal_ET0", "lastEDec", "targetGRes", "w_scaf", "sqn_g", "iterNum")
        return dict(zip(keys, s))

    def sizes(self):
        s = np.zeros(9, dtype=np.int64)
        lib().ref_opt_sizes(self.h, s.ctypes.data_as(_l))
        keys = ("nV", "nF", "nVa", "nFa", "nBnd", "nSys", "nnz", "nTriplets", "nFixedAir")
        return dict(zip(keys, (int(v) for v in s)))

    def uv(self):
        out = np.zeros((self.sizes()["nV"], 2), order="F")
        lib().ref_opt_get_uv(self.h, _pd(out))
        return out

    def air(self):
        sz = self.sizes()
        Va = np.zeros((sz["nVa"], 2), order="F")
        Fa = np.zeros((sz["nFa"], 3), np.int32, order="F")
        l2g = np.zeros(sz["nVa"], np.int32)
        rest8 = np.zeros((8, sz["nFa"]))
        sc = np.zeros(3)
        fx = np.zeros(sz["nFixedAir"], np.int32)
        lib().ref_opt_get_air(self.h, _pd(Va), _pi(Fa), _pi(l2g), _pd(rest8), _pd(sc), _pi(fx))
        return dict(V=Va, F=Fa, localVI2Global=l2g, rest8=rest8, nBnd=sz["nBnd"], fixed=fx,
                    surfaceArea=sc[0], avgEdgeLen=sc[1])

    def gradient(self):
        out = np.zeros(self.sizes()["nSys"])
        lib().ref_opt_get_gradient(self.h, _pd(out))
        return out

    def search_dir(self):
        out = np.zeros(self.sizes()["nSys"])
        lib().ref_opt_get_search_dir(self.h, _pd(out))
        return out

    def csr(self):
        sz = self.sizes()
        ia, ja, a = np.zeros(sz["nSys"] + 1, np.int32), np.zeros(sz["nnz"], np.int32), np.zeros(sz["nnz"])
        lib().ref_opt_get_csr(self.h, _pi(ia), _pi(ja), _pd(a))
        return ia, ja, a

    def triplets(self):
        n = self.sizes()["nTriplets"]
        I, J, V = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
        lib().ref_opt_get_triplets(self.h, _pi(I), _pi(J), _pd(V))
        return I, J, V

    def recompute_gradient(self):
        lib().ref_opt_recompute_gradient(self.h)
        return self.gradient()

    def recompute_energy(self):
        return lib().ref_opt_recompute_energy(self.h)


def timers_reset():
    lib().ref_timers_reset()


def timers():
    t4, s9 = np.zeros(4), np.zeros(9)
    lib().ref_timers_get(_pd(t4), _pd(s9))
    names4 = ("topology", "descent", "scaffolding", "energyUpdate")
    names9 = ("mtrComp", "mtrAssem", "symFac", "numFac", "backSolve", "lineSearch", "bSplit", "iSplit", "cMerge")
    return dict(zip(names4, t4)), dict(zip(names9, s9))


def set_output_folder(path):
    lib().ref_set_output_folder(path.encode())
