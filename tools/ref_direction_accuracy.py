"""How accurate is the REFERENCE's own search direction at a badly conditioned state?  (CPU only, build container)
For recorded states of the teacher-forced sweep (tests/golden/sweep_*.npz): the unmodified reference Optimizer runs ONE
iteration from the state (oracle/refapi: Eigen SimplicialLDLT), its matrix, gradient and direction are read back, and the
direction is compared with a solution of the SAME system refined in 80-bit arithmetic (residuals in np.longdouble, corrections
by a sparse LU), i.e. the exact solution to ~1e-16.  Also printed: what a residual-1e-12 Krylov solution (Jacobi-PCG here)
of the same system is off by, and how far each direction's step bound / first accepted energy are from the exact direction's.
    python tools/ref_direction_accuracy.py bimba_cfg2 10 [bimba_cfg1 8 ...]"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refapi  # noqa: E402


def refine(A, b, x0, lu, iters=6):
    Ac = A.tocoo()
    r_, c_, v_ = Ac.row, Ac.col, Ac.data.astype(np.longdouble)
    x = x0.astype(np.longdouble)
    bl = b.astype(np.longdouble)
    for _ in range(iters):
        Ax = np.zeros(len(b), np.longdouble)
        np.add.at(Ax, r_, v_ * x[c_])
        res = bl - Ax
        x = x + lu.solve(res.astype(np.float64)).astype(np.longdouble)
    Ax = np.zeros(len(b), np.longdouble)
    np.add.at(Ax, r_, v_ * x[c_])
    return x, float(np.linalg.norm((bl - Ax).astype(np.float64)) / np.linalg.norm(b))


def all_states(name):
    """--all <name>: the reference direction's error at EVERY recorded state of a sweep -> tests/golden/sweep_<name>_ref_direction_error.json
    (what tests/test_gpu_sweep.py scales its per-state tolerance with)"""
    import ctypes as C
    import json
    g = np.load(os.path.join(ROOT, "tests", "golden", "sweep_%s.npz" % name))
    out = {}
    for k in [int(v) for v in g["iters"]]:
        p, m = "k%d_" % k, "m%d_" % int(g["k%d_mesh" % k])
        mesh = refapi.RefMesh(g[m + "V_rest"], g[m + "F"], g[p + "V"])
        opt = refapi.RefOptimizer(mesh, float(g[p + "p0"]), scaffolding=True)
        sz0 = opt.sizes()
        gvec = opt.recompute_gradient()
        opt.solve(1)
        d_ref = np.zeros(sz0["nSys"])
        refapi.lib().ref_opt_get_search_dir(opt.h, d_ref.ctypes.data_as(C.POINTER(C.c_double)))
        ia = np.zeros(sz0["nSys"] + 1, np.int32); ja = np.zeros(sz0["nnz"], np.int32); a = np.zeros(sz0["nnz"])
        refapi.lib().ref_opt_get_csr(opt.h, ia.ctypes.data_as(C.POINTER(C.c_int32)), ja.ctypes.data_as(C.POINTER(C.c_int32)), a.ctypes.data_as(C.POINTER(C.c_double)))
        n = len(ia) - 1
        U = sp.csr_matrix((a, ja, ia), shape=(n, n))
        A = (U + sp.triu(U, 1).T).tocsc()
        lu = spla.splu(A)
        x_ex, res = refine(A, -gvec, lu.solve(-gvec), lu)
        x_ex = x_ex.astype(np.float64)
        err = float(np.linalg.norm(d_ref - x_ex) / np.linalg.norm(x_ex))
        dg = A.diagonal()
        out[str(k)] = {"ref_direction_rel_error": err, "diag_min": float(dg.min()), "diag_max": float(dg.max()), "refined_residual": res}
        print(name, k, out[str(k)]); sys.stdout.flush()
        opt.close(); mesh.close()
    with open(os.path.join(ROOT, "tests", "golden", "sweep_%s_ref_direction_error.json" % name), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--all":
        return all_states(sys.argv[2])
    args = sys.argv[1:] or ["bimba_cfg2", "10"]
    for name, k in zip(args[0::2], args[1::2]):
        k = int(k)
        g = np.load(os.path.join(ROOT, "tests", "golden", "sweep_%s.npz" % name))
        p, m = "k%d_" % k, "m%d_" % int(g["k%d_mesh" % k])
        mesh = refapi.RefMesh(g[m + "V_rest"], g[m + "F"], g[p + "V"])
        p0 = float(g[p + "p0"])
        opt = refapi.RefOptimizer(mesh, p0, scaffolding=True)
        sz0 = opt.sizes()                       # the sparse system of this iteration stays in the solver; the scaffold is rebuilt after it
        gvec = opt.recompute_gradient()
        opt.solve(1)
        import ctypes as C
        d_ref = np.zeros(sz0["nSys"])
        refapi.lib().ref_opt_get_search_dir(opt.h, d_ref.ctypes.data_as(C.POINTER(C.c_double)))
        ia = np.zeros(sz0["nSys"] + 1, np.int32); ja = np.zeros(sz0["nnz"], np.int32); a = np.zeros(sz0["nnz"])
        refapi.lib().ref_opt_get_csr(opt.h, ia.ctypes.data_as(C.POINTER(C.c_int32)), ja.ctypes.data_as(C.POINTER(C.c_int32)), a.ctypes.data_as(C.POINTER(C.c_double)))
        n = len(ia) - 1
        U = sp.csr_matrix((a, ja, ia), shape=(n, n))        # the reference solver keeps 0-based upper-triangular CSR
        A = (U + sp.triu(U, 1).T).tocsc()
        b = -gvec
        lu = spla.splu(A)
        x_lu = lu.solve(b)
        x_ex, res_ex = refine(A, b, x_lu, lu)
        x_ex64 = x_ex.astype(np.float64)
        dg = A.diagonal()
        Minv = sp.diags(1.0 / dg)
        # Jacobi-PCG to 1e-12 relative residual
        it = [0]
        def cb(_): it[0] += 1
        x_cg, info = spla.cg(A, b, rtol=1e-12, atol=0.0, M=Minv, maxiter=200000, callback=cb)
        def rel(x): return float(np.linalg.norm(x - x_ex64) / np.linalg.norm(x_ex64))
        def relmax(x): return float(np.max(np.abs(x - x_ex64)) / np.max(np.abs(x_ex64)))
        def resid(x): return float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
        print("%s it %d: n %d, diag %.1e..%.1e, refined solution residual %.1e" % (name, k, n, dg.min(), dg.max(), res_ex))
        print("   reference LDL^T direction : residual %.2e, error vs exact %.2e (max-norm %.2e)" % (resid(d_ref), rel(d_ref), relmax(d_ref)))
        print("   scipy SuperLU             : residual %.2e, error vs exact %.2e (max-norm %.2e)" % (resid(x_lu), rel(x_lu), relmax(x_lu)))
        print("   Jacobi-PCG, rtol 1e-12    : residual %.2e, error vs exact %.2e (max-norm %.2e), %d iterations" % (resid(x_cg), rel(x_cg), relmax(x_cg), it[0]))
        # step bound of the whole system (mesh + air mesh, SymDirichletEnergy::initStepSize :551-610 through the oracle port) for each
        # direction: the smallest positive root over all triangles, decided by whichever (tiny air) triangle inverts first
        from oracle import portapi
        nV = g[p + "V"].shape[0]
        aV, aF, l2g = g[p + "air_V"], g[p + "air_F"], g[p + "air_localVI2Global"]
        nB = len(g[p + "air_bnd"])
        Xall = np.vstack([g[p + "V"], aV[nB:]])
        Fall = np.vstack([g[m + "F"], l2g[aF]])
        for tag, x in (("reference LDL^T", d_ref), ("exact", x_ex64), ("Jacobi-PCG 1e-12", x_cg)):
            am = portapi.init_step_size(g[m + "F"], g[p + "V"], x[:2 * nV], 1.0)
            aa = portapi.init_step_size(Fall, Xall, x, 1.0)
            print("   step bound along the %-17s direction: mesh only %.12g, mesh + air %.12g" % (tag, am, aa))
        sc = opt.scalars()
        print("   reference after the iteration: E %.17g (trace E_next %.17g)" % (sc["lastEnergyVal"], float(g[p + "E_next"][0])))
        opt.close(); mesh.close()


if __name__ == "__main__":
    main()
