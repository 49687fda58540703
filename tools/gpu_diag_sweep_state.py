"""Diagnostic (GPU box): at recorded sweep states, the device's search direction against the solution of the SAME device matrix
refined in 80-bit arithmetic, the step bound along both, and the whole Newton iteration's fields, for a few solver settings.
    python tools/gpu_diag_sweep_state.py bimba_cfg2 9,10,12 [bimba_cfg1 8]"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import optcuts_b200 as ob  # noqa: E402
from ref_direction_accuracy import refine  # noqa: E402


def upload(ctx, g, k):
    p, m = "k%d_" % k, "m%d_" % int(g["k%d_mesh" % k])
    w_scaf = 0.01 * (1.0 - float(g["lambda_init"]))
    V_rest, F, UV = g[m + "V_rest"], g[m + "F"], g[p + "V"]
    rest8, sc = ctx.rest_features(V_rest, F)
    aV, aF = g[p + "air_V"], g[p + "air_F"]
    ctx.set_mesh(UV.shape[0], F, rest8, sc["surfaceArea"], g[m + "fixedVert"]); ctx.set_uv(UV)
    r8a, _ = ctx.rest_features(np.hstack([aV, np.zeros((len(aV), 1))]), aF, float(g[p + "air_scalars"][2]))
    ctx.set_air(aF, r8a, g[p + "air_localVI2Global"], len(g[p + "air_bnd"]), g[p + "air_fixedVert"], w_scaf / aF.shape[0]); ctx.set_uv(None, aV)
    return float(g[p + "p0"]), UV.shape[0]


def main():
    ctx = ob.Context(0)
    args = sys.argv[1:] or ["bimba_cfg2", "9,10,12"]
    for name, ks in zip(args[0::2], args[1::2]):
        g = np.load(os.path.join(ROOT, "tests", "golden", "sweep_%s.npz" % name))
        for k in [int(v) for v in ks.split(",")]:
            if ("k%d_V" % k) not in g.files:
                print(name, k, "not recorded"); continue
            p0, nV = upload(ctx, g, k)
            gr, _ = ctx.gradient(p0)
            ctx.set_pattern_from_elements(); ctx.hessian_assemble(p0)
            ia, ja, a = ctx.download_csr()
            n = len(ia) - 1
            U = sp.csr_matrix((a, ja - 1, ia - 1), shape=(n, n))
            A = (U + sp.triu(U, 1).T).tocsc()
            lu = spla.splu(A)
            x_ex, res = refine(A, -gr, lu.solve(-gr), lu)
            x_ex = x_ex.astype(np.float64)
            sb_ex = ctx.step_bound(x_ex)
            E_ref, alpha_ref = float(g["k%d_E_next" % k][0]), None
            print("%s it %d: n %d diag %.1e..%.1e; exact direction: step bound %.12g (x0.99 = %.12g)" % (name, k, n, A.diagonal().min(), A.diagonal().max(), sb_ex, 0.99 * sb_ex))
            for tag, opts, tol in (("default", {}, 1e-12), ("tol 1e-14", {}, 1e-14), ("scaled norm", {"pcg_scaled_norm": 1}, 1e-12), ("scaled norm, 1e-14", {"pcg_scaled_norm": 1}, 1e-14)):
                for kk, vv in opts.items():
                    ctx.set_option(kk, vv)
                ctx.factorize()
                x, info = ctx.solve(None, tol, 0, allow_not_converged=True)
                err = np.linalg.norm(x - x_ex) / np.linalg.norm(x_ex)
                emax = np.max(np.abs(x - x_ex)) / np.max(np.abs(x_ex))
                rres = np.linalg.norm(A @ x + gr) / np.linalg.norm(gr)
                sb = ctx.step_bound(x)
                print("   %-20s %5d CG it, status %d, rel res (device) %.1e (true) %.1e | error vs exact %.1e (max-norm %.1e) | step bound %.12g (%+.1e rel.)"
                      % (tag, info["iters"], info["status"], info["rel_res"], rres, err, emax, sb, sb / sb_ex - 1.0))
                for kk in opts:
                    ctx.set_option(kk, 0)
            upload(ctx, g, k)
            r = ctx.newton_step(p0, 0.0)
            print("   newton_step: alpha_init %.12g alpha %.12g halvings %d pcg %d it status %d rel_res %.1e E_new %.15g (reference %.15g, rel %.1e)"
                  % (r["alpha_init"], r["alpha"], r["n_halvings"], r["pcg_iters"], r["pcg_status"], r["pcg_rel_res"], r["E_new"], E_ref, abs(r["E_new"] - E_ref) / E_ref))
            sys.stdout.flush()
    ctx.close()


if __name__ == "__main__":
    main()
