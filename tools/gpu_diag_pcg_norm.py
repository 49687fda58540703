"""Accuracy of the PCG search direction against a dense/sparse LDL^T of the SAME device matrix (scipy), per DOF class
(mesh / air), for both stopping norms (plain 2-norm vs the block-Jacobi-scaled norm), on the torus states (soft scaffold
rows) and the bimba states.  Prints iteration counts too.   python tools/gpu_diag_pcg_norm.py > gpurun_out/..."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optcuts_b200 as ob  # noqa: E402


def states(name, picks):
    g = np.load(os.path.join(ROOT, "tests", "golden", "sweep_%s.npz" % name))
    w_scaf = 0.01 * (1.0 - float(g["lambda_init"]))
    for k in [int(i) for i in g["iters"]][:picks]:
        p, m = "k%d_" % k, "m%d_" % int(g["k%d_mesh" % k])
        yield name, k, g, p, m, w_scaf


def main():
    ctx = ob.Context(0)
    for name, picks in (("torus_cfg1", 4), ("bimba_cfg2", 2), ("bimba_cfg1", 3)):
        for name, k, g, p, m, w_scaf in states(name, picks):
            V_rest, F, UV = g[m + "V_rest"], g[m + "F"], g[p + "V"]
            rest8, sc = ctx.rest_features(V_rest, F)
            aV, aF = g[p + "air_V"], g[p + "air_F"]
            ctx.set_mesh(UV.shape[0], F, rest8, sc["surfaceArea"], g[m + "fixedVert"]); ctx.set_uv(UV)
            r8a, _ = ctx.rest_features(np.hstack([aV, np.zeros((len(aV), 1))]), aF, float(g[p + "air_scalars"][2]))
            ctx.set_air(aF, r8a, g[p + "air_localVI2Global"], len(g[p + "air_bnd"]), g[p + "air_fixedVert"], w_scaf / aF.shape[0]); ctx.set_uv(None, aV)
            p0 = float(g[p + "p0"])
            gr, _ = ctx.gradient(p0)
            ctx.set_pattern_from_elements(); ctx.hessian_assemble(p0)
            ia, ja, a = ctx.download_csr()
            n = len(ia) - 1
            U = sp.csr_matrix((a, ja - 1, ia - 1), shape=(n, n))
            A = (U + sp.triu(U, 1).T).tocsc()
            ref = spla.splu(A).solve(-gr)
            nm = 2 * UV.shape[0]
            line = "%s it %d: n %d, diag range %.1e..%.1e" % (name, k, n, A.diagonal().min(), A.diagonal().max())
            for plain in (1, 0):
                ctx.set_option("pcg_scaled_norm", 1 - plain)
                ctx.factorize()
                x, info = ctx.solve(None, 1e-12, 0)
                em = np.linalg.norm(x[:nm] - ref[:nm]) / np.linalg.norm(ref[:nm])
                ea = np.linalg.norm(x[nm:] - ref[nm:]) / max(np.linalg.norm(ref[nm:]), 1e-300)
                emax = np.max(np.abs(x - ref) / (np.abs(ref) + 1e-12 * np.max(np.abs(ref))))
                line += " | %s: %d it, mesh %.1e air %.1e worst component %.1e" % ("plain" if plain else "scaled", info["iters"], em, ea, emax)
            print(line)
    ctx.close()


if __name__ == "__main__":
    main()
