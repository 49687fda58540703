"""Local stencils in the shape the reference's candidate evaluation builds them (TriMesh.cpp:2270-2300): the triangles
around 1-2 vertices, local vertex numbering in first-seen order, everything fixed but the centre vertices."""
import numpy as np


def one_ring(F, V_rest, UV, centers):
    centers = list(centers)
    tris = np.nonzero(np.isin(F, centers).any(axis=1))[0]
    g2l, lF = {}, np.zeros((len(tris), 3), np.int32)
    for i, t in enumerate(tris):
        for k in range(3):
            v = int(F[t, k])
            if v not in g2l:
                g2l[v] = len(g2l)
            lF[i, k] = g2l[v]
    ids = np.array(list(g2l.keys()))
    free = np.isin(ids, centers)
    return np.ascontiguousarray(V_rest[ids]), np.ascontiguousarray(UV[ids]), lF, free


def port_local_solve(port, V_rest, F, UV, free, tol=1e-6, maxIter=100):
    """Optimizer::precompute + setRelGL2Tol(tol) + solve(maxIter) on a local mesh, restated with the oracle port
    (energyParams = {1}, Optimizer.cpp:203-261, 675-678)."""
    r8, sc, _ = port.rest_features(V_rest, F)
    fixed = np.nonzero(~np.asarray(free, bool))[0].astype(np.int32)
    nV = len(UV)
    tg = (nV - len(fixed)) / nV * tol
    uv = np.array(UV, dtype=np.float64)
    E0 = port.energy(F, uv, r8, sc["surfaceArea"])
    it = 0
    while it < maxIter:                       # globalIterNum counts every pass, the converged one included (Optimizer.cpp:215-229)
        uv, _, p, res = port.newton_step(F, uv, r8, sc["surfaceArea"], fixed, 1.0, tg)
        it += 1
        if res["converged"] or res["stopped"]:
            break
    return dict(E_init=E0, E_final=port.energy(F, uv, r8, sc["surfaceArea"]), iters=it, UV=uv)
