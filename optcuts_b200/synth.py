"""Synthetic workloads: n-section subdivision of a UV-parameterised triangle mesh.

Every triangle is split into n^2 congruent sub-triangles (barycentric grid); rest positions and UVs
are interpolated linearly, so the subdivided map has exactly the distortion distribution of the
input (a Tutte start stays a Tutte start) and stays inversion-free.  Edge points are shared between
the two triangles of an edge; seams (duplicated UV vertices) stay seams.
"""
import numpy as np


def subdivide(V_rest, F, UV, n):
    V_rest = np.asarray(V_rest, dtype=np.float64)
    UV = np.asarray(UV, dtype=np.float64)
    F = np.asarray(F, dtype=np.int64)
    if n <= 1:
        return V_rest.copy(), F.astype(np.int32), UV.copy()
    nV, nF = V_rest.shape[0], F.shape[0]
    # undirected edges
    he = np.stack([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]], axis=1).reshape(-1, 2)      # (3nF, 2): sides ab, bc, ca
    lo, hi = he.min(axis=1), he.max(axis=1)
    key = lo * nV + hi
    ukey, eid = np.unique(key, return_inverse=True)
    nE = len(ukey)
    eid = eid.reshape(nF, 3)
    fwd = (he[:, 0] == lo).reshape(nF, 3)            # side runs from the min-vertex end?
    elo, ehi = ukey // nV, ukey % nV
    nEdgePts = n - 1
    nIntPts = (n - 1) * (n - 2) // 2
    nVnew = nV + nE * nEdgePts + nF * nIntPts

    def lerp(A):
        out = np.empty((nVnew, A.shape[1]))
        out[:nV] = A
        if nEdgePts:
            t = (np.arange(1, n) / n)[None, :, None]
            out[nV:nV + nE * nEdgePts] = (A[elo][:, None, :] * (1 - t) + A[ehi][:, None, :] * t).reshape(-1, A.shape[1])
        return out

    Vr, Uv = lerp(V_rest), lerp(UV)
    # local grid index -> global id
    tri = np.arange(nF)
    gid = {}
    intBase = nV + nE * nEdgePts
    intCount = 0
    a, b, c = F[:, 0], F[:, 1], F[:, 2]

    def edge_pt(side, t):   # t = steps (1..n-1) from the START vertex of the side
        tt = np.where(fwd[:, side], t, n - t)
        return nV + eid[:, side] * nEdgePts + (tt - 1)

    for i in range(n + 1):
        for j in range(n + 1 - i):
            k = n - i - j
            if i == 0 and j == 0:
                g = a
            elif i == n:
                g = b
            elif j == n:
                g = c
            elif j == 0:          # on ab, i steps from a
                g = edge_pt(0, i)
            elif k == 0:          # on bc, j steps from b
                g = edge_pt(1, j)
            elif i == 0:          # on ca, (n - j) steps from c
                g = edge_pt(2, n - j)
            else:
                g = intBase + tri * nIntPts + intCount
                w = np.array([k, i, j]) / n
                Vr[g] = w[0] * V_rest[a] + w[1] * V_rest[b] + w[2] * V_rest[c]
                Uv[g] = w[0] * UV[a] + w[1] * UV[b] + w[2] * UV[c]
                intCount += 1
            gid[(i, j)] = g
    tris = []
    for i in range(n):
        for j in range(n - i):
            tris.append(np.stack([gid[(i, j)], gid[(i + 1, j)], gid[(i, j + 1)]], axis=1))
            if i + j <= n - 2:
                tris.append(np.stack([gid[(i + 1, j)], gid[(i + 1, j + 1)], gid[(i, j + 1)]], axis=1))
    # keep the sub-triangles of one parent contiguous (locality)
    Fn = np.stack(tris, axis=1).reshape(-1, 3)
    return Vr, Fn.astype(np.int32), Uv


def locality_order(UV, F):
    """Renumber vertices and triangles along a Morton curve of the UV domain (better gather locality).
    Returns (perm_v, perm_f): new vertex i = old perm_v[i]."""
    UV = np.asarray(UV)
    q = ((UV - UV.min(axis=0)) / (np.ptp(UV, axis=0) + 1e-300) * 65535).astype(np.uint64)

    def spread(x):
        x = (x | (x << 16)) & 0x0000FFFF0000FFFF
        x = (x | (x << 8)) & 0x00FF00FF00FF00FF
        x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0F
        x = (x | (x << 2)) & 0x3333333333333333
        x = (x | (x << 1)) & 0x5555555555555555
        return x
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1)
    perm_v = np.argsort(code, kind="stable")
    inv = np.empty_like(perm_v)
    inv[perm_v] = np.arange(len(perm_v))
    Fn = inv[np.asarray(F)]
    perm_f = np.argsort(Fn.min(axis=1), kind="stable")
    return perm_v, perm_f


def apply_order(V_rest, F, UV, perm_v, perm_f):
    inv = np.empty_like(perm_v)
    inv[perm_v] = np.arange(len(perm_v))
    return V_rest[perm_v], inv[np.asarray(F)][perm_f].astype(np.int32), UV[perm_v]
