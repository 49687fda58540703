"""CPU prototype of the multilevel additive Schwarz (MAS) preconditioner used by the GPU PCG
(optcuts_b200/csrc/ocb_mas.cu).  Design tool only: measures CG iteration counts on the reference's
own matrices (tests/golden) and on subdivided meshes assembled by the oracle port.

    python tools/mas_proto.py [s1|s100|x2|x4] [--leaf 32] [--group 8] [--grid 16] [--prec f32]
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_golden(tag):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bimba_cfg2_states.npz"))
    s, r = ("s1_", "r1_") if tag == "s1" else ("s100_", "r100_")
    UV, airV, l2g = g[s + "V"], g[s + "air_V"], g[s + "air_localVI2Global"]
    nB = len(g[s + "air_bnd"])
    X = np.vstack([UV, airV[nB:]])
    if tag == "s1":
        ia, ja, a = g[r + "ia"].astype(np.int64), g[r + "ja"].astype(np.int64), g[r + "a"]
        n = len(ia) - 1
        U = sp.csr_matrix((a, ja, ia), shape=(n, n))
        A = U + sp.triu(U, 1).T
    else:
        from oracle import portapi
        A = assemble(g[s + "F"], UV, g[r + "rest8"], float(g[r + "surfaceArea"]), g[s + "fixedVert"], float(g["energyParam0"]),
                     air=dict(F=g[s + "air_F"], V=airV, l2g=l2g, rest8=g[r + "air_rest8"], w=float(g[r + "w_scaf"])), nTot=X.shape[0])
    b = -g[r + "gradient"]
    fixed = np.zeros(X.shape[0], bool)
    fixed[g[s + "fixedVert"]] = True
    return A.tocsr(), b, X, fixed


def assemble(F, UV, rest8, surf, fixed, p0, air=None, nTot=None):
    from oracle import portapi
    nV = UV.shape[0]
    nTot = nTot or nV
    I, J, S = portapi.hessian_triplets(F, UV, rest8, surf, False, fixed)
    A = sp.coo_matrix((p0 * S, (I, J)), shape=(2 * nTot, 2 * nTot)).tocsr()
    fx = set(int(v) for v in fixed)
    if air is not None:
        Ia, Ja, Sa = portapi.hessian_triplets(air["F"], air["V"], air["rest8"], 1.0, True, ())
        m = air["l2g"].astype(np.int64)
        gi, gj = 2 * m[Ia // 2] + Ia % 2, 2 * m[Ja // 2] + Ja % 2
        keep = np.array([(i // 2 not in fx) and (j // 2 not in fx) for i, j in zip(gi, gj)])
        A = A + sp.coo_matrix((air["w"] / air["F"].shape[0] * Sa[keep], (gi[keep], gj[keep])), shape=A.shape).tocsr()
    return A


def load_subdiv(k):
    from oracle import portapi
    from optcuts_b200 import synth
    g = np.load(os.path.join(ROOT, "tests", "golden", "bimba_cfg2_states.npz"))
    Vr, F, UV = synth.subdivide(g["s1_V_rest"], g["s1_F"], g["s1_V"], k)
    rest8, sc, _ = portapi.rest_features(Vr, F)
    A = assemble(F, UV, rest8, sc["surfaceArea"], [0], float(g["energyParam0"]))
    b = -float(g["energyParam0"]) * portapi.gradient(F, UV, rest8, sc["surfaceArea"], False, [0])
    fixed = np.zeros(UV.shape[0], bool)
    fixed[0] = True
    return A, b, UV, fixed


# ------------------------------------------------------------------------------------------------
def rcb(X, idx, nparts_sizes):
    """Recursive coordinate bisection of points idx into consecutive parts with the given sizes."""
    out = []

    def rec(ids, sizes):
        if len(sizes) == 1:
            out.append(ids)
            return
        h = len(sizes) // 2
        nl = int(sum(sizes[:h]))
        P = X[ids]
        ax = int(np.argmax(P.max(0) - P.min(0)))
        o = np.argpartition(P[:, ax], nl - 1) if 0 < nl < len(ids) else np.arange(len(ids))
        rec(ids[o[:nl]], sizes[:h])
        rec(ids[o[nl:]], sizes[h:])
    rec(idx, list(nparts_sizes))
    return out


def hilbert_order(X, bits=int(os.environ.get("HBITS", "10"))):
    """Order points along a Hilbert curve of a 2^bits x 2^bits grid (alternative to recursive bisection)."""
    lo, hi = X.min(0), X.max(0)
    ext = max(float((hi - lo).max()), 1e-300)
    q = np.minimum(((X - lo) / ext * (1 << bits)).astype(np.int64), (1 << bits) - 1)
    x, y = q[:, 0].copy(), q[:, 1].copy()
    d = np.zeros(len(X), np.int64)
    s = 1 << (bits - 1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate
        m = ry == 0
        flip = m & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        x, y = np.where(m, y, x), np.where(m, x, y)
        s >>= 1
    return np.argsort(d, kind="stable")


def build_hierarchy_curve(X, grid, leaf, group):
    """Hierarchy from a space-filling-curve order: CTA chunks, leaves and groups are consecutive runs."""
    n = X.shape[0]
    order = hilbert_order(X)
    rowsPer = -(-n // grid)
    leaves, ctaOf = [], []
    for c in range(grid):
        b, e = c * rowsPer, min(n, (c + 1) * rowsPer)
        if e <= b:
            continue
        nl = -(-(e - b) // leaf)
        o = b
        for sz in even_sizes(e - b, nl):
            leaves.append((o, o + sz)); ctaOf.append(c); o += sz
    levels = [dict(rng=leaves, cta=ctaOf)]
    while len(levels[-1]["rng"]) > 1:
        cur = levels[-1]
        rng, cta, parent = [], [], []
        nn = len(cur["rng"])
        multi = len(set(cur["cta"])) < nn
        i = 0
        while i < nn:
            # even group sizes inside a CTA
            j = i
            while j < nn and (not multi or cur["cta"][j] == cur["cta"][i]):
                j += 1
            k = j - i
            for sz in even_sizes(k, -(-k // group)):
                rng.append((cur["rng"][i][0], cur["rng"][i + sz - 1][1])); cta.append(cur["cta"][i]); parent.extend([len(rng) - 1] * sz); i += sz
            if not multi:
                break
        cur["parent"] = parent
        levels.append(dict(rng=rng, cta=cta))
    levels[-1]["parent"] = [0]
    return order, levels


def even_sizes(n, k):
    return [n // k + (1 if i < n % k else 0) for i in range(k)]


def build_hierarchy(X, grid, leaf, group):
    """Returns order (new -> old vertex) and, per level >= 1, node ranges [beg, end) in the new order
    plus parent pointers.  Level 1 nodes = leaves.  CTA boundaries are group boundaries at every level."""
    n = X.shape[0]
    rowsPer = -(-n // grid)
    sizes = [min(rowsPer, n - i * rowsPer) for i in range(grid) if n - i * rowsPer > 0]
    chunks = rcb(X, np.arange(n), sizes)
    order, leaves, ctaOf = [], [], []
    for c, ids in enumerate(chunks):
        nl = -(-len(ids) // leaf)
        for lf in rcb(X, ids, even_sizes(len(ids), nl)):
            leaves.append((len(order), len(order) + len(lf)))
            order.extend(lf.tolist())
            ctaOf.append(c)
    levels = [dict(rng=leaves, cta=ctaOf)]
    while len(levels[-1]["rng"]) > 1:
        cur = levels[-1]
        rng, cta, parent = [], [], []
        i = 0
        nn = len(cur["rng"])
        # group consecutive nodes, never across a CTA boundary while a CTA still has > 1 node
        multi = len(set(cur["cta"])) < nn
        while i < nn:
            j = i + 1
            while j < nn and j - i < group and (not multi or cur["cta"][j] == cur["cta"][i]):
                j += 1
            rng.append((cur["rng"][i][0], cur["rng"][j - 1][1]))
            cta.append(cur["cta"][i])
            parent.extend([len(rng) - 1] * (j - i))
            i = j
        cur["parent"] = parent
        levels.append(dict(rng=rng, cta=cta))
    levels[-1]["parent"] = [0]
    return np.array(order), levels


def lib_hierarchy(X, grid):
    """The hierarchy the CUDA library builds (ocb_precond_hierarchy), in this prototype's format."""
    from optcuts_b200 import _capi
    vert_of, lv, lloc = _capi.precond_hierarchy(X, grid)
    levels = []
    rng = [(int(lv[0][k]), int(lv[0][k + 1])) for k in range(len(lv[0]) - 1)]
    levels.append(dict(rng=rng))
    for l in range(1, len(lv)):
        cb = lv[l]
        parent = np.zeros(len(levels[-1]["rng"]), np.int64)
        nr = []
        for k in range(len(cb) - 1):
            parent[cb[k]:cb[k + 1]] = k
            nr.append((levels[-1]["rng"][cb[k]][0], levels[-1]["rng"][cb[k + 1] - 1][1]))
        levels[-1]["parent"] = parent.tolist()
        levels.append(dict(rng=nr))
    levels[-1]["parent"] = [0] * len(levels[-1]["rng"])
    return np.asarray(vert_of, np.int64), levels


def mas_setup(A, X, fixed, order, levels, basis=3, prec="f64", leaf_dense=True, exact_from=99):
    """Returns a function r -> z.  A in the NEW order."""
    n2 = A.shape[0]
    ops = []

    def quant(M):
        M = 0.5 * (M + M.T)
        if prec == "f32":
            return M.astype(np.float32).astype(np.float64)
        if prec == "f16":
            s = np.abs(M).max()
            return (M / s).astype(np.float16).astype(np.float64) * s
        if prec == "hi32":                  # the high 32 bits of the fp64 value (20-bit mantissa), rounded to nearest
            v = np.ascontiguousarray(M, dtype=np.float64).view(np.uint64)
            v = (v + np.uint64(0x80000000)) & np.uint64(0xFFFFFFFF00000000)
            return v.view(np.float64)
        if prec == "bf16":
            v = M.astype(np.float32).view(np.uint32)
            v = ((v + 0x8000) & 0xFFFF0000).astype(np.uint32)
            return v.view(np.float32).astype(np.float64)
        return M
    # level 0
    blocks = []
    if leaf_dense:
        for (b, e) in levels[0]["rng"]:
            blocks.append(quant(np.linalg.inv(A[2 * b:2 * e, 2 * b:2 * e].toarray())))
    else:
        for v in range(n2 // 2):
            blocks.append(np.linalg.inv(A[2 * v:2 * v + 2, 2 * v:2 * v + 2].toarray()))
    ops.append((None, sp.block_diag(blocks, format="csr")))
    # coarse levels: nodes of level l carry `2*basis` DOFs; the solve blocks are the groups (= nodes of level l+1)
    free = (~fixed).astype(np.float64)
    for l in range(len(levels)):
        rng = levels[l]["rng"]
        rows, cols, vals = [], [], []
        for k, (b, e) in enumerate(rng):
            P = X[b:e]
            c = 0.5 * (P.max(0) + P.min(0))
            s = max(1e-300, 0.5 * float((P.max(0) - P.min(0)).max()))
            fn = [np.ones(e - b), (P[:, 0] - c[0]) / s, (P[:, 1] - c[1]) / s][:basis]
            for comp in range(2):
                for q, f in enumerate(fn):
                    rows.append(2 * np.arange(b, e) + comp)
                    cols.append(np.full(e - b, 2 * basis * k + comp * basis + q))
                    vals.append(f * free[b:e])
        Pm = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n2, 2 * basis * len(rng))).tocsr()
        Al = (Pm.T @ A @ Pm).tocsr()
        if l + 1 >= exact_from:             # exact coarse solve on this level's Galerkin matrix, no levels above
            D = Al.toarray()
            d = np.diag(D).copy(); bad = d <= 1e-300 * max(1.0, d.max()); D[bad, bad] = 1.0
            ops.append((Pm, sp.csr_matrix(quant(np.linalg.pinv(D, hermitian=True)))))
            break
        parent = np.asarray(levels[l]["parent"])
        blocks = []
        for gidx in range(parent.max() + 1):
            ch = np.nonzero(parent == gidx)[0]
            sl = slice(2 * basis * ch[0], 2 * basis * (ch[-1] + 1))
            D = Al[sl, sl].toarray()
            # nodes made of fixed vertices only (or degenerate): regularise
            d = np.diag(D).copy()
            bad = d <= 1e-300 * max(1.0, d.max())
            D[bad, bad] = 1.0
            blocks.append(quant(np.linalg.pinv(D, hermitian=True)))
        ops.append((Pm, sp.block_diag(blocks, format="csr")))

    def apply(r):
        z = ops[0][1] @ r
        for Pm, Dinv in ops[1:]:
            z = z + Pm @ (Dinv @ (Pm.T @ r))
        return z
    return apply


def pcg(A, b, M, tol=1e-12, maxit=100000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    d = z.copy()
    rz = r @ z
    bb = b @ b
    for it in range(1, maxit + 1):
        Ad = A @ d
        alpha = rz / (d @ Ad)
        x += alpha * d
        r -= alpha * Ad
        if r @ r <= tol * tol * bb:
            return x, it
        z = M(r)
        rz2 = r @ z
        d = z + (rz2 / rz) * d
        rz = rz2
    return x, maxit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", nargs="?", default="s1")
    ap.add_argument("--leaf", type=int, default=32)
    ap.add_argument("--group", type=int, default=8)
    ap.add_argument("--grid", type=int, default=16)
    ap.add_argument("--basis", type=int, default=3)
    ap.add_argument("--prec", default="f64")
    ap.add_argument("--no-leaf-dense", action="store_true")
    ap.add_argument("--jacobi", action="store_true")
    ap.add_argument("--exact-from", type=int, default=99, help="level (1-based) solved exactly on its Galerkin matrix")
    ap.add_argument("--curve", action="store_true", help="Hilbert-curve order instead of recursive bisection")
    ap.add_argument("--lib", action="store_true", help="use the hierarchy built by liboptcuts_b200.so")
    a = ap.parse_args()
    t0 = time.time()
    if a.case in ("s1", "s100"):
        A, b, X, fixed = load_golden(a.case)
    else:
        A, b, X, fixed = load_subdiv(int(a.case[1:]))
    n = X.shape[0]
    print("case %s: %d vertices, nnz %d (%.1fs)" % (a.case, n, A.nnz, time.time() - t0))
    if a.jacobi:
        blocks = [np.linalg.inv(A[2 * v:2 * v + 2, 2 * v:2 * v + 2].toarray()) for v in range(n)]
        Minv = sp.block_diag(blocks, format="csr")
        _, it = pcg(A, b, lambda r: Minv @ r)
        print("block-Jacobi: %d iterations" % it)
    order, levels = lib_hierarchy(X, a.grid) if a.lib else (build_hierarchy_curve if a.curve else build_hierarchy)(X, a.grid, a.leaf, a.group)
    if a.lib:
        a.exact_from = len(levels)          # the library's last level is the exactly solved coarse level
    perm2 = np.stack([2 * order, 2 * order + 1], 1).ravel()
    Ap = A[perm2][:, perm2].tocsr()
    t0 = time.time()
    M = mas_setup(Ap, X[order], fixed[order], order, levels, a.basis, a.prec, not a.no_leaf_dense, a.exact_from)
    t1 = time.time()
    x, it = pcg(Ap, b[perm2], M)
    print("MAS leaf %d group %d grid %d basis %d prec %s: levels %s -> %d iterations (setup %.1fs, solve %.1fs) resid %.2e"
          % (a.leaf, a.group, a.grid, a.basis, a.prec, [len(l["rng"]) for l in levels], it, t1 - t0, time.time() - t1,
             np.linalg.norm(Ap @ x - b[perm2]) / np.linalg.norm(b)))


if __name__ == "__main__":
    main()
