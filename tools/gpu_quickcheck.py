"""Ad-hoc first-light check of every C-ABI entry point on a GPU against the oracle port + golden fixtures."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optcuts_b200 as ob
from optcuts_b200 import synth
from oracle import portapi

g = np.load("tests/golden/bimba_cfg2_states.npz")
p0 = float(g["energyParam0"])
ctx = ob.Context(0)
print(ctx.version())

def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / (np.max(np.abs(b)) + 1e-300))

for tag, rt in (("s1_", "r1_"), ("s100_", "r100_")):
    Vr, F, UV = g[tag + "V_rest"], g[tag + "F"], g[tag + "V"]
    nV = UV.shape[0]
    r8, sc = ctx.rest_features(Vr, F)
    print(tag, "features bitexact", np.array_equal(r8, g[rt + "rest8"]), sc["surfaceArea"] - g[rt + "surfaceArea"], sc["avgEdgeLen"] - g[rt + "avgEdgeLen"])
    ctx.set_mesh(nV, F, g[rt + "rest8"], float(g[rt + "surfaceArea"]), g[tag + "fixedVert"])
    ctx.set_uv(UV)
    per = ctx.energy_per_elem()
    pper = portapi.energy_per_elem(F, UV, g[rt + "rest8"], float(g[rt + "surfaceArea"]))
    print(tag, "energy per elem bitexact vs port", np.array_equal(per, pper))
    # scaffold
    air_rest8 = g[rt + "air_rest8"]
    w_scaf = float(g[rt + "w_scaf"])
    Fa = g[tag + "air_F"]
    ctx.set_air(Fa, air_rest8, g[tag + "air_localVI2Global"], len(g[tag + "air_bnd"]), g[rt + "air_fixed"], w_scaf / Fa.shape[0])
    ctx.set_uv(None, g[tag + "air_V"])
    et, esd, escaf = ctx.energy(p0)
    print(tag, "E", et, "ref", float(g[rt + "E_last"]), "rel", abs(et - g[rt + "E_last"]) / g[rt + "E_last"], "Escaf", escaf, float(g[rt + "E_scaf_last"]))
    gr, sq = ctx.gradient(p0)
    print(tag, "gradient rel", rel(gr, g[rt + "gradient"]), "sqn", sq, float(g[rt + "sqn_g"]))
    ctx.set_pattern_from_elements()
    print(tag, "sizes", ctx.sizes())
    t = time.time(); ctx.hessian_assemble(p0); ctx.synchronize(); print("assemble s", time.time() - t)
    ia, ja, a = ctx.download_csr()
    if rt + "a" in g:
        print(tag, "ia eq", np.array_equal(ia, g[rt + "ia"] + 1), "ja eq", np.array_equal(ja, g[rt + "ja"] + 1), "a rel", rel(a, g[rt + "a"]))
    else:
        print(tag, "a sum rel", abs(a.sum() - g[rt + "a_sum"]) / abs(g[rt + "a_sum"]), abs(np.abs(a).sum() - g[rt + "a_abs_sum"]) / g[rt + "a_abs_sum"])
    # spmv vs scipy-free check: y = A x using csr
    x = np.random.default_rng(0).standard_normal(len(ia) - 1)
    y = ctx.multiply(x)
    yr = np.zeros_like(x)
    for i in range(len(ia) - 1):
        seg = slice(ia[i] - 1, ia[i + 1] - 1)
        cols = ja[seg] - 1
        yr[i] += a[seg] @ x[cols]
        off = cols != i
        np.add.at(yr, cols[off], a[seg][off] * x[i])
    print(tag, "spmv rel", rel(y, yr))
    ctx.factorize()
    t = time.time()
    p, info = ctx.solve(None, 1e-12, 0, allow_not_converged=True)
    print(tag, "pcg", info, "time", time.time() - t, "searchDir rel", np.linalg.norm(p - g[rt + "searchDir"]) / np.linalg.norm(g[rt + "searchDir"]))
    res = ctx.multiply(p) + gr
    print(tag, "true residual rel", np.linalg.norm(res) / np.linalg.norm(gr))
    alpha = ctx.step_bound(None, 1.0)
    print(tag, "step bound", alpha, "x0.99", alpha * 0.99, "ref alpha", float(g[rt + "alpha"]))
    ls = ctx.line_search(p0, 0.0, alpha * 0.99)
    print(tag, "line search", ls)
    print(tag, "E_new rel", abs(ls["E_new"] - g[rt + "E_new"]) / g[rt + "E_new"], "E_sd_new rel", abs(ls["E_sd_new"] - g[rt + "E_sd_new"]) / g[rt + "E_sd_new"], "lastEDec", ls["lastEDec"], float(g[rt + "lastEDec"]))
    nxt = "s2_" if tag == "s1_" else "s101_"
    print(tag, "UV after step: max abs diff vs reference next state", np.max(np.abs(ctx.get_uv() - g[nxt + "V"])), "scale", np.max(np.abs(g[nxt + "V"])))
    dg = ctx.divgrad_scores()   # at the NEW uv; compare with port
    print(tag, "divgrad rel vs port", rel(dg, portapi.divgrad(F, ctx.get_uv(), g[rt + "rest8"], float(g[rt + "surfaceArea"]))))
    H = ctx.hessian_blocks()
    Hp = portapi.hessian_blocks(F, ctx.get_uv(), g[rt + "rest8"], float(g[rt + "surfaceArea"]))
    print(tag, "hessian blocks rel-to-block vs port", float(np.max(np.abs(H - Hp).reshape(len(F), -1).max(axis=1) / np.abs(Hp).reshape(len(F), -1).max(axis=1))))
    if len(g[tag + "cohE"]):
        ctx.set_uv(UV)
        ese = ctx.seam_energy(g[tag + "cohE"], g[rt + "edgeLen"], g[rt + "boundaryEdge"], 0.0, float(g[rt + "virtualRadius"]), float(g[rt + "avgEdgeLen"]), False)
        print(tag, "seam", ese * float(g[rt + "virtualRadius"]), float(g[rt + "seam_sparsity"]))
        ese = ctx.seam_energy(g[tag + "cohE"], g[rt + "edgeLen"], g[rt + "boundaryEdge"], 0.0, float(g[rt + "virtualRadius"]), float(g[rt + "avgEdgeLen"]), True)
        print(tag, "seam soup", ese * float(g[rt + "virtualRadius"]), float(g[rt + "seam_sparsity_soup"]))

# newton_step end to end from s1
tag, rt = "s1_", "r1_"
ctx.set_mesh(g["s1_V"].shape[0], g["s1_F"], g["r1_rest8"], float(g["r1_surfaceArea"]), g["s1_fixedVert"])
ctx.set_uv(g["s1_V"])
Fa = g["s1_air_F"]
ctx.set_air(Fa, g["r1_air_rest8"], g["s1_air_localVI2Global"], len(g["s1_air_bnd"]), g["r1_air_fixed"], float(g["r1_w_scaf"]) / Fa.shape[0])
ctx.set_uv(None, g["s1_air_V"])
t = time.time()
r = ctx.newton_step(p0, float(g["r1_targetGRes"]))
print("newton_step", r, "wall", time.time() - t)
print("E_new rel", abs(r["E_new"] - g["r1_E_new"]) / g["r1_E_new"])

# scale test
for n in (4, 10):
    Vr, F, UV = synth.subdivide(g["s1_V_rest"], g["s1_F"], g["s1_V"], n)
    t = time.time(); r8, sc = ctx.rest_features(Vr, F); print("n", n, "faces", len(F), "features", time.time() - t)
    ctx.set_mesh(UV.shape[0], F, r8, sc["surfaceArea"], [0]); ctx.set_uv(UV)
    t = time.time(); ctx.set_pattern_from_elements(); print(" pattern s", time.time() - t, ctx.sizes())
    for rep in range(2):
        ctx.timer_start(); e = ctx.energy(p0); ms = ctx.timer_stop_ms(); print(" energy", e[0], "ms", ms)
    ctx.timer_start(); gr, sq = ctx.gradient(p0, download=False); print(" gradient ms", ctx.timer_stop_ms(), sq)
    ctx.timer_start(); ctx.hessian_assemble(p0); print(" hessian ms", ctx.timer_stop_ms())
    ctx.timer_start(); ctx.factorize(); print(" jacobi ms", ctx.timer_stop_ms())
    ctx.timer_start(); p, info = ctx.solve(None, 1e-8, 2000, download=False, allow_not_converged=True); ms = ctx.timer_stop_ms()
    print(" pcg", info, "ms", ms, "us/iter", 1e3 * ms / max(1, info["iters"]))
print("launches", ctx.launch_count())
