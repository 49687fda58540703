"""Per-section host wall-clock of the e2e Newton iteration (OCB_HOST_TIMING=1), bimba 10k state 1."""
import os, sys, time
os.environ["OCB_HOST_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optcuts_b200 as ob

g = np.load("tests/golden/bimba_cfg2_states.npz")
p0 = float(g["energyParam0"])
ctx = ob.Context(0)
tag, rt = "s1_", "r1_"
Fa = g[tag + "air_F"]
ctx.set_mesh(g[tag + "V"].shape[0], g[tag + "F"], g[rt + "rest8"], float(g[rt + "surfaceArea"]), g[tag + "fixedVert"])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
t0 = time.perf_counter()
for i in range(N):
    ctx.set_air(Fa, g[rt + "air_rest8"], g[tag + "air_localVI2Global"], len(g[tag + "air_bnd"]), g[rt + "air_fixed"], float(g[rt + "w_scaf"]) / Fa.shape[0])
    ctx.set_uv(g[tag + "V"], g[tag + "air_V"])
    ctx.set_pattern_from_elements()
    r = ctx.newton_step(p0, float(g[rt + "targetGRes"]))
    uv = ctx.get_uv()
ctx.synchronize()
print("wall per e2e step: %.3f ms  (pcg iters %d)" % (1e3 * (time.perf_counter() - t0) / N, r["pcg_iters"]))
ctx.close()
