import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import optcuts_b200 as ob
from oracle import refapi as ref
from conftest import State, GOLDEN
g = np.load(os.path.join(GOLDEN, "bimba_cfg2_states.npz")); g = {k: g[k] for k in g.files}
s = State(g, "s1_", "r1_", "s2_")
trace = [dict(kv.split("=") for kv in ln.split()) for ln in open(os.path.join(GOLDEN, "bimba_cfg2_trace.txt"))]
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    c2 = ob.Context(0)
    mesh = ob.TriMesh(s.V_rest, s.F, s.UV, ctx=c2)
    def builder(m):
        rm = ref.RefMesh(m.V_rest, m.F, m.V); sc = ref.build_scaffold(rm); rm.close()
        return ob.scaffold.Scaffold(sc["V"], sc["F"], sc["bnd"], m.nV, fixedAir=sc["fixed"], rest8=sc["rest8"])
    opt = ob.Optimizer(mesh, energyParams=(s.p0,), scaffolding=True, scaffold_builder=builder, ctx=c2)
    opt.precompute()
    out = []
    for it in range(6):
        opt.solve(1)
        want = trace[it + 1]
        out.append("%.1e/%d" % (abs(opt.getLastEnergyVal() - float(want["E"])) / float(want["E"]), opt.last_step["pcg_iters"]))
    print("free-run rel E diff / pcg iters per iteration:", out, "fallbacks", c2.precond_info()["fallbacks"])
    c2.close()
