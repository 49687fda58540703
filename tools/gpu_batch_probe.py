"""PCG iteration counts / time per Newton step on synthetic disks of the batch71 workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import optcuts_b200 as ob
from optcuts_b200 import batch
for faces in (500, 2000, 8000, 20000):
    V_rest, F, UV = batch.synthetic_disk(faces, seed=1)
    ctx = ob.Context(0)
    mesh = ob.TriMesh(V_rest, F, UV, ctx=ctx)
    opt = ob.Optimizer(mesh, energyParams=(0.975,), ctx=ctx)
    opt.precompute()
    its, ts = [], []
    for _ in range(8):
        t = time.perf_counter()
        if opt.solve(1): break
        ctx.synchronize(); ts.append(1e3 * (time.perf_counter() - t)); its.append(opt.last_step["pcg_iters"])
    print("faces", faces, "verts", UV.shape[0], "precond", ctx.precond_info(), "pcg iters", its, "ms/step", [round(x, 2) for x in ts])
    ctx.close()
