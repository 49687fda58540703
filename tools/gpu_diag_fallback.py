"""Diagnostic (GPU box): does the two-level preconditioner survive on the badly scaled sweep states, or does the solve fall back to
block-Jacobi?   python tools/gpu_diag_fallback.py bimba_cfg2 3,7,10,14 [bimba_cfg1 8,57,86]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import optcuts_b200 as ob  # noqa: E402
from gpu_diag_sweep_state import upload  # noqa: E402


def main():
    ctx = ob.Context(0)
    args = sys.argv[1:] or ["bimba_cfg2", "3,7,10,14"]
    for name, ks in zip(args[0::2], args[1::2]):
        g = np.load(os.path.join(ROOT, "tests", "golden", "sweep_%s.npz" % name))
        for k in [int(v) for v in ks.split(",")]:
            p0, nV = upload(ctx, g, k)
            gr, _ = ctx.gradient(p0)
            ctx.set_pattern_from_elements(); ctx.hessian_assemble(p0)
            for eq in (0, 1):
                ctx.set_option("mas_equilibrate", eq)
                f0 = ctx.precond_info()["fallbacks"]
                ctx.factorize()
                x, info = ctx.solve(None, 1e-12, 0, allow_not_converged=True)
                pi = ctx.precond_info()
                print("%s it %d equilibrate=%d: %d CG iterations, status %d, rel res %.1e, preconditioner fallbacks in this solve: %d, levels %s"
                      % (name, k, eq, info["iters"], info["status"], info["rel_res"], pi["fallbacks"] - f0, pi["nodes"]))
                sys.stdout.flush()
            ctx.set_option("mas_equilibrate", 0)
    ctx.close()


if __name__ == "__main__":
    main()
