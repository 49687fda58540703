"""How well defined is the reference's own op sequence?  Runs the UNMODIFIED reference (oracle/_ref/OptCuts_probe) on inputs whose
vertex coordinates are perturbed by a relative 1e-13 (a hundred ulps: far below anything a mesh file carries) and compares every
run with the recorded trace of the unperturbed input: Newton iterations, connectivity stages, first differing stage, finals.
A different linear solver perturbs every Newton iteration by kappa * eps >> 1e-13, so this is a LOWER bound of what swapping
the solver does to a free run.      python tools/ref_sensitivity.py [name] [n_seeds]   (build container, CPU only)"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_diag_run import GOLDEN, INPUTS, RUNS, parse_trace, stages  # noqa: E402

PROBE = os.path.join(ROOT, "oracle", "_ref", "OptCuts_probe")


def perturbed(src, dst, seed, rel):
    rng = np.random.default_rng(seed)
    with open(src) as f, open(dst, "w") as g:
        for ln in f:
            if ln.startswith("v "):
                xyz = np.array(ln.split()[1:4], float)
                xyz = xyz * (1.0 + rel * rng.uniform(-1, 1, 3))
                g.write("v %.17g %.17g %.17g\n" % tuple(xyz))
            else:
                g.write(ln)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "bimba_cfg2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rel = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-13
    save = int(sys.argv[4]) if len(sys.argv) > 4 else -1     # seed whose trace is kept as tests/golden/traces/<name>_alt_{trace,info}.txt
    seeds = range(n) if save < 0 else [save]
    mesh, args = RUNS[name]
    want = parse_trace(os.path.join(GOLDEN, "traces", name + "_trace.txt"))
    sw = stages(want)
    print("%s: reference trace %d Newton iterations, %d connectivity stages; perturbation %.0e relative" % (name, len(want), len(sw), rel))
    for seed in seeds:
        with tempfile.TemporaryDirectory() as wd:
            for f in os.listdir(INPUTS):
                shutil.copy(os.path.join(INPUTS, f), wd)
            if seed > 0:                      # seed 0 = the unperturbed input (must reproduce the trace bit for bit)
                perturbed(os.path.join(INPUTS, mesh), os.path.join(wd, mesh), seed, rel)
            env = dict(os.environ, ORACLE_TRACE=os.path.join(wd, "trace.txt"))
            subprocess.run([PROBE, "100", os.path.join(wd, mesh)] + args + ["golden"], cwd=wd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            got = parse_trace(env["ORACLE_TRACE"])
            out = os.path.join(wd, "output")
            info = open(os.path.join(out, os.listdir(out)[0], "info.txt")).read().split("\n")
            if seed == save:
                shutil.copy(env["ORACLE_TRACE"], os.path.join(GOLDEN, "traces", name + "_alt_trace.txt"))
                shutil.copy(os.path.join(out, os.listdir(out)[0], "info.txt"), os.path.join(GOLDEN, "traces", name + "_alt_info.txt"))
        sg = stages(got)
        lead = 0
        for x, y in zip(got, want):
            if x["Fhash"] != y["Fhash"] or abs(float(x["Enoscaf"]) - float(y["Enoscaf"])) > 1e-9 * abs(float(y["Enoscaf"])):
                break
            lead += 1
        first = next((k for k in range(min(len(sg), len(sw))) if sg[k][0] != sw[k][0]), None)
        print("seed %d: %d iterations, %d stages, first differing stage %s, %d leading iterations within 1e-9, info %s, finals %s"
              % (seed, len(got), len(sg), first, lead, info[1], info[3]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
