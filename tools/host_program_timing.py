"""Whole run of the reference HOST PROGRAM on one input, three ways, on this box's host cores (+ its GPU):
   cuda   shim/_build/OptCuts_cuda, Optimizer hooks: the Newton iteration is device resident (shim/CudaOptimizer.cpp)
   cuda0  same binary, OCB_DEVICE_NEWTON=0: only the Energy / LinSysSolver virtuals are served by the GPU, call by call
   ref    oracle/_ref/OptCuts_bin, the unmodified reference (Eigen SimplicialLDLT, TBB shim on all cores)
Prints the reference's own info.txt (line 2: Newton iterations, topology steps ...; line 3: its timers; line 4: final E_SD,
E_se) and the wall clock of each.   python tools/host_program_timing.py [mesh] [args...] > gpurun_out/..."""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")


def main():
    mesh = sys.argv[1] if len(sys.argv) > 1 else "bimba_i_f10000.obj"
    args = sys.argv[2:] if len(sys.argv) > 2 else ["0.025", "1", "2", "4.1", "1", "0"]
    runs = [("cuda", os.path.join(ROOT, "shim", "_build", "OptCuts_cuda"), {"OCB_DEVICE_NEWTON": "1", "OCB_CANDIDATES_REPORT": "1"}),
            ("cuda0", os.path.join(ROOT, "shim", "_build", "OptCuts_cuda"), {"OCB_DEVICE_NEWTON": "0"}),
            ("ref", os.path.join(ROOT, "oracle", "_ref", "OptCuts_bin"), {})]
    if os.environ.get("OCB_TIMING_SKIP"):
        runs = [r for r in runs if r[0] not in os.environ["OCB_TIMING_SKIP"].split(",")]
    print("== whole run: %s %s, %d host cores" % (mesh, " ".join(args), os.cpu_count()))
    res = {}
    for name, exe, env in runs:
        with tempfile.TemporaryDirectory() as wd:
            for f in os.listdir(INPUTS):
                shutil.copy(os.path.join(INPUTS, f), wd)
            t0 = time.perf_counter()
            r = subprocess.run([exe, "100", os.path.join(wd, mesh)] + args + ["t"], cwd=wd, env=dict(os.environ, **env), capture_output=True, text=True)
            dt = time.perf_counter() - t0
            out = os.path.join(wd, "output")
            info = open(os.path.join(out, os.listdir(out)[0], "info.txt")).read().split("\n") if os.path.isdir(out) else ["", "", "", ""]
        print("-- %-5s rc=%d wall %.2f s\n   iterations: %s\n   timers: %s\n   final E_SD, E_se: %s" % (name, r.returncode, dt, info[1], info[2], info[3]))
        if r.returncode != 0:
            print(r.stderr[-1500:])
        for ln in r.stderr.split("\n"):
            if ln.startswith("[ocb candidates]"):
                print("   " + ln)
        res[name] = dt
    if "ref" in res:
        for k in res:
            if k != "ref":
                print("speed-up %s vs ref (wall clock, whole run): %.2fx" % (k, res["ref"] / res[k]))


if __name__ == "__main__":
    main()
