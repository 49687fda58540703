"""Regenerates tests/golden/traces/*.txt FROM THE REFERENCE ITSELF (build container only: needs oracle/_ref).

    python tests/golden/make_traces.py [name ...]

Runs oracle/_ref/OptCuts_probe (= unmodified reference + the %.17g per-iteration hook of oracle/probe_hook.hpp) on the
input fixtures of tests/golden/inputs/ with the reference's own command lines (BASELINE.json configs; batch.py:11-14):
one line per Newton iteration with E, E without scaffold, E_se, lambda, the element / vertex / seam counts and an FNV-1a
hash of F and cohE (the connectivity after every topology operation, i.e. type AND path of the whole op sequence).
The drop-in host program (shim/_build/OptCuts_cuda_probe) must reproduce these traces on the GPU (tests/test_gpu_runs.py).
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(ROOT, "oracle", "_ref", "OptCuts_probe")
INPUTS = os.path.join(HERE, "inputs")
# name: (mesh file, arguments after the mesh path = lambda_init testID methodType distortionBound bijectivity initCut)
RUNS = {
    "bimba_cfg1": ("bimba_i_f10000.obj", ["0.999", "1", "0", "4.1", "1", "0"]),      # BASELINE configs[0]: dual update on
    "bimba_cfg2": ("bimba_i_f10000.obj", ["0.025", "1", "2", "4.1", "1", "0"]),      # configs[1]: fixed lambda
    "torus_cfg1": ("torus.obj", ["0.999", "1", "0", "4.1", "1", "0"]),               # configs[4]: highGenus (cut_to_disk initial seams)
    "face_rsp_cfg1": ("face_f10000.obj", ["0.999", "1", "0", "4.1", "1", "0"]),      # configs[4]: regional seam placement (vertWeight)
    "lucy6k_cfg1": ("lucy_o_f6032.obj", ["0.999", "1", "0", "4.1", "1", "0"]),       # configs[2]: scalability set, smallest member
}


def run(name):
    mesh, args = RUNS[name]
    out_dir = os.path.join(HERE, "traces")
    os.makedirs(out_dir, exist_ok=True)
    with tempfile.TemporaryDirectory() as wd:
        # the reference looks for <mesh>_selected.txt next to the mesh (main.cpp:1585-1600)
        for f in os.listdir(INPUTS):
            shutil.copy(os.path.join(INPUTS, f), wd)
        env = dict(os.environ, ORACLE_TRACE=os.path.join(wd, "trace.txt"))
        t0 = time.time()
        subprocess.run([PROBE, "100", os.path.join(wd, mesh)] + args + ["golden"], cwd=wd, env=env, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.time() - t0
        folder = os.listdir(os.path.join(wd, "output"))[0]
        shutil.copy(os.path.join(wd, "trace.txt"), os.path.join(out_dir, name + "_trace.txt"))
        shutil.copy(os.path.join(wd, "output", folder, "info.txt"), os.path.join(out_dir, name + "_info.txt"))
    print("%s: %.1f s, %d iterations" % (name, dt, sum(1 for _ in open(os.path.join(out_dir, name + "_trace.txt")))))


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(RUNS)):
        run(n)
