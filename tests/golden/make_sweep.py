"""Regenerates tests/golden/sweep_<run>.npz FROM THE REFERENCE ITSELF (build container only: needs oracle/_ref).

    python tests/golden/make_sweep.py

Teacher-forced sweep fixtures (SURVEY H2 / 7.3 "Step"): the full state of the reference run BEFORE Newton iteration k + 1
(= after iteration k: mesh, UVs, the air mesh the reference triangulated from them) and the UVs AFTER it, for ~20
iterations spread over the whole run of BASELINE.json configs[1] (fixed lambda) and ~12 of configs[0] (dual update:
lambda, hence energyParam0, changes along the run).  tests/test_gpu_sweep.py uploads state k, runs ONE Newton iteration on
the GPU and compares with what the reference computed (energies of the trace at 1e-9, UVs).  Meshes that did not change
between two picked iterations are stored once.
"""
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import state_io  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(ROOT, "oracle", "_ref", "OptCuts_probe")
INPUTS = os.path.join(HERE, "inputs")
RUNS = {
    "bimba_cfg2": ("bimba_i_f10000.obj", ["0.025", "1", "2", "4.1", "1", "0"], 20),
    "bimba_cfg1": ("bimba_i_f10000.obj", ["0.999", "1", "0", "4.1", "1", "0"], 20),
    # highly distorted start after cut_to_disk: the step bound comes from air triangles whose rows are 1e6 times softer
    # than the mesh's (the case that exposed the plain 2-norm PCG stopping test)
    "torus_cfg1": ("torus.obj", ["0.999", "1", "0", "4.1", "1", "0"], 8),
}


def parse_trace(path):
    return [dict(kv.split("=") for kv in ln.split()) for ln in open(path) if ln.strip()]


def pick(trace, n):
    """iterations k such that k -> k + 1 is a plain Newton iteration (same connectivity, not converged at k + 1's start)"""
    ok = [i for i in range(len(trace) - 1)
          if trace[i]["Fhash"] == trace[i + 1]["Fhash"] and trace[i]["cohEhash"] == trace[i + 1]["cohEhash"]
          and trace[i]["topo"] == trace[i + 1]["topo"] and trace[i + 1]["conv"] == "0" and trace[i]["p0"] == trace[i + 1]["p0"]]
    sel = [ok[int(round(j * (len(ok) - 1) / (n - 1)))] for j in range(n)]
    return sorted(set(sel))


def run(name):
    mesh, args, n = RUNS[name]
    trace = parse_trace(os.path.join(HERE, "traces", name + "_trace.txt"))
    rows = pick(trace, n)                                     # 0-based rows of the trace; row r is iteration r + 1
    its = sorted({int(trace[r]["it"]) for r in rows} | {int(trace[r + 1]["it"]) for r in rows})
    with tempfile.TemporaryDirectory() as wd:
        for f in os.listdir(INPUTS):
            shutil.copy(os.path.join(INPUTS, f), wd)
        os.makedirs(os.path.join(wd, "dumps"))
        env = dict(os.environ, ORACLE_DUMP_DIR=os.path.join(wd, "dumps"), ORACLE_DUMP_ITERS=",".join(map(str, its)))
        subprocess.run([PROBE, "100", os.path.join(wd, mesh)] + args + ["golden"], cwd=wd, env=env, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        st = {i: state_io.read_state(os.path.join(wd, "dumps", "state_%06d.bin" % i)) for i in its}
    out, meshes = {}, {}
    for r in rows:
        k, k1 = int(trace[r]["it"]), int(trace[r + 1]["it"])
        s, s1 = st[k], st[k1]
        key = hashlib.sha1(s["F"].tobytes() + s["V_rest"].tobytes()).hexdigest()[:12]
        if key not in meshes:
            meshes[key] = len(meshes)
            m = "m%d_" % meshes[key]
            out[m + "V_rest"], out[m + "F"], out[m + "fixedVert"] = s["V_rest"], s["F"], s["fixedVert"]
        p = "k%d_" % k
        out[p + "mesh"] = np.int32(meshes[key])
        for f in ("V", "air_V", "air_F", "air_bnd", "air_localVI2Global", "air_fixedVert", "air_scalars", "scalars"):
            out[p + f] = s[f]
        out[p + "V_next"], out[p + "scalars_next"] = s1["V"], s1["scalars"]
        out[p + "E_next"] = np.array([float(trace[r + 1]["E"]), float(trace[r + 1]["Enoscaf"])])
        out[p + "p0"] = np.float64(float(trace[r]["p0"]))
    out["iters"] = np.array([int(trace[r]["it"]) for r in rows], np.int32)
    out["lambda_init"] = np.float64(float(args[0]))
    out["nV0"] = np.int32(st[its[0]]["V"].shape[0] if False else 0)
    path = os.path.join(HERE, "sweep_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(name, "iterations", list(out["iters"]), "meshes", len(meshes), "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    for nme in (sys.argv[1:] or list(RUNS)):
        run(nme)
