"""The reference's 71 benchmark meshes TO CONVERGENCE through a host program, one process per mesh with the reference's own
command line (batch.py:11-14, headless mode 100, lambda_init 0.999, OptCuts, b_d 4.1, bijective), several processes at a time.

    python tools/batch_to_convergence.py ref  out.json [procs] [timeout_s] [first_n]     the unmodified reference (oracle/_ref/OptCuts_bin); first_n < 0: a sample spread over the sizes
    python tools/batch_to_convergence.py cuda out.json [procs] [timeout_s] [first_n]     the host program with the GPU plugins (shim/_build/OptCuts_cuda)
    python tools/batch_to_convergence.py compare ref.json cuda.json                      per-mesh table + summary

Per mesh: process wall clock, the reference's own info.txt (Newton iterations, topology steps, its timers, final E_SD and
E_se).  s/mesh of BASELINE.json's metric = the wall clock here (process start, CUDA context creation and file I/O inside)."""
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optcuts_b200 import batch  # noqa: E402

EXE = {"ref": os.path.join(ROOT, "oracle", "_ref", "OptCuts_bin"), "cuda": os.path.join(ROOT, "shim", "_build", "OptCuts_cuda")}


def run_one(exe, mesh_path, wd, timeout, extra_env=None):
    os.makedirs(wd, exist_ok=True)
    t0 = time.perf_counter()
    try:
        r = subprocess.run([exe, "100", mesh_path] + batch.MESH_ARGS + ["b"], cwd=wd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=timeout, text=True, errors="replace",
                           env=dict(os.environ, **(extra_env or {})))
        rc, err = r.returncode, r.stderr[-300:]
    except subprocess.TimeoutExpired:
        rc, err = -999, "timeout"
    dt = time.perf_counter() - t0
    row = {"rc": rc, "wall_s": dt}
    out = os.path.join(wd, "output")
    try:
        info = open(os.path.join(out, os.listdir(out)[0], "info.txt")).read().split("\n")
        row.update(newton_iters=int(info[1].split()[0]), topo_steps=int(info[1].split()[1]), lambda_final=float(info[1].split()[5]),
                   in_process_s=float(info[2].split()[2]), timers=" ".join(info[2].split()[4:]),
                   E_SD=float(info[3].split()[0]), E_se=float(info[3].split()[1]))
    except Exception:   # noqa: BLE001
        row["stderr_tail"] = err
    return row


def run(kind, out_path, procs, timeout, first_n):
    items = batch.benchmark71()
    order = sorted(range(len(items)), key=lambda i: -items[i][1])                     # largest first
    order = order[:first_n] if first_n > 0 else order[::max(1, len(order) // -first_n)][:-first_n]      # first_n < 0: that many meshes spread over the size range
    import contextlib
    with tempfile.TemporaryDirectory() as wd, (batch.MpsDaemon(0) if kind == "cuda" else contextlib.nullcontext()) as mps:
        paths = batch.extract_benchmark(os.path.join(wd, "in"))
        cenv = mps.child_env() if kind == "cuda" else None
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=procs) as ex:
            rows = list(ex.map(lambda i: run_one(EXE[kind], paths[items[i][0]], os.path.join(wd, "m%d" % i), timeout, cenv), order))
        wall = time.perf_counter() - t0
    res = {"kind": kind, "mps": bool(kind == "cuda" and mps.up), "procs": procs, "cores": os.cpu_count(), "timeout_s": timeout, "batch_wall_s": wall,
           "meshes": {items[i][0]: dict(rows[k], faces=items[i][1]) for k, i in enumerate(order)}}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    ok = [r for r in rows if r["rc"] == 0 and "E_SD" in r]
    print("%s: %d meshes, %d finished, batch wall %.1f s with %d processes at a time, sum of process walls %.1f s (%.2f s/mesh)"
          % (kind, len(rows), len(ok), wall, procs, sum(r["wall_s"] for r in ok), sum(r["wall_s"] for r in ok) / max(1, len(ok))))


def compare(ref_path, cuda_path):
    A, B = json.load(open(ref_path)), json.load(open(cuda_path))
    print("reference: %d processes at a time on %d cores, batch wall %.1f s;  GPU host program: %d at a time, batch wall %.1f s"
          % (A["procs"], A["cores"], A["batch_wall_s"], B["procs"], B["batch_wall_s"]))
    print("%-28s %6s | %7s %5s %8s %8s %8s | %7s %5s %8s %8s %8s | %s" % ("mesh", "faces", "ref it", "topo", "E_SD", "E_se", "wall s", "gpu it", "topo", "E_SD", "E_se", "wall s", "verdict"))
    same = close = differ = failed = 0
    sw_a = sw_b = sp_a = sp_b = 0.0
    for name in sorted(A["meshes"], key=lambda n: -A["meshes"][n]["faces"]):
        a, b = A["meshes"][name], B["meshes"].get(name)
        if b is None:
            continue
        if "E_SD" not in a or "E_SD" not in b:
            failed += 1
            print("%-28s %6d | %s | %s" % (name, a["faces"], "rc %d" % a["rc"] if "E_SD" not in a else "ok", "rc %d %s" % (b["rc"], b.get("stderr_tail", "")[-80:]) if "E_SD" not in b else "ok"))
            continue
        eS = abs(a["E_SD"] - b["E_SD"]) / abs(a["E_SD"])
        eE = abs(a["E_se"] - b["E_se"]) / max(abs(a["E_se"]), 1e-300)
        if eS <= 2e-5 and eE <= 2e-5 and a["topo_steps"] == b["topo_steps"]:
            verdict = "same result"; same += 1
        elif eS <= 2e-2 and eE <= 5e-2:
            verdict = "other branch, close (dE_SD %.1e, dE_se %.1e)" % (eS, eE); close += 1
        else:
            verdict = "other branch (dE_SD %.1e, dE_se %.1e)" % (eS, eE); differ += 1
        sw_a += a["wall_s"]; sw_b += b["wall_s"]; sp_a += a["in_process_s"]; sp_b += b["in_process_s"]
        print("%-28s %6d | %7d %5d %8.5f %8.5f %8.2f | %7d %5d %8.5f %8.5f %8.2f | %s"
              % (name, a["faces"], a["newton_iters"], a["topo_steps"], a["E_SD"], a["E_se"], a["wall_s"], b["newton_iters"], b["topo_steps"], b["E_SD"], b["E_se"], b["wall_s"], verdict))
    n = same + close + differ
    print("\n%d meshes finished in both arms: %d same result (finals within 2e-5, same number of topology steps), %d on another branch with close finals, "
          "%d on another branch; %d not finished in one arm" % (n, same, close, differ, failed))
    if n:
        print("s/mesh (process wall clock, mean over those %d): reference %.2f, GPU host program %.2f (%.2fx); in-process timer total: %.2f vs %.2f (%.2fx)"
              % (n, sw_a / n, sw_b / n, sw_a / sw_b, sp_a / n, sp_b / n, sp_a / sp_b))


if __name__ == "__main__":
    if sys.argv[1] == "compare":
        compare(sys.argv[2], sys.argv[3])
    else:
        run(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 4, float(sys.argv[4]) if len(sys.argv) > 4 else 900.0,
            int(sys.argv[5]) if len(sys.argv) > 5 else 71)
