"""Diagnostic: whole runs of the host program (shim/_build/OptCuts_cuda_probe) under library / option variants, each compared
with the recorded reference trace: where the free run leaves the reference (first iteration whose E differs by > 1e-9, first
differing connectivity stage), iteration counts, finals.

    python tools/gpu_diag_run.py name variant[,variant...]     variant = label:LIBDIR-or-"-":ENV=V;ENV=V
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
INPUTS = os.path.join(GOLDEN, "inputs")
PROBE = os.path.join(ROOT, "shim", "_build", "OptCuts_cuda_probe")
RUNS = {
    "torus_cfg1": ("torus.obj", ["0.999", "1", "0", "4.1", "1", "0"]),
    "bimba_cfg2": ("bimba_i_f10000.obj", ["0.025", "1", "2", "4.1", "1", "0"]),
    "bimba_cfg1": ("bimba_i_f10000.obj", ["0.999", "1", "0", "4.1", "1", "0"]),
    "face_rsp_cfg1": ("face_f10000.obj", ["0.999", "1", "0", "4.1", "1", "0"]),
    "lucy6k_cfg1": ("lucy_o_f6032.obj", ["0.999", "1", "0", "4.1", "1", "0"]),
}


def parse_trace(path):
    return [dict(kv.split("=") for kv in ln.split()) for ln in open(path) if ln.strip()]


def stages(trace):
    out = []
    for ln in trace:
        key = (ln["Fhash"], ln["cohEhash"])
        if not out or out[-1][0] != key:
            out.append((key, []))
        out[-1][1].append(ln)
    return out


def main():
    name = sys.argv[1]
    mesh, args = RUNS[name]
    want = parse_trace(os.path.join(GOLDEN, "traces", name + "_trace.txt"))
    sw = stages(want)
    for var in sys.argv[2].split(","):
        label, libdir, envs = (var.split(":") + ["", ""])[:3]
        env = dict(os.environ)
        if libdir and libdir != "-":
            env["LD_LIBRARY_PATH"] = os.path.join(ROOT, libdir) + ":" + env.get("LD_LIBRARY_PATH", "")
        for kv in envs.split(";"):
            if kv:
                k, v = kv.split("=")
                env[k] = v
        with tempfile.TemporaryDirectory() as wd:
            for f in os.listdir(INPUTS):
                shutil.copy(os.path.join(INPUTS, f), wd)
            env["ORACLE_TRACE"] = os.path.join(wd, "trace.txt")
            r = subprocess.run([PROBE, "100", os.path.join(wd, mesh)] + args + ["t"], cwd=wd, env=env, capture_output=True, text=True, errors="replace")
            if r.returncode != 0:
                print(label, "rc", r.returncode, r.stderr[-500:])
                continue
            got = parse_trace(env["ORACLE_TRACE"])
            out = os.path.join(wd, "output")
            info = open(os.path.join(out, os.listdir(out)[0], "info.txt")).read().split("\n")
        sg = stages(got)
        lead = 0
        for x, y in zip(got, want):
            if x["Fhash"] != y["Fhash"] or abs(float(x["Enoscaf"]) - float(y["Enoscaf"])) > 1e-9 * abs(float(y["Enoscaf"])):
                break
            lead += 1
        first = next((k for k in range(min(len(sg), len(sw))) if sg[k][0] != sw[k][0]), None)
        print("%-22s its %d (ref %d) stages %d (ref %d) first differing stage %s, %d leading iterations within 1e-9; info %s | finals %s"
              % (label, len(got), len(want), len(sg), len(sw), first, lead, info[1], info[3]))
        if len(sg) != len(sw) or first is not None:
            k0 = first if first is not None else min(len(sg), len(sw)) - 1
            for tag, s in (("got", sg), ("ref", sw)):
                for k in range(max(0, k0 - 1), min(len(s), k0 + 3)):
                    print("    %s stage %d: %s" % (tag, k, ["it=%s conv=%s topo=%s E=%s" % (l["it"], l["conv"], l["topo"], l["Enoscaf"]) for l in s[k][1]]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
