"""Regenerates the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Needs /root/reference and oracle/_ref (make -C oracle ref); runs only in the build container:

    python tests/golden/make_golden.py            # config 2 states + traces of configs 1 and 2

What it does
  1. runs oracle/_ref/OptCuts_probe (= unmodified reference + %.17g dump hook) on
     input/bimba_i_f10000.obj with BASELINE.json's configs[1] arguments (0.025 1 2 4.1 1 0) and
     configs[0] arguments (0.999 1 0 4.1 1 0), keeping the per-iteration trace;
  2. keeps the full state after Newton iterations 1,2 (Tutte start, no seams) and 100,101 (seams);
  3. re-plays ONE reference Newton iteration from states 1 and 100 through the real
     OptCuts::Optimizer (oracle/ref_capi.cpp) — verifying that this reproduces states 2 / 101 bit for
     bit — and records what the reference computed on the way: energies, gradient, CSR matrix,
     search direction, step size.
"""
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refapi, state_io  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
PROBE = os.path.join(ROOT, "oracle", "_ref", "OptCuts_probe")
MESH = os.path.join(REF, "input", "bimba_i_f10000.obj")


def run_probe(args, dump_iters, workdir):
    os.makedirs(os.path.join(workdir, "dumps"), exist_ok=True)
    env = dict(os.environ, ORACLE_TRACE=os.path.join(workdir, "trace.txt"),
               ORACLE_DUMP_DIR=os.path.join(workdir, "dumps"), ORACLE_DUMP_ITERS=",".join(map(str, dump_iters)))
    subprocess.run([PROBE, "100", MESH] + args + ["golden"], cwd=workdir, env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    info = [f for f in os.listdir(os.path.join(workdir, "output"))][0]
    with open(os.path.join(workdir, "output", info, "info.txt")) as f:
        info_txt = f.read()
    with open(os.path.join(workdir, "trace.txt")) as f:
        trace = f.read()
    return trace, info_txt


def replay(state, nxt, p0):
    """one reference Newton iteration from `state`; returns the recorded intermediates"""
    m = refapi.RefMesh(state["V_rest"], state["F"], state["V"], cohE=state["cohE"] if len(state["cohE"]) else None)
    opt = refapi.RefOptimizer(m, p0, scaffolding=True, mute=True)
    sz0, sc0 = opt.sizes(), opt.scalars()
    air0 = opt.air()
    assert np.array_equal(air0["V"], state["air_V"]) and np.array_equal(air0["F"], state["air_F"])
    g = opt.recompute_gradient()
    rest8, msc = m.features()
    opt.solve(1)
    # the sparse system of this iteration is still in the solver; searchDir has the old size
    import ctypes as C
    p = np.zeros(sz0["nSys"])
    refapi.lib().ref_opt_get_search_dir(opt.h, p.ctypes.data_as(C.POINTER(C.c_double)))
    ia = np.zeros(sz0["nSys"] + 1, np.int32); ja = np.zeros(sz0["nnz"], np.int32); a = np.zeros(sz0["nnz"])
    # pattern/values were rebuilt inside solve(1) for the SAME scaffold (Optimizer.cpp:522-530) -> same sizes
    refapi.lib().ref_opt_get_csr(opt.h, ia.ctypes.data_as(C.POINTER(C.c_int32)), ja.ctypes.data_as(C.POINTER(C.c_int32)),
                                 a.ctypes.data_as(C.POINTER(C.c_double)))
    sc1 = opt.scalars()
    uv1 = opt.uv()
    assert np.array_equal(uv1, nxt["V"]), "replay through ref_capi does not reproduce the reference run"
    alpha = float(np.median((uv1 - state["V"]).ravel()[np.abs(p[:uv1.size]).argsort()[-50:] * 0 + np.abs(p[:2 * uv1.shape[0]]).argsort()[-50:] // 1 % 1])) if False else None
    # accepted step: x1 = x0 + alpha p  -> alpha from the largest component
    k = int(np.argmax(np.abs(p[:2 * state["V"].shape[0]])))
    alpha = (uv1[k // 2, k % 2] - state["V"][k // 2, k % 2]) / p[k]
    out = dict(E_last=sc0["lastEnergyVal"], E_scaf_last=sc0["energyVal_scaffold"], E_sd_last=sc0["energyVal_ET0"],
               targetGRes=sc0["targetGRes"], w_scaf=sc0["w_scaf"], gradient=g, sqn_g=float(g @ g),
               searchDir=p, ia=ia, ja=ja, a=a, alpha=alpha,
               E_new=sc1["lastEnergyVal"], E_scaf_new=sc1["energyVal_scaffold"], E_sd_new=sc1["energyVal_ET0"],
               lastEDec=sc1["lastEDec"], rest8=rest8, surfaceArea=msc["surfaceArea"], avgEdgeLen=msc["avgEdgeLen"],
               virtualRadius=msc["virtualRadius"], air_rest8=air0["rest8"], air_fixed=air0["fixed"],
               energy_per_elem=m.energy_per_elem(), divgrad=m.divgrad(),
               seam_sparsity=m.seam_sparsity(False), seam_sparsity_soup=m.seam_sparsity(True))
    if len(state["cohE"]):
        b, e = m.coh_features()
        out.update(boundaryEdge=b, edgeLen=e)
    opt.close(); m.close()
    return out


def pack_state(prefix, s, full):
    d = {}
    keys = ["V", "air_V", "air_F", "air_bnd", "air_localVI2Global", "air_fixedVert", "air_scalars", "scalars"]
    if full:
        keys += ["V_rest", "F", "cohE", "fixedVert"]
    for k in keys:
        d[prefix + k] = s[k]
    return d


def main():
    out = {}
    with tempfile.TemporaryDirectory() as wd:
        trace2, info2 = run_probe(["0.025", "1", "2", "4.1", "1", "0"], [1, 2, 100, 101], wd)
        st = {i: state_io.read_state(os.path.join(wd, "dumps", "state_%06d.bin" % i)) for i in (1, 2, 100, 101)}
    open(os.path.join(HERE, "bimba_cfg2_trace.txt"), "w").write(trace2)
    open(os.path.join(HERE, "bimba_cfg2_info.txt"), "w").write(info2)
    p0 = 1.0 - 0.025
    out.update(pack_state("s1_", st[1], True)); out.update(pack_state("s2_", st[2], False))
    out.update(pack_state("s100_", st[100], True)); out.update(pack_state("s101_", st[101], False))
    assert st[100]["V"].shape == st[101]["V"].shape, "pick another pair: topology changed between 100 and 101"
    for tag, a, b in (("r1_", 1, 2), ("r100_", 100, 101)):
        r = replay(st[a], st[b], p0)
        if tag == "r100_":      # keep the file small: the full matrix is stored for state 1 only
            r["a_sum"] = float(np.sum(r["a"])); r["a_abs_sum"] = float(np.sum(np.abs(r["a"])))
            for k in ("ia", "ja", "a", "energy_per_elem"):
                r.pop(k)
        out.update({tag + k: v for k, v in r.items()})
    out["energyParam0"] = p0
    np.savez_compressed(os.path.join(HERE, "bimba_cfg2_states.npz"), **out)
    print("wrote bimba_cfg2_states.npz", os.path.getsize(os.path.join(HERE, "bimba_cfg2_states.npz")) / 1e6, "MB")
    if "--skip-cfg1" not in sys.argv:
        with tempfile.TemporaryDirectory() as wd:
            trace1, info1 = run_probe(["0.999", "1", "0", "4.1", "1", "0"], [], wd)
        open(os.path.join(HERE, "bimba_cfg1_trace.txt"), "w").write(trace1)
        open(os.path.join(HERE, "bimba_cfg1_info.txt"), "w").write(info1)


if __name__ == "__main__":
    main()
