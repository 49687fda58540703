"""CPU: consistency of the fixtures recorded from the UNMODIFIED reference that the GPU tests lean on.

* the two branches of configs[1] (traces/bimba_cfg2_trace.txt, bimba_cfg2_alt_trace.txt: the reference on its input perturbed by a
  relative 1e-9, tools/ref_sensitivity.py): the same 77 connectivity stages, then two more splits in the second branch;
* the reference's own direction error per sweep state (sweep_*_ref_direction_error.json) covers every state of the sweeps;
* the reference's results on the 71 benchmark meshes to convergence: all finished, every run ends below the distortion bound."""
import json
import os

import numpy as np
from conftest import GOLDEN


def _stages(path):
    out = []
    for ln in open(path):
        if not ln.strip():
            continue
        kv = dict(x.split("=") for x in ln.split())
        key = (kv["Fhash"], kv["cohEhash"])
        if not out or out[-1] != key:
            out.append(key)
    return out


def test_reference_branches_of_configs1_share_their_first_77_stages():
    main = _stages(os.path.join(GOLDEN, "traces", "bimba_cfg2_trace.txt"))
    alt = _stages(os.path.join(GOLDEN, "traces", "bimba_cfg2_alt_trace.txt"))
    assert len(main) == 77 and len(alt) == 79 and alt[:77] == main
    fin = lambda n: [float(v) for v in open(os.path.join(GOLDEN, "traces", n)).read().split("\n")[3].split()]
    a, b = fin("bimba_cfg2_info.txt"), fin("bimba_cfg2_alt_info.txt")
    assert abs(a[0] - 4.28086) < 1e-5 and abs(a[1] - 2.65232) < 1e-5
    assert b[0] < a[0] and b[1] > a[1]                  # two more splits: less distortion, longer seams


def test_reference_direction_error_covers_every_sweep_state():
    for name in ("torus_cfg1", "bimba_cfg2", "bimba_cfg1"):
        g = np.load(os.path.join(GOLDEN, "sweep_%s.npz" % name))
        err = json.load(open(os.path.join(GOLDEN, "sweep_%s_ref_direction_error.json" % name)))
        for k in g["iters"]:
            e = err[str(int(k))]
            assert 0.0 <= e["ref_direction_rel_error"] < 1e-3 and e["refined_residual"] < 1e-14
        # the reference is NOT exact on the distorted early states: that is what the sweep's tolerance scales with
        worst = max(v["ref_direction_rel_error"] for v in err.values())
        assert worst > (1e-11 if name == "torus_cfg1" else 1e-8), (name, worst)


def test_reference_batch71_to_convergence_fixture():
    from optcuts_b200 import batch
    d = json.load(open(os.path.join(GOLDEN, "benchmark71_reference_to_convergence.json")))
    items = {n: f for n, f, _ in batch.benchmark71()}
    assert d["kind"] == "ref" and set(d["meshes"]) == set(items)
    for name, r in d["meshes"].items():
        assert r["rc"] == 0 and r["faces"] == items[name] and r["newton_iters"] > 0
        assert r["E_SD"] <= 4.1 + 1e-9, (name, r["E_SD"])          # distortion bound b_d = 4.1 of the batch's command line
