"""GPU: kernel (6) for real -- the nested (bijective) optimizers of the topology step's candidate evaluation, run in lock
step on the device inside the reference's own TriMesh::querySplit / queryMerge (shim/CudaCandidates.cpp), against the
reference's nested Optimizer on the SAME local problems in the same process (OCB_CANDIDATES_VERIFY): for every query of a
whole run (all boundary-split, interior-split, merge and propagation queries: >= 10 topology steps) every candidate's final
local E_SD, its new vertex positions and its Newton iteration count are compared; the decisions themselves (arg-max, op
type, path) are covered by the connectivity hashes of tests/test_gpu_runs.py."""
import os
import shutil
import subprocess

import pytest
from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
CUDA_PROBE = os.path.join(ROOT, "shim", "_build", "OptCuts_cuda_probe")


def _run(tmp_path, mesh, args, max_iters=None):
    if not os.path.exists(CUDA_PROBE):
        pytest.skip("shim/_build/OptCuts_cuda_probe not built")
    for f in os.listdir(os.path.join(GOLDEN, "inputs")):
        shutil.copy(os.path.join(GOLDEN, "inputs", f), tmp_path)
    env = dict(os.environ, ORACLE_TRACE=str(tmp_path / "trace.txt"), OCB_CANDIDATES_VERIFY=str(tmp_path / "verify.txt"), OCB_HOST_TIMING="1")
    if max_iters:
        env["ORACLE_MAX_ITERS"] = str(max_iters)
    r = subprocess.run([CUDA_PROBE, "100", str(tmp_path / mesh)] + args + ["t"], cwd=tmp_path, env=env, capture_output=True, text=True, errors="replace", timeout=3000)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [dict(kv.split("=") for kv in ln.split()) for ln in open(tmp_path / "verify.txt")]
    stats = [ln for ln in r.stderr.split("\n") if "ocb candidates" in ln]
    return rows, stats


@pytest.mark.parametrize("mesh,args,min_queries", [("bimba_i_f10000.obj", ["0.025", "1", "2", "4.1", "1", "0"], 40),
                                                   ("torus.obj", ["0.999", "1", "0", "4.1", "1", "0"], 1)])
def test_device_candidates_match_reference_nested_optimizers(tmp_path, mesh, args, min_queries):
    rows, stats = _run(tmp_path, mesh, args)
    assert len(rows) >= min_queries
    n = sum(int(r["problems"]) for r in rows)
    on_dev = sum(int(r["on_device"]) for r in rows)
    bij = sum(int(r["bijective"]) for r in rows)
    worstE = max(float(r["worst_rel_Esd"]) for r in rows)
    worstV = max(float(r["worst_rel_UV"]) for r in rows)
    itdiff = sum(int(r["iter_count_differs"]) for r in rows)
    print("%s: %d queries, %d local problems (%d bijective, %d on the device), worst rel E_SD %.2e, worst rel UV %.2e, iteration count differs in %d; %s"
          % (mesh, len(rows), n, bij, on_dev, worstE, worstV, itdiff, stats))
    assert on_dev == n                                   # nothing fell back to the CPU (limits: 128 vertices, 192 triangles, 32 free)
    assert bij > 0.5 * n                                 # bijectivity is on: the air-mesh path is what is being tested
    # a nested solve stops at relGL2Tol 1e-6 / the relative-decrease test: two correct solvers agree to ~1e-9 in E unless one
    # takes an extra iteration at the threshold
    assert worstE <= 1e-7 and worstV <= 1e-5
    assert itdiff <= 0.02 * n + 1
