"""The drop-in boundary inside the REAL host program: the reference's own main.cpp / Optimizer.cpp / TriMesh.cpp /
Scaffold.cpp (compiled from /root/reference by shim/Makefile, travelling as a prebuilt) with
  * CudaLinSysSolver   in place of EigenLibSolver  (through the reference's compile-time solver switch) and
  * CudaSymDirichletEnergy registered in place of SymDirichletEnergy (main.cpp:1570),
both thin C++ subclasses over the C-ABI, run on the GPU and compared with the reference's trace."""
import os
import subprocess

import numpy as np
import pytest
from conftest import GOLDEN, ROOT
from objfixture import write_obj

pytestmark = pytest.mark.gpu
CUDA_PROBE = os.path.join(ROOT, "shim", "_build", "OptCuts_cuda_probe")
CUDA_BIN = os.path.join(ROOT, "shim", "_build", "OptCuts_cuda")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "OptCuts_bin")
ARGS = ["0.025", "1", "2", "4.1", "1", "0", "t"]          # BASELINE.json configs[1]


def parse_trace(path):
    return [dict(kv.split("=") for kv in ln.split()) for ln in open(path) if ln.strip()]


@pytest.fixture(scope="module")
def obj(tmp_path_factory, golden):
    d = tmp_path_factory.mktemp("dropin")
    p = str(d / "bimba_s1.obj")
    write_obj(p, golden["s1_V_rest"], golden["s1_F"], golden["s1_V"])
    return p


@pytest.mark.parametrize("device_newton", ["1", "0"], ids=["optimizer-hooks", "plugins-call-by-call"])
def test_newton_iterations_inside_reference_host(obj, tmp_path, device_newton):
    """device_newton=1: the Optimizer hooks keep the whole Newton iteration on the device (shim/CudaOptimizer.cpp);
    device_newton=0: only the Energy / LinSysSolver virtuals are served by the GPU, call by call through Eigen triplets."""
    if not os.path.exists(CUDA_PROBE):
        pytest.skip("shim/_build/OptCuts_cuda_probe not built (make -C shim needs the reference headers)")
    n = 10
    env = dict(os.environ, ORACLE_TRACE=str(tmp_path / "trace.txt"), ORACLE_MAX_ITERS=str(n), OCB_DEVICE_NEWTON=device_newton)
    r = subprocess.run([CUDA_PROBE, "100", obj] + ARGS, cwd=tmp_path, env=env, capture_output=True, text=True, errors="replace", timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = parse_trace(tmp_path / "trace.txt")
    # the fixture is the reference's state after iteration 1: iteration k here == iteration k+1 of the golden trace
    # (the unmodified reference started from the same OBJ reproduces that trace bit for bit)
    want = parse_trace(os.path.join(GOLDEN, "bimba_cfg2_trace.txt"))[1:]
    assert len(got) == n
    for k in range(n):
        # free-running (not teacher-forced): the ~1e-11 PCG-vs-LDLT difference of every step is amplified ~6x per
        # iteration on the steep Tutte-start landscape (see test_optimizer_free_run_with_reference_scaffold)
        tol = 1e-9 if k < 2 else (1e-8 if k < 3 else (1e-6 if k < 5 else 1e-4))
        for key in ("E", "Enoscaf"):
            assert abs(float(got[k][key]) - float(want[k][key])) <= tol * float(want[k][key]), (k, key, got[k], want[k])
        for key in ("F", "V", "amF", "amV", "bnd", "cohE"):
            assert got[k][key] == want[k][key], (k, key)


@pytest.mark.slow
def test_whole_run_matches_reference(obj, tmp_path):
    """configs[1] to convergence (geometry + topology steps, ~170 Newton iterations) with the CUDA plugins vs the
    unmodified reference, both started from the same OBJ: same number of topology steps, final E_SD / E_se equal
    to the 6 digits info.txt carries (north_star: within 1e-6), Newton iteration count within 5 % (the free-running
    trajectories separate at the 1e-6 level after ~8 iterations, so the count is not expected to be identical)."""
    if not (os.path.exists(CUDA_BIN) and os.path.exists(REF_BIN)):
        pytest.skip("prebuilt binaries missing")
    out = {}
    for name, exe in (("cuda", CUDA_BIN), ("ref", REF_BIN)):
        d = tmp_path / name
        d.mkdir()
        r = subprocess.run([exe, "100", obj] + ARGS, cwd=d, capture_output=True, text=True, errors="replace", timeout=1500)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        folder = os.listdir(d / "output")[0]
        info = open(d / "output" / folder / "info.txt").read().split("\n")
        out[name] = dict(iters=[int(v) for v in info[1].split()[:2]], E=[float(v) for v in info[3].split()],
                         timers=info[2])
    print("reference:", out["ref"], "\ncuda:", out["cuda"])
    assert out["cuda"]["iters"][1] == out["ref"]["iters"][1]
    assert abs(out["cuda"]["iters"][0] - out["ref"]["iters"][0]) <= 0.05 * out["ref"]["iters"][0]
    for a, b in zip(out["cuda"]["E"], out["ref"]["E"]):
        assert abs(a - b) <= 2e-6 * abs(b)


def test_linsys_solver_virtual_surface():
    """Every virtual of OptCuts::LinSysSolver on shim/CudaLinSysSolver.hpp (C++ program shim/test_linsys_surface.cpp): the
    Optimizer's set_pattern(vNeighbor) + update_a + factorize + solve sequence, set_pattern(SparseMatrix), multiply, coeffMtr and
    getNumNonzeros against Eigen::SimplicialLDLT on the same matrix (solution 1e-9, products and entries 1e-12)."""
    exe = os.path.join(ROOT, "shim", "_build", "linsys_surface_test")
    if not os.path.exists(exe):
        pytest.skip("shim/_build/linsys_surface_test not built (make -C shim needs the reference headers)")
    r = subprocess.run([exe], capture_output=True, text=True, errors="replace", timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "linsys surface: ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
