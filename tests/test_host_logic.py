"""Host-side index logic of the Python mirror (adjacency, scaffold numbering, synthetic subdivision,
batch partitioning).  CPU only."""
import numpy as np
from test_oracle_vs_reference import random_mesh


def test_adjacency_matches_reference_sets(ref):
    from optcuts_b200.trimesh import adjacency_from_faces
    V_rest, F, UV = random_mesh(5, n=8)
    m = ref.RefMesh(V_rest, F, UV)
    ptr, idx = m.adjacency()                               # TriMesh::vNeighbor, TriMesh.cpp:444-453
    p2, i2 = adjacency_from_faces(F, len(UV))
    assert np.array_equal(ptr, p2) and np.array_equal(idx, i2)
    m.close()


def test_scaffold_numbering(state1):
    from optcuts_b200.scaffold import Scaffold
    a = state1.air
    s = Scaffold(a["V"], a["F"], a["bnd"], state1.nV, fixedAir=a["fixed"], rest8=a["rest8"])
    assert np.array_equal(s.localVI2Global, a["localVI2Global"])          # Scaffold.cpp:179-184
    assert s.wholeMeshSize == state1.nV + a["V"].shape[0] - a["nBnd"]


def test_subdivide_preserves_energy_density(port):
    from optcuts_b200 import synth
    V_rest, F, UV = random_mesh(6, n=6)
    r8, sc, _ = port.rest_features(V_rest, F)
    e0 = port.energy(F, UV, r8, sc["surfaceArea"])
    for n in (2, 3, 5):
        Vr, Fn, Uv = synth.subdivide(V_rest, F, UV, n)
        assert Fn.shape[0] == n * n * F.shape[0]
        # closed count: V + E(n-1) + F(n-1)(n-2)/2
        E = len(np.unique(np.sort(np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]]), axis=1), axis=0))
        assert Vr.shape[0] == len(UV) + E * (n - 1) + F.shape[0] * (n - 1) * (n - 2) // 2
        r8n, scn, rc = port.rest_features(Vr, Fn)
        assert rc == 0 and abs(scn["surfaceArea"] - sc["surfaceArea"]) < 1e-12 * sc["surfaceArea"]
        # linear interpolation keeps the Jacobian of every sub-triangle: same area-weighted energy
        assert abs(port.energy(Fn, Uv, r8n, scn["surfaceArea"]) - e0) < 1e-11 * e0


def test_locality_order_is_a_permutation():
    from optcuts_b200 import synth
    V_rest, F, UV = random_mesh(7, n=7)
    pv, pf = synth.locality_order(UV, F)
    assert sorted(pv) == list(range(len(UV))) and sorted(pf) == list(range(len(F)))
    Vr, Fn, Uv = synth.apply_order(V_rest, F, UV, pv, pf)
    # same triangles geometrically
    a = np.sort(np.round(UV[F].reshape(len(F), -1), 12), axis=0)
    b = np.sort(np.round(Uv[Fn].reshape(len(F), -1), 12), axis=0)
    assert np.allclose(a, b)


# ---------------------------------------------------------------------------------------------------------------------
# the two-pass record / replay machinery of the candidate evaluation (shim/CudaCandidates.cpp) is decision-neutral
import os
import shutil
import subprocess

import pytest
from conftest import GOLDEN, ROOT

SELFCHECK = os.path.join(ROOT, "shim", "_build", "OptCuts_selfcheck_probe")


@pytest.mark.parametrize("name,mesh,args", [("torus_cfg1", "torus.obj", ["0.999", "1", "0", "4.1", "1", "0"]),
                                            ("bimba_cfg2", "bimba_i_f10000.obj", ["0.025", "1", "2", "4.1", "1", "0"])])
def test_candidate_record_replay_is_decision_neutral(tmp_path, name, mesh, args):
    """The reference host program with the patched TriMesh (querySplit / queryMerge run twice: record, then replay) and
    OCB_CANDIDATES_SELFCHECK=1 (the recorded local problems are solved by the reference's own nested Optimizer): the
    whole run must reproduce the reference's trace BIT FOR BIT -- same candidates, same keys, same decisions."""
    if not os.path.exists(SELFCHECK):
        pytest.skip("shim/_build/OptCuts_selfcheck_probe not built (make -C shim needs the reference headers)")
    for f in os.listdir(os.path.join(GOLDEN, "inputs")):
        shutil.copy(os.path.join(GOLDEN, "inputs", f), tmp_path)
    env = dict(os.environ, ORACLE_TRACE=str(tmp_path / "trace.txt"), OCB_CANDIDATES_SELFCHECK="1")
    r = subprocess.run([SELFCHECK, "100", str(tmp_path / mesh)] + args + ["t"], cwd=tmp_path, env=env, capture_output=True, text=True, errors="replace", timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(tmp_path / "trace.txt").read() == open(os.path.join(GOLDEN, "traces", name + "_trace.txt")).read()
