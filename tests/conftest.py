"""Shared fixtures.  `-m "not gpu"` covers the oracle (against the golden vectors produced by the
reference itself), the host-side logic and the C-ABI's symbol table; `-m gpu` are the parity tests
proper: CUDA path through the C-ABI against the oracle on the same inputs.

Nothing here reads /root/reference: the golden fixtures in tests/golden/ were generated from it by
tests/golden/make_golden.py, and oracle/_ref/ (the compiled, unmodified reference) travels as a
prebuilt.  Tests that need oracle/_ref skip when it is absent.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running")


@pytest.fixture(scope="session")
def golden():
    g = np.load(os.path.join(GOLDEN, "bimba_cfg2_states.npz"))
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def port():
    from oracle import portapi
    portapi.lib()
    return portapi


@pytest.fixture(scope="session")
def ref():
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref/liboptcuts_ref.so not built (make -C oracle ref needs /root/reference)")
    refapi.lib()
    return refapi


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the gpu-marked tests; fails loudly (no fallback) if the library is missing."""
    import optcuts_b200 as ob
    c = ob.Context(0)
    yield c
    c.close()


class State:
    """A reference state from the golden file: mesh + air mesh + what the reference computed from it."""

    def __init__(self, g, tag, rt, nxt):
        self.g, self.tag, self.rt, self.nxt = g, tag, rt, nxt
        self.V_rest, self.F, self.UV = g[tag + "V_rest"], g[tag + "F"], g[tag + "V"]
        self.cohE = g[tag + "cohE"].reshape(-1, 4) if g[tag + "cohE"].size else np.zeros((0, 4), np.int32)
        self.fixed = g[tag + "fixedVert"]
        self.rest8, self.surfaceArea = g[rt + "rest8"], float(g[rt + "surfaceArea"])
        self.avgEdgeLen, self.virtualRadius = float(g[rt + "avgEdgeLen"]), float(g[rt + "virtualRadius"])
        self.air = dict(V=g[tag + "air_V"], F=g[tag + "air_F"], localVI2Global=g[tag + "air_localVI2Global"],
                        nBnd=len(g[tag + "air_bnd"]), rest8=g[rt + "air_rest8"], fixed=g[rt + "air_fixed"],
                        bnd=g[tag + "air_bnd"], areaThres_AM=float(g[tag + "air_scalars"][2]))
        self.w_scaf = float(g[rt + "w_scaf"])
        self.p0 = float(g["energyParam0"])
        self.nV, self.nF = self.UV.shape[0], self.F.shape[0]

    def r(self, key):
        return self.g[self.rt + key]

    def next_uv(self):
        return self.g[self.nxt + "V"]

    def upload(self, ctx, with_air=True):
        ctx.set_mesh(self.nV, self.F, self.rest8, self.surfaceArea, self.fixed)
        ctx.set_uv(self.UV)
        if with_air:
            a = self.air
            ctx.set_air(a["F"], a["rest8"], a["localVI2Global"], a["nBnd"], a["fixed"], self.w_scaf / a["F"].shape[0])
            ctx.set_uv(None, a["V"])


@pytest.fixture(scope="session", params=["s1", "s100"])
def state(request, golden):
    if request.param == "s1":
        return State(golden, "s1_", "r1_", "s2_")
    return State(golden, "s100_", "r100_", "s101_")


@pytest.fixture(scope="session")
def state1(golden):
    return State(golden, "s1_", "r1_", "s2_")


@pytest.fixture(scope="session")
def state100(golden):
    return State(golden, "s100_", "r100_", "s101_")


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))
