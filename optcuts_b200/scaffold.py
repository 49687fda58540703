"""Air-mesh scaffold data as the hot path consumes it (src/Scaffold.hpp:18-33).

The triangulation itself (boundary loops + Triangle, Scaffold.cpp:27-208) is host-side work of the
caller and out of scope: the host program hands the result over in this form.
"""
import numpy as np


class Scaffold:
    def __init__(self, V_air, F_air, bnd, nV_mesh, fixedAir=(), areaThres_AM=0.0, rest8=None, ctx=None):
        self.V = np.asfortranarray(V_air, dtype=np.float64)
        self.F = np.asfortranarray(F_air, dtype=np.int32)
        self.bnd = np.asarray(bnd, dtype=np.int32).ravel()
        nVa, nB = self.V.shape[0], len(self.bnd)
        # Scaffold.cpp:179-184
        self.localVI2Global = np.concatenate([self.bnd, nV_mesh + np.arange(nVa - nB, dtype=np.int32)]).astype(np.int32)
        self.wholeMeshSize = nV_mesh + nVa - nB
        self.fixedAir = np.asarray(sorted(int(v) for v in fixedAir), dtype=np.int32)
        self.areaThres_AM = float(areaThres_AM)
        self.rest8 = rest8
        if rest8 is None:
            if ctx is None:
                raise ValueError("Scaffold needs rest8 or a Context to compute it")
            V3 = np.hstack([self.V, np.zeros((nVa, 1))])
            self.rest8, _ = ctx.rest_features(V3, self.F, self.areaThres_AM)

    @property
    def nBnd(self):
        return len(self.bnd)

    def F_global(self):
        return self.localVI2Global[self.F]
