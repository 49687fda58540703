"""Host-side mesh container mirroring the data members of OptCuts::TriMesh that the hot path
reads (src/TriMesh.hpp:30-68).  Feature arithmetic (computeFeatures, TriMesh.cpp:343-398) runs on
the device through ocb_rest_features; adjacency and seam bookkeeping are small host-side index work.
"""
import numpy as np

REST_NAMES = ("triArea", "triAreaSq", "e0SqLen", "e1SqLen", "e0dote1",
              "e0SqLen_div_dbAreaSq", "e1SqLen_div_dbAreaSq", "e0dote1_div_dbAreaSq")


class TriMesh:
    """V_rest (nV x 3), F (nF x 3), V = UV (nV x 2); separateTri=false flavour (TriMesh.cpp:41-199)."""

    def __init__(self, V_rest, F, UV, cohE=None, initSeamLen=0.0, areaThres_AM=0.0, ctx=None, fixedVert=(0,)):
        self.V_rest = np.asfortranarray(V_rest, dtype=np.float64)
        if self.V_rest.shape[1] == 2:   # air meshes are planar (Scaffold.cpp:174-175)
            self.V_rest = np.asfortranarray(np.hstack([self.V_rest, np.zeros((self.V_rest.shape[0], 1))]))
        self.F = np.asfortranarray(F, dtype=np.int32)
        self.V = np.asfortranarray(UV, dtype=np.float64).copy(order="F")
        self.cohE = np.zeros((0, 4), np.int32) if cohE is None or len(cohE) == 0 else np.asarray(cohE, np.int32).reshape(-1, 4)
        self.initSeamLen = float(initSeamLen)
        self.areaThres_AM = float(areaThres_AM)
        self.fixedVert = set(int(v) for v in fixedVert)      # computeFeatures(resetFixedV=true) pins vertex 0
        self.vertWeight = np.ones(self.V.shape[0])
        self.rest8 = None
        self._adj = None
        if ctx is not None:
            self.computeFeatures(ctx)

    @property
    def nV(self):
        return self.V.shape[0]

    @property
    def nF(self):
        return self.F.shape[0]

    def computeFeatures(self, ctx):
        """TriMesh::computeFeatures (TriMesh.cpp:323-399): per-triangle rest features on the device."""
        self.rest8, sc = ctx.rest_features(self.V_rest, self.F, self.areaThres_AM)
        for k, name in enumerate(REST_NAMES):
            setattr(self, name, self.rest8[k])
        self.surfaceArea, self.avgEdgeLen, self.virtualRadius = sc["surfaceArea"], sc["avgEdgeLen"], sc["virtualRadius"]
        c = self.cohE
        if len(c):
            self.boundaryEdge = (c.min(axis=1) < 0).astype(np.int32)
            self.edgeLen = np.linalg.norm(self.V_rest[c[:, 0]] - self.V_rest[c[:, 1]], axis=1)
        else:
            self.boundaryEdge = np.zeros(0, np.int32)
            self.edgeLen = np.zeros(0)
        self._adj = None
        return self

    def adjacency(self):
        """vNeighbor (TriMesh.cpp:444-453) as CSR (ptr, idx), rows ascending like std::set."""
        if self._adj is None:
            self._adj = adjacency_from_faces(self.F, self.nV)
        return self._adj

    def copy(self):
        m = TriMesh.__new__(TriMesh)
        m.__dict__.update(self.__dict__)
        m.V = self.V.copy(order="F")
        m.fixedVert = set(self.fixedVert)
        return m


def adjacency_from_faces(F, nV):
    F = np.asarray(F)
    a = np.concatenate([F[:, 0], F[:, 1], F[:, 1], F[:, 2], F[:, 2], F[:, 0]]).astype(np.int64)
    b = np.concatenate([F[:, 1], F[:, 0], F[:, 2], F[:, 1], F[:, 0], F[:, 2]]).astype(np.int64)
    key = np.unique(a * nV + b)
    rows, cols = key // nV, key % nV
    ptr = np.zeros(nV + 1, np.int64)
    np.add.at(ptr, rows + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr.astype(np.int32), cols.astype(np.int32)


def merge_adjacency(adj_mesh, nV, F_air_global, nVtot):
    """Scaffold::mergeVNeighbor (Scaffold.cpp:295-305): mesh adjacency + air-mesh adjacency in global ids."""
    ptr, idx = adj_mesh
    rows = np.repeat(np.arange(nV, dtype=np.int64), np.diff(ptr))
    Fa = np.asarray(F_air_global)
    a = np.concatenate([rows, Fa[:, 0], Fa[:, 1], Fa[:, 1], Fa[:, 2], Fa[:, 2], Fa[:, 0]]).astype(np.int64)
    b = np.concatenate([idx.astype(np.int64), Fa[:, 1], Fa[:, 0], Fa[:, 2], Fa[:, 1], Fa[:, 0], Fa[:, 2]]).astype(np.int64)
    key = np.unique(a * nVtot + b)
    r, c = key // nVtot, key % nVtot
    p = np.zeros(nVtot + 1, np.int64)
    np.add.at(p, r + 1, 1)
    return np.cumsum(p).astype(np.int32), c.astype(np.int32)
