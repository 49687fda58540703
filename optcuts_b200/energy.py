"""SymDirichletEnergy — Python mirror of OptCuts::Energy / OptCuts::SymDirichletEnergy
(src/Energy/Energy.hpp:16-46, src/Energy/SymDirichletEnergy.hpp) on top of the C-ABI.

Same method names, argument meaning and statelessness as the reference: every call takes the
TriMesh, uploads what changed and runs the CUDA kernels.  No CPU path.
"""
import numpy as np
from ._capi import Context


class SymDirichletEnergy:
    needRefactorize = True   # Energy(true), SymDirichletEnergy.cpp:647-651

    def __init__(self, ctx=None, device=0):
        self.ctx = ctx if ctx is not None else Context(device)
        self._key = None

    def getNeedRefactorize(self):
        return self.needRefactorize

    def _bind(self, data, uniformWeight):
        key = (id(data), id(data.F), id(data.rest8), bool(uniformWeight), tuple(sorted(data.fixedVert)))
        if key != self._key:
            rest8, surf = data.rest8, data.surfaceArea
            if uniformWeight:   # w = 1: unit areas over unit surface (Optimizer.cpp:775,794,838)
                rest8 = data.rest8.copy()
                rest8[0] = 1.0
                surf = 1.0
            self.ctx.set_mesh(data.nV, data.F, rest8, surf, sorted(data.fixedVert))
            self._key = key
        self.ctx.set_uv(data.V)

    # Energy::computeEnergyVal (Energy.cpp:35-40)
    def computeEnergyVal(self, data, uniformWeight=False):
        self._bind(data, uniformWeight)
        return self.ctx.energy(1.0, check_inversion=False)[1]

    # SymDirichletEnergy::getEnergyValPerElem (SymDirichletEnergy.cpp:24-46)
    def getEnergyValPerElem(self, data, uniformWeight=False):
        self._bind(data, uniformWeight)
        return self.ctx.energy_per_elem(False)

    # SymDirichletEnergy::getEnergyValByElemID (:48-68)
    def getEnergyValByElemID(self, data, elemI, uniformWeight=False):
        self._bind(data, uniformWeight)
        return self.ctx.energy_by_elem(int(elemI), False)

    # SymDirichletEnergy::computeHessian, dense flavour (:306-427)
    def computeHessianDense(self, data, uniformWeight=False):
        self._bind(data, uniformWeight)
        return self.ctx.hessian_dense(False)

    # SymDirichletEnergy::computeGradient (:258-304)
    def computeGradient(self, data, uniformWeight=False):
        self._bind(data, uniformWeight)
        return self.ctx.gradient(1.0)[0]

    # SymDirichletEnergy::computeHessian, triplet flavour (:429-549): returns (V, I, J)
    def computeHessian(self, data, uniformWeight=False):
        self._bind(data, uniformWeight)
        I, J, V = self.ctx.hessian_triplets(False)
        return V, I, J

    def computeHessianBlocks(self, data, uniformWeight=False):
        self._bind(data, uniformWeight)
        return self.ctx.hessian_blocks(False)

    # SymDirichletEnergy::initStepSize (:551-610)
    def initStepSize(self, data, searchDir, stepSize):
        self._bind(data, False)
        return self.ctx.step_bound(np.asarray(searchDir, dtype=np.float64), stepSize)

    # SymDirichletEnergy::computeDivGradPerVert (:108-149)
    def computeDivGradPerVert(self, data):
        self._bind(data, False)
        return self.ctx.divgrad_scores()

    # SymDirichletEnergy::checkEnergyVal (:612-645): isometric map => every triangle gives 4w
    def checkEnergyVal(self, data):
        P = data.V_rest
        F = data.F
        e0, e1 = P[F[:, 1]] - P[F[:, 0]], P[F[:, 2]] - P[F[:, 0]]
        l0 = np.linalg.norm(e0, axis=1)
        x2 = np.einsum("ij,ij->i", e0, e1) / l0
        y2 = np.linalg.norm(np.cross(e0, e1), axis=1) / l0
        soupF = np.arange(3 * len(F), dtype=np.int32).reshape(-1, 3)
        UV = np.zeros((3 * len(F), 2))
        UV[1::3, 0] = l0
        UV[2::3, 0] = x2
        UV[2::3, 1] = y2
        ctx = self.ctx
        rest8 = data.rest8
        ctx.set_mesh(3 * len(F), soupF, rest8, data.surfaceArea, [])
        ctx.set_uv(UV)
        self._key = None
        per = ctx.energy_per_elem(False)
        return float(np.sum(per - 4.0 * data.triArea / data.surfaceArea))
