"""Optimizer — Python mirror of OptCuts::Optimizer's geometry step (src/Optimizer.hpp:22-143,
src/Optimizer.cpp:154-261, 505-704, 764-850) running device-resident through the C-ABI.

State (UVs, gradient, matrix, search direction) stays in HBM between calls; per Newton iteration
only scalars come back, plus the mesh UVs when the host has to rebuild the air mesh.
The topology step (createFracture) and the scaffold triangulation stay with the host program.
"""
import numpy as np
from ._capi import Context
from .trimesh import merge_adjacency


class Optimizer:
    def __init__(self, data0, energyTerms=None, energyParams=(1.0,), propagateFracture=0, mute=True,
                 scaffolding=False, scaffold_builder=None, ctx=None, device=0,
                 pcg_rel_tol=1e-12, pcg_max_it=0):
        self.data0 = data0
        self.energyParams = list(energyParams)
        self.energyTerms = energyTerms
        self.mute = mute
        self.scaffolding = bool(scaffolding)
        self.scaffold_builder = scaffold_builder
        if self.scaffolding and scaffold_builder is None:
            raise ValueError("scaffolding needs a scaffold_builder(TriMesh) -> Scaffold (host-side Triangle call)")
        self.ctx = ctx if ctx is not None else Context(device)
        self.pcg_rel_tol, self.pcg_max_it = pcg_rel_tol, pcg_max_it
        self.allowEDecRelTol = True
        self.propagateFracture = propagateFracture
        self.globalIterNum = 0
        self.topoIter = 0
        self.relGL2Tol = 1.0e-12                        # Optimizer.cpp:70
        self.w_scaf = self.energyParams[0] * 0.01        # Optimizer.cpp:87 (frozen at construction)
        self.energyParamSum = float(sum(self.energyParams))
        self.scaffold = None
        self.result = None
        self.lastEnergyVal = 0.0
        self.energyVal_scaffold = 0.0
        self.energyVal_ET = [0.0]
        self.lastEDec = 0.0
        self.last_step = None
        self.history = []
        self._uv_on_host = True

    # ---- helpers -------------------------------------------------------------------------------
    def _upload_mesh(self):
        r = self.result
        self.ctx.set_mesh(r.nV, r.F, r.rest8, r.surfaceArea, sorted(r.fixedVert))
        self.ctx.set_uv(r.V)
        self._uv_on_host = True

    def _sync_uv_to_host(self):
        if not self._uv_on_host:
            self.result.V = self.ctx.get_uv()
            self._uv_on_host = True

    def _rebuild_scaffold(self):
        """Optimizer.cpp:236-239 / 253-256: new air mesh from the current UVs, merged adjacency."""
        self._sync_uv_to_host()
        s = self.scaffold_builder(self.result)
        self.scaffold = s
        self.ctx.set_air(s.F, s.rest8, s.localVI2Global, s.nBnd, s.fixedAir, self.w_scaf / s.F.shape[0])
        self.ctx.set_uv(None, s.V)
        adj = merge_adjacency(self.result.adjacency(), self.result.nV, s.F_global(), s.wholeMeshSize)
        fixed = sorted(set(self.result.fixedVert) | set(int(s.localVI2Global[v]) for v in s.fixedAir))
        self.ctx.set_pattern(adj[0], adj[1], fixed)

    def updateTargetGRes(self):
        d0 = self.data0
        self.targetGRes = self.energyParamSum * float(d0.nV - len(d0.fixedVert)) / float(d0.nV) * self.relGL2Tol

    def setRelGL2Tol(self, tol):
        self.relGL2Tol = tol
        self.updateTargetGRes()

    def setAllowEDecRelTol(self, b):
        self.allowEDecRelTol = bool(b)

    # ---- API -----------------------------------------------------------------------------------
    def precompute(self):
        """Optimizer::precompute (Optimizer.cpp:154-201)."""
        self.result = self.data0.copy()
        self._upload_mesh()
        et, esd, escaf = self.ctx.energy(self.energyParams[0])     # also the inversion check of :65-67
        if self.scaffolding:
            self._rebuild_scaffold()
        else:
            self.ctx.set_pattern_from_elements()
        self.lastEDec = 0.0
        self.updateTargetGRes()
        self.computeLastEnergyVal()

    def computeLastEnergyVal(self):
        et, esd, escaf = self.ctx.energy(self.energyParams[0])
        self.lastEnergyVal, self.energyVal_ET[0], self.energyVal_scaffold = et, esd, escaf
        return et

    def solve(self, maxIter=100):
        """Optimizer::solve (Optimizer.cpp:203-261) without the fracture-propagation branch."""
        for _ in range(maxIter):
            r = self.ctx.newton_step(self.energyParams[0], self.targetGRes, self.pcg_rel_tol, self.pcg_max_it,
                                     self.allowEDecRelTol)
            self.last_step = r
            self.history.append(r)
            if r["converged"]:
                self.lastEDec = 0.0
                self.globalIterNum += 1
                return 1
            self._uv_on_host = False
            self.lastEDec = r["lastEDec"]
            self.lastEnergyVal, self.energyVal_scaffold, self.energyVal_ET[0] = r["E_new"], r["E_scaf_new"], r["E_sd_new"]
            self.globalIterNum += 1
            if r["stopped"]:
                return 1
            if self.scaffolding:
                self._rebuild_scaffold()
        return 0

    def updateEnergyData(self, updateEVal=True, updateGradient=True, updateHessian=True):
        """Optimizer::updateEnergyData (Optimizer.cpp:323-371) — energy part; matrices are rebuilt per step."""
        self.energyParamSum = float(sum(self.energyParams))
        self.updateTargetGRes()
        if updateEVal:
            self.computeLastEnergyVal()

    def setConfig(self, config, iterNum, topoIter):
        """Optimizer::setConfig (Optimizer.cpp:287-300): re-upload everything from a TriMesh."""
        self.topoIter, self.globalIterNum = topoIter, iterNum
        self.result = config.copy()
        self._upload_mesh()
        if self.scaffolding:
            self._rebuild_scaffold()
        else:
            self.ctx.set_pattern_from_elements()
        self.updateEnergyData()

    # ---- getters ---------------------------------------------------------------------------------
    def getResult(self):
        self._sync_uv_to_host()
        return self.result

    def getScaffold(self):
        return self.scaffold

    def getAirMesh(self):
        return self.scaffold

    def isScaffolding(self):
        return self.scaffolding

    def getIterNum(self):
        return self.globalIterNum

    def getTopoIter(self):
        return self.topoIter

    def getLastEnergyVal(self, excludeScaffold=False):
        return self.lastEnergyVal - self.energyVal_scaffold if (excludeScaffold and self.scaffolding) else self.lastEnergyVal

    def getGradient(self):
        return self.ctx.gradient(self.energyParams[0])[0]

    def getSearchDir(self):
        return self.ctx.get_search_dir()
