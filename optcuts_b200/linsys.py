"""CudaLinSysSolver — Python mirror of OptCuts::LinSysSolver<VectorXi,VectorXd>
(src/LinSysSolver/LinSysSolver.hpp:22-256) with the EigenLibSolver role
(src/LinSysSolver/EigenLibSolver.cpp) played by the device PCG (two-level additive Schwarz preconditioner).
"""
import numpy as np
from ._capi import Context


class CudaLinSysSolver:
    def __init__(self, ctx=None, device=0, rel_tol=1e-12, max_it=0):
        self.ctx = ctx if ctx is not None else Context(device)
        self.rel_tol, self.max_it = rel_tol, max_it
        self.numRows = 0
        self.last = None

    def set_type(self, threadAmt, mtype, is_upper_half=False):
        pass   # no-op, as in EigenLibSolver::set_type

    # LinSysSolver::set_pattern(vNeighbor, fixedVert), LinSysSolver.hpp:37-135
    def set_pattern(self, vNeighbor, fixedVert):
        if isinstance(vNeighbor, tuple):
            ptr, idx = vNeighbor
        else:   # list of sets, as the reference passes it
            ptr = np.zeros(len(vNeighbor) + 1, np.int32)
            ptr[1:] = np.cumsum([len(s) for s in vNeighbor])
            idx = np.fromiter((v for s in vNeighbor for v in sorted(s)), dtype=np.int32, count=int(ptr[-1]))
        self.numRows = 2 * (len(ptr) - 1)
        self.ctx.set_pattern(ptr, idx, sorted(fixedVert))

    # LinSysSolver::update_a(I, J, S), LinSysSolver.hpp:138-159
    def update_a(self, II, JJ, SS):
        self.ctx.update_values_triplets(II, JJ, SS)

    def analyze_pattern(self):
        pass   # nothing symbolic to do for PCG

    def factorize(self):
        self.ctx.factorize()
        return True

    def solve(self, rhs):
        x, self.last = self.ctx.solve(rhs, self.rel_tol, self.max_it)
        return x

    def multiply(self, x):
        return self.ctx.multiply(x)

    def getNumRows(self):
        return self.numRows

    def getNumNonzeros(self):
        return self.ctx.sizes()["nnz_upper"]

    def get_csr(self):
        """(ia, ja, a): 1-based upper-triangular CSR exactly as LinSysSolver::set_pattern lays it out."""
        return self.ctx.download_csr()

    def coeffMtr(self, rowI, colI):
        if rowI > colI:
            rowI, colI = colI, rowI
        ia, ja, a = self.get_csr()
        seg = slice(ia[rowI] - 1, ia[rowI + 1] - 1)
        hit = np.nonzero(ja[seg] == colI + 1)[0]
        return float(a[seg][hit[0]]) if len(hit) else 0.0
