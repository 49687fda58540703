#include "CudaSymDirichletEnergy.hpp"
#include "CudaCoordinateHint.hpp"

#include <cstring>
#include <stdexcept>
#include <string>

namespace OptCuts {

CudaSymDirichletEnergy::CudaSymDirichletEnergy(int p_minFaces)
    : minFaces(p_minFaces), ctx(NULL), boundSurface(0.0), boundUniform(false), boundNV(-1)
{
    if (ocb_create(&ctx, 0) != OCB_OK) throw std::runtime_error("ocb_create failed");
}
CudaSymDirichletEnergy::~CudaSymDirichletEnergy(void) { ocb_destroy(ctx); }

void CudaSymDirichletEnergy::check(int rc, const char* what) const
{
    if (rc < 0 && rc != OCB_ERR_INVERTED) throw std::runtime_error(std::string(what) + ": " + ocb_last_error(ctx));
}

// upload topology + rest features when they changed (TriMesh::F / fixedVert / features: TriMesh.hpp:30-68),
// then the current UVs (TriMesh::V) — the plugin is stateless towards its caller, like the reference's
bool CudaSymDirichletEnergy::bind(const TriMesh& data, bool uniformWeight) const
{
    const int nF = static_cast<int>(data.F.rows()), nV = static_cast<int>(data.V.rows());
    if (nF < minFaces) return false;
    std::vector<int> fixed(data.fixedVert.begin(), data.fixedVert.end());
    // keyed on CONTENT (F, fixed vertices, triangle areas): Eigen keeps a buffer's address across reassignments of the same
    // size, and a freed buffer can be handed to another TriMesh, so the address says nothing
    const bool same = boundNV == nV && boundUniform == uniformWeight && boundF.rows() == nF && fixed == boundFixed &&
                      boundTriArea.size() == data.triArea.size() && boundSurface == data.surfaceArea &&
                      std::memcmp(boundF.data(), data.F.data(), sizeof(int) * 3 * nF) == 0 &&
                      std::memcmp(boundTriArea.data(), data.triArea.data(), sizeof(double) * nF) == 0;
    if (!same) {
        Eigen::MatrixXd rest8(nF, 8);                 // column k = feature k  ==  8 x nF SoA in memory
        if (uniformWeight) rest8.col(0).setOnes(); else rest8.col(0) = data.triArea;   // w = 1 (Optimizer.cpp:775,794,838)
        rest8.col(1) = data.triAreaSq; rest8.col(2) = data.e0SqLen; rest8.col(3) = data.e1SqLen; rest8.col(4) = data.e0dote1;
        rest8.col(5) = data.e0SqLen_div_dbAreaSq; rest8.col(6) = data.e1SqLen_div_dbAreaSq; rest8.col(7) = data.e0dote1_div_dbAreaSq;
        check(ocb_set_mesh(ctx, nV, nF, data.F.data(), rest8.data(), uniformWeight ? 1.0 : data.surfaceArea,
                           fixed.data(), static_cast<int>(fixed.size())), "ocb_set_mesh");
        boundF = data.F; boundFixed = fixed; boundTriArea = data.triArea; boundSurface = data.surfaceArea; boundUniform = uniformWeight; boundNV = nV;
    }
    check(ocb_set_uv(ctx, data.V.data(), NULL), "ocb_set_uv");
    if (!uniformWeight) {                             // the mesh (not the air mesh): publish its UVs for the solver's hierarchy
        std::vector<double> xy(data.V.data(), data.V.data() + 2 * static_cast<size_t>(nV));
        int n = nV;
        cudaCoordinateHint(true, n, xy);
    }
    return true;
}

void CudaSymDirichletEnergy::computeEnergyVal(const TriMesh& data, double& energyVal, bool uniformWeight) const
{
    if (!bind(data, uniformWeight)) { SymDirichletEnergy::computeEnergyVal(data, energyVal, uniformWeight); return; }
    double tot, scaf;
    check(ocb_energy(ctx, 1.0, &tot, &energyVal, &scaf), "ocb_energy");
}

void CudaSymDirichletEnergy::getEnergyValPerElem(const TriMesh& data, Eigen::VectorXd& e, bool uniformWeight) const
{
    if (!bind(data, uniformWeight)) { SymDirichletEnergy::getEnergyValPerElem(data, e, uniformWeight); return; }
    e.resize(data.F.rows());
    check(ocb_energy_per_elem(ctx, 0, e.data()), "ocb_energy_per_elem");
}

void CudaSymDirichletEnergy::getEnergyValByElemID(const TriMesh& data, int elemI, double& energyVal, bool uniformWeight) const
{
    if (!bind(data, uniformWeight)) { SymDirichletEnergy::getEnergyValByElemID(data, elemI, energyVal, uniformWeight); return; }
    check(ocb_energy_by_elem(ctx, elemI, 0, &energyVal), "ocb_energy_by_elem");
}

void CudaSymDirichletEnergy::computeHessian(const TriMesh& data, Eigen::MatrixXd& Hessian, bool uniformWeight) const
{
    // dense flavour (SymDirichletEnergy.cpp:306-427): only sensible for small meshes; the device path takes meshes the plugin is
    // bound to (>= minFaces) that still fit a dense matrix
    if (data.V.rows() > 4096 || !bind(data, uniformWeight)) { SymDirichletEnergy::computeHessian(data, Hessian, uniformWeight); return; }
    Hessian.resize(data.V.rows() * 2, data.V.rows() * 2);
    check(ocb_hessian_dense(ctx, 0, Hessian.data()), "ocb_hessian_dense");       // symmetric: row- and column-major coincide
}

void CudaSymDirichletEnergy::computeGradient(const TriMesh& data, Eigen::VectorXd& gradient, bool uniformWeight) const
{
    if (!bind(data, uniformWeight)) { SymDirichletEnergy::computeGradient(data, gradient, uniformWeight); return; }
    gradient.resize(data.V.rows() * 2);
    double sqn;
    check(ocb_gradient(ctx, 1.0, gradient.data(), &sqn), "ocb_gradient");
}

void CudaSymDirichletEnergy::computeHessian(const TriMesh& data, Eigen::VectorXd* V, Eigen::VectorXi* I, Eigen::VectorXi* J,
                                            bool uniformWeight) const
{
    if (!bind(data, uniformWeight)) { SymDirichletEnergy::computeHessian(data, V, I, J, uniformWeight); return; }
    int64_t n = 0;
    check(ocb_hessian_triplets(ctx, 0, NULL, NULL, NULL, &n), "ocb_hessian_triplets");
    // the reference APPENDS to V/I/J (IglUtils::addBlockToMatrix grows them); Optimizer passes empty vectors
    const int64_t n0 = V->size();
    V->conservativeResize(n0 + n);
    Eigen::VectorXi tI, tJ;
    Eigen::VectorXi* pI = I ? I : &tI; Eigen::VectorXi* pJ = J ? J : &tJ;
    const int64_t i0 = pI->size(), j0 = pJ->size();
    pI->conservativeResize(i0 + n); pJ->conservativeResize(j0 + n);
    check(ocb_hessian_triplets(ctx, 0, V->data() + n0, pI->data() + i0, pJ->data() + j0, &n), "ocb_hessian_triplets");
}

void CudaSymDirichletEnergy::initStepSize(const TriMesh& data, const Eigen::VectorXd& searchDir, double& stepSize) const
{
    // searchDir covers the whole system (mesh + air DOFs, Optimizer.cpp:692-704); the mesh part comes first
    if (searchDir.size() < data.V.rows() * 2 || !bind(data, false)) { SymDirichletEnergy::initStepSize(data, searchDir, stepSize); return; }
    check(ocb_step_bound(ctx, searchDir.data(), &stepSize), "ocb_step_bound");
}

void CudaSymDirichletEnergy::computeDivGradPerVert(const TriMesh& data, Eigen::VectorXd& divGradPerVert) const
{
    if (!bind(data, false)) { SymDirichletEnergy::computeDivGradPerVert(data, divGradPerVert); return; }
    divGradPerVert.resize(data.V.rows());
    check(ocb_divgrad_scores(ctx, divGradPerVert.data()), "ocb_divgrad_scores");
}

}  // namespace OptCuts
