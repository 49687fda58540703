// CudaSymDirichletEnergy — OptCuts::Energy plugin over the optcuts_b200 C-ABI.
//
// Subclass of OptCuts::SymDirichletEnergy (src/Energy/SymDirichletEnergy.hpp) that overrides the virtuals
// the Optimizer calls through its std::vector<Energy*> (Optimizer.cpp:694-696, 766-769, 785-788, 802-824):
// computeEnergyVal, getEnergyValPerElem, computeGradient, computeHessian (triplet flavour), initStepSize.
// Registered by the host program in place of `new OptCuts::SymDirichletEnergy()` (main.cpp:1570).
// Meshes below `minFaces` triangles (the local stencils of the topology step) stay on the inherited CPU
// code: a kernel launch costs more than their arithmetic.
#ifndef CudaSymDirichletEnergy_hpp
#define CudaSymDirichletEnergy_hpp

#include "SymDirichletEnergy.hpp"
#include "optcuts_b200.h"

#include <vector>

namespace OptCuts {

class CudaSymDirichletEnergy : public SymDirichletEnergy
{
public:
    explicit CudaSymDirichletEnergy(int minFaces = 2000);
    virtual ~CudaSymDirichletEnergy(void);

    virtual void computeEnergyVal(const TriMesh& data, double& energyVal, bool uniformWeight = false) const;
    virtual void getEnergyValPerElem(const TriMesh& data, Eigen::VectorXd& energyValPerElem, bool uniformWeight = false) const;
    virtual void computeGradient(const TriMesh& data, Eigen::VectorXd& gradient, bool uniformWeight = false) const;
    virtual void computeHessian(const TriMesh& data, Eigen::VectorXd* V,
                                Eigen::VectorXi* I = NULL, Eigen::VectorXi* J = NULL, bool uniformWeight = false) const;
    virtual void getEnergyValByElemID(const TriMesh& data, int elemI, double& energyVal, bool uniformWeight = false) const;
    virtual void computeHessian(const TriMesh& data, Eigen::MatrixXd& Hessian, bool uniformWeight = false) const;   // dense flavour
    virtual void initStepSize(const TriMesh& data, const Eigen::VectorXd& searchDir, double& stepSize) const;
    virtual void computeDivGradPerVert(const TriMesh& data, Eigen::VectorXd& divGradPerVert) const;

protected:
    bool bind(const TriMesh& data, bool uniformWeight) const;   // false: stay on the CPU path
    void check(int rc, const char* what) const;

    int minFaces;
    mutable ocb_ctx* ctx;
    mutable Eigen::MatrixXi boundF;                  // what is on the device
    mutable std::vector<int> boundFixed;
    mutable Eigen::VectorXd boundTriArea;
    mutable double boundSurface;
    mutable bool boundUniform;
    mutable int boundNV;
};

}  // namespace OptCuts
#endif
