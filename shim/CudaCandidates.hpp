// Candidate evaluation of the topology step on the B200 (north_star kernel 6) inside the reference's own
// TriMesh::querySplit / queryMerge (TriMesh.cpp:548-759, 801-951).
//
// The reference scores every candidate split / merge by running a NESTED dense Optimizer on a local stencil
// (computeLocalEdDec_bSplit / _inSplit / _merge, TriMesh.cpp:2262-2794): with bijectivity on, ~100-1000 such solves per
// topology step, each re-triangulating its local air mesh (Triangle) after every Newton iteration.  Here the reference's
// code keeps doing what only it can do -- candidate generation and filtering, the local meshes after the local operation,
// the air loops (Scaffold::get1RingAirLoop / getCornerAirLoop), the decision logic -- and the nested solves run in lock
// step on the GPU:
//
//   pass 1 (record)  the query function runs once with OcbLocalOptimizer (a stand-in with Optimizer's call surface at the
//                    three nested-solve sites) RECORDING every local problem; results of this pass are discarded;
//   solve            all recorded problems advance one Newton iteration per round: the host triangulates the air region of
//                    every active problem (the same igl::triangle::triangulate("qYQ") call as Scaffold.cpp:169, all host
//                    threads), one ocb_stencil_newton_step launch does the iteration of every problem (one CTA each);
//   pass 2 (replay)  the query function runs again; OcbLocalOptimizer hands back the stored results (looked up by the
//                    CONTENT of the local problem, so no ordering assumption is made), and the reference's own code takes
//                    the decisions.
//
// shim/Makefile generates _build/TriMesh_cuda.cpp from the reference's TriMesh.cpp with sed: the include below, `Optimizer
// optimizer(localMesh` -> `OcbLocalOptimizer optimizer(localMesh` at the three sites, and OCB_HOOK_QUERY(...) at the top
// of querySplit and queryMerge.  OCB_DEVICE_CANDIDATES=0 keeps the reference's nested optimizers.
#ifndef CudaCandidates_hpp
#define CudaCandidates_hpp

#include "Optimizer.hpp"

namespace OptCuts {

class OcbLocalOptimizer
{
public:
    OcbLocalOptimizer(const TriMesh& data0, const std::vector<Energy*>& energyTerms, const std::vector<double>& energyParams,
                      int propagateFracture, bool mute, bool scaffolding,
                      const Eigen::MatrixXd& UV_bnds, const Eigen::MatrixXi& E, const Eigen::VectorXi& bnd, bool useDense);
    ~OcbLocalOptimizer(void);
    void precompute(void);
    void setRelGL2Tol(double tol);
    int solve(int maxIter = 100);
    void computeEnergyVal(const TriMesh& data, const Scaffold& scaffoldData, double& energyVal, bool excludeScaffold = false);
    TriMesh& getResult(void);
    const Scaffold& getScaffold(void) const;

private:
    void passThrough(int maxIter);
    const TriMesh& data0;
    const std::vector<Energy*>& energyTerms; const std::vector<double>& energyParams;
    bool scaffolding;
    const Eigen::MatrixXd& UV_bnds; const Eigen::MatrixXi& E; const Eigen::VectorXi& bnd;
    double tol;
    Optimizer* real;              // the reference's nested optimizer (hooks off, or a stencil over the kernel's limits)
    TriMesh result;
    Scaffold noScaffold;
    double Esd;
};

struct OcbBatchScope {
    OcbBatchScope(void);
    ~OcbBatchScope(void);
    bool outermost(void) const { return outer; }
    void solve(void);
private:
    bool outer;
};

}  // namespace OptCuts

#define OCB_HOOK_QUERY(call) OptCuts::OcbBatchScope ocbScope_; if (ocbScope_.outermost()) { call; ocbScope_.solve(); }
#endif
