// Device-resident Newton iteration inside the reference's own Optimizer (Optimizer.cpp:154-261, 323-371, 505-673, 764-843).
//
// The reference's plugin surface is leaky: Optimizer instantiates `SymDirichletEnergy SD` on the stack for the
// scaffold term (Optimizer.cpp:701, 774, 793, 811, 835) and drives the solve call by call through Eigen triplets.  These
// hooks are the "5-line backend hook in a patched copy of Optimizer.cpp" the survey proposes: shim/Makefile inserts ONE
// macro at the top of computeEnergyVal, computeGradient, computeHessian and solve_oneStep of a build-time copy of the
// reference's Optimizer.cpp (sed, nothing of the reference is committed; INTEGRATION.md shows the diff).  Each macro
// collects references to the Optimizer's members (they are protected, a macro expands inside the member function) and
// calls a plain function of shim/CudaOptimizer.cpp, which talks to liboptcuts_b200.so through the C-ABI:
//
//   computeEnergyVal  -> ocb_energy          mesh term + scaffold term in one launch
//   computeGradient   -> ocb_gradient        fused gradient + energy + norm pass (the gradient comes back for the host's
//                                            own convergence test, Optimizer.cpp:210-221)
//   computeHessian    -> ocb_hessian_assemble   no triplets: the following set_pattern / update_a calls are no-ops
//   solve_oneStep     -> ocb_newton_step_ex     Hessian (unless fractureInitiated) + preconditioner + PCG + step bound +
//                                            line search, then the new UVs of mesh and air mesh are read back
//
// Only the global (sparse) optimizer takes this path; the nested dense optimizers of the topology step (useDense) keep
// the reference's code.  OCB_DEVICE_NEWTON=0 turns the hooks off (the call-by-call plugin path of round 1 remains).
#ifndef CudaOptimizerHooks_hpp
#define CudaOptimizerHooks_hpp

#include "TriMesh.hpp"
#include "Scaffold.hpp"
#include "LinSysSolver.hpp"

#include <vector>

namespace OptCuts {

struct OcbOptView {
    TriMesh& result; Scaffold& scaffold; bool scaffolding; bool mute;
    const std::vector<double>& energyParams; double w_scaf; double targetGRes; bool allowEDecRelTol; bool fractureInitiated;
    Eigen::VectorXd& gradient; Eigen::VectorXd& searchDir; double& lastEnergyVal; double& lastEDec;
    std::vector<double>& energyVal_ET; double& energyVal_scaffold; std::vector<Eigen::VectorXd>& gradient_ET;
    LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* solver;
};

bool ocbHookEnergy(OcbOptView& v, double& energyVal, bool excludeScaffold);
bool ocbHookGradient(OcbOptView& v, Eigen::VectorXd& gradient, bool excludeScaffold);
bool ocbHookHessian(OcbOptView& v);
bool ocbHookStep(OcbOptView& v, bool& stopped);
bool ocbDeviceResident(LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* solver);      // the hooks drive this optimizer's solver

}  // namespace OptCuts

#define OCB_OPT_VIEW(G) OptCuts::OcbOptView ocbV_{result, scaffold, scaffolding, mute, energyParams, w_scaf, targetGRes, allowEDecRelTol, \
                                                  fractureInitiated, G, searchDir, lastEnergyVal, lastEDec, energyVal_ET, energyVal_scaffold, gradient_ET, linSysSolver}
#define OCB_HOOK_ENERGY   if (!useDense && &data == &result) { OCB_OPT_VIEW(this->gradient); if (OptCuts::ocbHookEnergy(ocbV_, energyVal, excludeScaffold)) return; }
#define OCB_HOOK_GRADIENT if (!useDense && &data == &result && !excludeScaffold) { OCB_OPT_VIEW(this->gradient); if (OptCuts::ocbHookGradient(ocbV_, gradient, excludeScaffold)) return; }
#define OCB_HOOK_HESSIAN  if (!useDense && &data == &result) { OCB_OPT_VIEW(this->gradient); if (OptCuts::ocbHookHessian(ocbV_)) return; }
#define OCB_HOOK_STEP     if (!useDense) { OCB_OPT_VIEW(this->gradient); bool ocbStopped_ = false; \
                              if (OptCuts::ocbHookStep(ocbV_, ocbStopped_)) { fractureInitiated = false; \
                                  if (!mute && !ocbStopped_) writeEnergyValToFile(false); return ocbStopped_; } }
// Scaffold::mergeVNeighbor only feeds LinSysSolver::set_pattern (Optimizer.cpp:174, 358, 524), which the device-resident solver
// ignores: the 5 000 std::set copies per Newton iteration are skipped while the hooks are active
// `data_findExtrema = result` (Optimizer.cpp:381, 452: "potentially time-consuming", a whole-TriMesh copy with its std::map / std::set
// members, ~5 ms at 10k faces, once per createFracture call, i.e. once per Newton iteration while a fracture propagates) only feeds
// the viewer's "find extrema" channel (main.cpp:81, 1577): nothing reads it in headless runs.  OCB_KEEP_FIND_EXTREMA=1 keeps the copy.
namespace OptCuts { inline bool ocbKeepFindExtrema(void) { static const bool on = []() { const char* e = std::getenv("OCB_KEEP_FIND_EXTREMA"); return e && std::atoi(e) != 0; }(); return on; } }
#define OCB_COPY_FIND_EXTREMA if (useDense || OptCuts::ocbKeepFindExtrema() || !OptCuts::ocbDeviceResident(linSysSolver)) data_findExtrema = result;
#define OCB_MERGE_VNEIGHBOR if (useDense || !OptCuts::ocbDeviceResident(linSysSolver)) scaffold.mergeVNeighbor(result.vNeighbor, vNeighbor_withScaf);
#endif
