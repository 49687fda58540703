// TEST INFRASTRUCTURE — not on the product path.
//
// C API over the UNMODIFIED reference classes (OptCuts::TriMesh / SymDirichletEnergy /
// EigenLibSolver / Optimizer / Scaffold, compiled from /root/reference by oracle/Makefile into
// oracle/_ref/liboptcuts_ref.so).  Used by tests/ to (i) pin our C restatement (oracle/port) and
// (ii) check the CUDA path against the real reference, and by bench.py's cpu_baseline /
// --impl reference legs to time the reference's own TBB+Eigen path.
//
// Everything here only *calls* the reference; the globals below are the ones the reference's
// translation units `extern` from main.cpp (main.cpp:36-104), which is not linked into the library.
#include "TriMesh.hpp"
#include "Optimizer.hpp"
#include "Scaffold.hpp"
#include "SymDirichletEnergy.hpp"
#include "EigenLibSolver.hpp"
#include "IglUtils.hpp"
#include "Timer.hpp"

#include <fstream>
#include <cstring>
#include <cstdint>
#include <sys/time.h>

// ---- globals the reference TUs expect (Optimizer.cpp:28-33, TriMesh.cpp:25-32) ----
OptCuts::MethodType methodType = OptCuts::MT_OPTCUTS;
std::string outputFolderPath = "/tmp/";
bool fractureMode = false;
std::ofstream logFile;
Timer timer, timer_step;
std::vector<std::pair<double, double>> energyChanges_bSplit, energyChanges_iSplit, energyChanges_merge;
std::vector<std::vector<int>> paths_bSplit, paths_iSplit, paths_merge;
std::vector<Eigen::MatrixXd> newVertPoses_bSplit, newVertPoses_iSplit, newVertPoses_merge;
double filterExp_in = 0.6;
int inSplitTotalAmt = 0;

using namespace OptCuts;
typedef Eigen::Map<const Eigen::MatrixXd> CMapXd;
typedef Eigen::Map<const Eigen::MatrixXi> CMapXi;

namespace {
struct TimerInit {
    TimerInit() {
        // same activity order as main.cpp:1552-1565
        timer.new_activity("topology"); timer.new_activity("descent");
        timer.new_activity("scaffolding"); timer.new_activity("energyUpdate");
        const char* n[9] = {"matrixComputation", "matrixAssembly", "symbolicFactorization",
                            "numericalFactorization", "backSolve", "lineSearch",
                            "boundarySplit", "interiorSplit", "cornerMerge"};
        for (auto s : n) timer_step.new_activity(s);
    }
} timerInit;

struct OptProbe : public Optimizer {   // opens the protected members for read-out
    using Optimizer::Optimizer;
    using Optimizer::gradient; using Optimizer::searchDir; using Optimizer::linSysSolver;
    using Optimizer::I_mtr; using Optimizer::J_mtr; using Optimizer::V_mtr;
    using Optimizer::lastEnergyVal; using Optimizer::energyVal_scaffold; using Optimizer::energyVal_ET;
    using Optimizer::lastEDec; using Optimizer::targetGRes; using Optimizer::w_scaf;
    using Optimizer::scaffold; using Optimizer::result;
    using Optimizer::computeGradient; using Optimizer::computeEnergyVal; using Optimizer::computeHessian;
};
struct OptHandle {
    TriMesh* mesh;                       // data0 is held by reference inside Optimizer
    std::vector<Energy*> terms;
    std::vector<double> params;
    OptProbe* opt;
};
struct SolverHandle {
    EigenLibSolver<Eigen::VectorXi, Eigen::VectorXd> s;
};
}  // namespace

extern "C" {

// ---------------------------------------------------------------- mesh
// V_rest: nV x 3, F: nF x 3, UV: nV x 2, cohE: nCoh x 4 — all column-major (Eigen default).
void* ref_mesh_create(int nV, int nF, const double* V_rest, const int32_t* F, const double* UV,
                      int nCoh, const int32_t* cohE, double areaThres_AM)
{
    Eigen::MatrixXd Vr = CMapXd(V_rest, nV, 3);
    Eigen::MatrixXi Fm = CMapXi(F, nF, 3);
    Eigen::MatrixXd uv = CMapXd(UV, nV, 2);
    TriMesh* m = new TriMesh(Vr, Fm, uv, Eigen::MatrixXi(), false, 0.0, areaThres_AM);
    if (nCoh > 0) {
        m->cohE = CMapXi(cohE, nCoh, 4);
        m->computeFeatures(false, true);
    }
    return m;
}
void ref_mesh_destroy(void* h) { delete (TriMesh*)h; }
void ref_mesh_set_uv(void* h, const double* UV) {
    TriMesh* m = (TriMesh*)h;
    m->V = CMapXd(UV, m->V.rows(), 2);
}
void ref_mesh_get_uv(void* h, double* UV) {
    TriMesh* m = (TriMesh*)h;
    std::memcpy(UV, m->V.data(), sizeof(double) * m->V.size());
}
int ref_mesh_nV(void* h) { return (int)((TriMesh*)h)->V.rows(); }
int ref_mesh_nF(void* h) { return (int)((TriMesh*)h)->F.rows(); }
// rest8: 8 x nF SoA rows = triArea, triAreaSq, e0SqLen, e1SqLen, e0dote1, e0SqLen_div_dbAreaSq,
// e1SqLen_div_dbAreaSq, e0dote1_div_dbAreaSq ; scalars = surfaceArea, avgEdgeLen, virtualRadius
void ref_mesh_features(void* h, double* rest8, double* scalars)
{
    TriMesh* m = (TriMesh*)h;
    const Eigen::VectorXd* f[8] = {&m->triArea, &m->triAreaSq, &m->e0SqLen, &m->e1SqLen, &m->e0dote1,
                                   &m->e0SqLen_div_dbAreaSq, &m->e1SqLen_div_dbAreaSq, &m->e0dote1_div_dbAreaSq};
    const long n = m->F.rows();
    for (int k = 0; k < 8; ++k) std::memcpy(rest8 + k * n, f[k]->data(), sizeof(double) * n);
    scalars[0] = m->surfaceArea; scalars[1] = m->avgEdgeLen; scalars[2] = m->virtualRadius;
}
void ref_mesh_coh_features(void* h, int32_t* boundaryEdge, double* edgeLen)
{
    TriMesh* m = (TriMesh*)h;
    for (long i = 0; i < m->cohE.rows(); ++i) { boundaryEdge[i] = m->boundaryEdge[i]; edgeLen[i] = m->edgeLen[i]; }
}
// vertex adjacency (vNeighbor) as CSR; returns total length; pass NULL to query sizes
long ref_mesh_adjacency(void* h, int32_t* ptr, int32_t* idx)
{
    TriMesh* m = (TriMesh*)h;
    long tot = 0;
    for (size_t v = 0; v < m->vNeighbor.size(); ++v) {
        if (ptr) ptr[v] = (int32_t)tot;
        for (int nb : m->vNeighbor[v]) { if (idx) idx[tot] = nb; ++tot; }
    }
    if (ptr) ptr[m->vNeighbor.size()] = (int32_t)tot;
    return tot;
}
int ref_mesh_check_inversion(void* h) { return ((TriMesh*)h)->checkInversion(true) ? 1 : 0; }
double ref_mesh_seam_sparsity(void* h, int triSoup) { double s; ((TriMesh*)h)->computeSeamSparsity(s, triSoup != 0); return s; }

// ---------------------------------------------------------------- SymDirichletEnergy
double ref_sd_energy(void* h, int uniform) { SymDirichletEnergy SD; double e; SD.computeEnergyVal(*(TriMesh*)h, e, uniform != 0); return e; }
void ref_sd_energy_per_elem(void* h, int uniform, double* out) {
    SymDirichletEnergy SD; Eigen::VectorXd e; SD.getEnergyValPerElem(*(TriMesh*)h, e, uniform != 0);
    std::memcpy(out, e.data(), sizeof(double) * e.size());
}
void ref_sd_gradient(void* h, int uniform, double* g) {
    SymDirichletEnergy SD; Eigen::VectorXd v; SD.computeGradient(*(TriMesh*)h, v, uniform != 0);
    std::memcpy(g, v.data(), sizeof(double) * v.size());
}
// triplets; call with V==NULL to get the count
long ref_sd_hessian_triplets(void* h, int uniform, double* V, int32_t* I, int32_t* J) {
    static thread_local Eigen::VectorXd Vv; static thread_local Eigen::VectorXi Iv, Jv;
    if (!V) { SymDirichletEnergy SD; Vv.resize(0); Iv.resize(0); Jv.resize(0); SD.computeHessian(*(TriMesh*)h, &Vv, &Iv, &Jv, uniform != 0); return Vv.size(); }
    std::memcpy(V, Vv.data(), sizeof(double) * Vv.size());
    std::memcpy(I, Iv.data(), sizeof(int32_t) * Iv.size());
    std::memcpy(J, Jv.data(), sizeof(int32_t) * Jv.size());
    return Vv.size();
}
// per-triangle projected 6x6 blocks (row-major 36 per triangle), recovered from the triplet
// stream of a mesh WITHOUT fixed vertices influence: we use the dense path on 1 triangle at a time
void ref_sd_hessian_dense(void* h, int uniform, double* H /* (2nV)^2 col-major */) {
    SymDirichletEnergy SD; Eigen::MatrixXd M; SD.computeHessian(*(TriMesh*)h, M, uniform != 0);
    std::memcpy(H, M.data(), sizeof(double) * M.size());
}
double ref_sd_init_step_size(void* h, const double* searchDir, double stepSize0) {
    TriMesh* m = (TriMesh*)h; SymDirichletEnergy SD;
    Eigen::VectorXd p = Eigen::Map<const Eigen::VectorXd>(searchDir, m->V.rows() * 2);
    double s = stepSize0; SD.initStepSize(*m, p, s); return s;
}
void ref_sd_divgrad(void* h, double* out) {
    SymDirichletEnergy SD; Eigen::VectorXd v; SD.computeDivGradPerVert(*(TriMesh*)h, v);
    std::memcpy(out, v.data(), sizeof(double) * v.size());
}
void ref_make_pd6(double* M /* 36, symmetric */) {
    Eigen::Matrix<double, 6, 6> A = Eigen::Map<Eigen::Matrix<double, 6, 6>>(M);
    IglUtils::makePD(A);
    std::memcpy(M, A.data(), sizeof(double) * 36);
}

// ---------------------------------------------------------------- LinSysSolver / EigenLibSolver
void* ref_solver_create(void) { SolverHandle* s = new SolverHandle; s->s.set_type(1, -2); return s; }
void ref_solver_destroy(void* h) { delete (SolverHandle*)h; }
void ref_solver_set_pattern(void* h, int nV, const int32_t* adjPtr, const int32_t* adjIdx, int nFixed, const int32_t* fixed) {
    std::vector<std::set<int>> nb(nV);
    for (int v = 0; v < nV; ++v) for (int k = adjPtr[v]; k < adjPtr[v + 1]; ++k) nb[v].insert(adjIdx[k]);
    std::set<int> fx(fixed, fixed + nFixed);
    ((SolverHandle*)h)->s.set_pattern(nb, fx);
}
void ref_solver_update_a(void* h, long n, const int32_t* I, const int32_t* J, const double* S) {
    Eigen::VectorXi Iv = Eigen::Map<const Eigen::VectorXi>(I, n), Jv = Eigen::Map<const Eigen::VectorXi>(J, n);
    Eigen::VectorXd Sv = Eigen::Map<const Eigen::VectorXd>(S, n);
    ((SolverHandle*)h)->s.update_a(Iv, Jv, Sv);
}
int ref_solver_num_rows(void* h) { return ((SolverHandle*)h)->s.getNumRows(); }
long ref_solver_nnz(void* h) { return ((SolverHandle*)h)->s.get_ja().size(); }
void ref_solver_get_csr(void* h, int32_t* ia, int32_t* ja, double* a) {
    auto& s = ((SolverHandle*)h)->s;
    std::memcpy(ia, s.get_ia().data(), sizeof(int32_t) * s.get_ia().size());
    std::memcpy(ja, s.get_ja().data(), sizeof(int32_t) * s.get_ja().size());
    // EigenLibSolver keeps values in its own coefMtr; the base-class `a` is what update_a filled
    std::memcpy(a, s.get_a().data(), sizeof(double) * s.get_a().size());
}
int ref_solver_factorize(void* h) { auto& s = ((SolverHandle*)h)->s; s.analyze_pattern(); return s.factorize() ? 1 : 0; }
void ref_solver_solve(void* h, const double* rhs, double* x) {
    auto& s = ((SolverHandle*)h)->s; const int n = s.getNumRows();
    Eigen::VectorXd b = Eigen::Map<const Eigen::VectorXd>(rhs, n), r;
    s.solve(b, r);
    std::memcpy(x, r.data(), sizeof(double) * n);
}

// ---------------------------------------------------------------- Optimizer (global, sparse path)
// energyParam0 = 1 - lambda (main.cpp:1569).  propagateFracture = 0, mute selectable.
void* ref_opt_create(void* meshH, double energyParam0, int scaffolding, int mute)
{
    OptHandle* o = new OptHandle;
    o->mesh = (TriMesh*)meshH;
    o->terms.push_back(new SymDirichletEnergy());
    o->params.push_back(energyParam0);
    o->opt = new OptProbe(*o->mesh, o->terms, o->params, 0, mute != 0, scaffolding != 0);
    o->opt->precompute();
    return o;
}
void ref_opt_destroy(void* h) { OptHandle* o = (OptHandle*)h; delete o->opt; for (auto t : o->terms) delete t; delete o; }
int ref_opt_solve(void* h, int maxIter) { return ((OptHandle*)h)->opt->solve(maxIter); }
void ref_opt_set_energy_param(void* h, double p0) { OptHandle* o = (OptHandle*)h; o->params[0] = p0; o->opt->updateEnergyData(true, false, false); }
// scalars: lastEnergyVal, energyVal_scaffold, energyVal_ET[0], lastEDec, targetGRes, w_scaf, ||g||^2, iterNum
void ref_opt_scalars(void* h, double* out) {
    OptProbe* p = ((OptHandle*)h)->opt;
    out[0] = p->lastEnergyVal; out[1] = p->energyVal_scaffold; out[2] = p->energyVal_ET[0];
    out[3] = p->lastEDec; out[4] = p->targetGRes; out[5] = p->w_scaf;
    out[6] = p->gradient.size() ? p->gradient.squaredNorm() : 0.0; out[7] = p->getIterNum();
}
// sizes: nV, nF, nVa, nFa, nBnd, nSys(=wholeMeshSize*2 or 2nV), nnz(ja), nTriplets, nFixedAir
void ref_opt_sizes(void* h, long* out) {
    OptProbe* p = ((OptHandle*)h)->opt;
    const bool sc = p->isScaffolding();
    out[0] = p->result.V.rows(); out[1] = p->result.F.rows();
    out[2] = sc ? p->scaffold.airMesh.V.rows() : 0; out[3] = sc ? p->scaffold.airMesh.F.rows() : 0;
    out[4] = sc ? p->scaffold.bnd.size() : 0;
    out[5] = sc ? p->scaffold.wholeMeshSize * 2 : p->result.V.rows() * 2;
    out[6] = p->linSysSolver->get_ja().size(); out[7] = p->V_mtr.size();
    out[8] = sc ? (long)p->scaffold.airMesh.fixedVert.size() : 0;
}
void ref_opt_get_uv(void* h, double* UV) { OptProbe* p = ((OptHandle*)h)->opt; std::memcpy(UV, p->result.V.data(), sizeof(double) * p->result.V.size()); }
void ref_opt_get_air(void* h, double* Va, int32_t* Fa, int32_t* localVI2Global, double* rest8, double* scalars, int32_t* fixedAir) {
    OptProbe* p = ((OptHandle*)h)->opt; const TriMesh& am = p->scaffold.airMesh;
    if (Va) std::memcpy(Va, am.V.data(), sizeof(double) * am.V.size());
    if (Fa) std::memcpy(Fa, am.F.data(), sizeof(int32_t) * am.F.size());
    if (localVI2Global) std::memcpy(localVI2Global, p->scaffold.localVI2Global.data(), sizeof(int32_t) * p->scaffold.localVI2Global.size());
    if (rest8) ref_mesh_features((void*)&am, rest8, scalars);
    if (fixedAir) { int k = 0; for (int v : am.fixedVert) fixedAir[k++] = v; }
}
void ref_opt_get_gradient(void* h, double* g) { OptProbe* p = ((OptHandle*)h)->opt; std::memcpy(g, p->gradient.data(), sizeof(double) * p->gradient.size()); }
void ref_opt_get_search_dir(void* h, double* d) { OptProbe* p = ((OptHandle*)h)->opt; std::memcpy(d, p->searchDir.data(), sizeof(double) * p->searchDir.size()); }
void ref_opt_get_csr(void* h, int32_t* ia, int32_t* ja, double* a) {
    OptProbe* p = ((OptHandle*)h)->opt; auto* s = p->linSysSolver;
    std::memcpy(ia, s->get_ia().data(), sizeof(int32_t) * s->get_ia().size());
    std::memcpy(ja, s->get_ja().data(), sizeof(int32_t) * s->get_ja().size());
    std::memcpy(a, s->get_a().data(), sizeof(double) * s->get_a().size());
}
void ref_opt_get_triplets(void* h, int32_t* I, int32_t* J, double* V) {
    OptProbe* p = ((OptHandle*)h)->opt;
    std::memcpy(I, p->I_mtr.data(), sizeof(int32_t) * p->I_mtr.size());
    std::memcpy(J, p->J_mtr.data(), sizeof(int32_t) * p->J_mtr.size());
    std::memcpy(V, p->V_mtr.data(), sizeof(double) * p->V_mtr.size());
}
// recompute pieces at the CURRENT state (what the next solve(1) will start from)
void ref_opt_recompute_gradient(void* h) { OptProbe* p = ((OptHandle*)h)->opt; p->computeGradient(p->result, p->scaffold, p->gradient); }
double ref_opt_recompute_energy(void* h) { OptProbe* p = ((OptHandle*)h)->opt; double e; p->computeEnergyVal(p->result, p->scaffold, e); return e; }


// ---------------------------------------------------------------- local stencil solve (topology candidates)
// What TriMesh::computeLocalEdDec_* do after building the local mesh (TriMesh.cpp:2376-2381, 2498-2504, 2771-2781):
// nested Optimizer(dense = true, mute, no scaffold), setRelGL2Tol(tol), solve(maxIter), energy of the result.
void ref_mesh_reset_fixed(void* h, int n, const int32_t* fixed) {
    std::set<int> fx(fixed, fixed + n);
    ((TriMesh*)h)->resetFixedVert(fx);
}
// out: E_init (after precompute), E_final, iterations ; UV_out: nV x 2 col-major
void ref_local_solve(void* meshH, double relGL2Tol, int maxIter, double* out, double* UV_out) {
    TriMesh* m = (TriMesh*)meshH;
    SymDirichletEnergy SD;
    std::vector<Energy*> terms(1, &SD);
    std::vector<double> params(1, 1.0);
    OptProbe opt(*m, terms, params, 0, true, false, Eigen::MatrixXd(), Eigen::MatrixXi(), Eigen::VectorXi(), true);
    opt.precompute();
    out[0] = opt.getLastEnergyVal();
    opt.setRelGL2Tol(relGL2Tol);
    opt.solve(maxIter);
    double e; opt.computeEnergyVal(opt.result, opt.scaffold, e, true);
    out[1] = e; out[2] = opt.getIterNum();
    std::memcpy(UV_out, opt.getResult().V.data(), sizeof(double) * opt.getResult().V.size());
}

// ---------------------------------------------------------------- Scaffold (air mesh) from a mesh
// OptCuts::Scaffold(mesh) — boundary loops + bbox ring + Triangle (Scaffold.cpp:27-208); host-side work
// of the caller that the parity tests need in order to free-run with bijectivity on.
void* ref_scaffold_create(void* meshH) { TriMesh* m = (TriMesh*)meshH; return new Scaffold(*m); }
void ref_scaffold_destroy(void* h) { delete (Scaffold*)h; }
// sizes: nVa, nFa, nBnd, nFixedAir, wholeMeshSize
void ref_scaffold_sizes(void* h, long* out) {
    Scaffold* s = (Scaffold*)h;
    out[0] = s->airMesh.V.rows(); out[1] = s->airMesh.F.rows(); out[2] = s->bnd.size();
    out[3] = (long)s->airMesh.fixedVert.size(); out[4] = s->wholeMeshSize;
}
void ref_scaffold_get(void* h, double* Va, int32_t* Fa, int32_t* bnd, double* rest8, double* scalars, int32_t* fixedAir, double* areaThres) {
    Scaffold* s = (Scaffold*)h; const TriMesh& am = s->airMesh;
    std::memcpy(Va, am.V.data(), sizeof(double) * am.V.size());
    std::memcpy(Fa, am.F.data(), sizeof(int32_t) * am.F.size());
    std::memcpy(bnd, s->bnd.data(), sizeof(int32_t) * s->bnd.size());
    ref_mesh_features((void*)&am, rest8, scalars);
    int k = 0; for (int v : am.fixedVert) fixedAir[k++] = v;
    *areaThres = am.areaThres_AM;
}

// ---------------------------------------------------------------- timers (main.cpp:357-361 order)
void ref_timers_reset(void) { for (int i = 0; i < 4; ++i) timer.reset(i); for (int i = 0; i < 9; ++i) timer_step.reset(i); }
void ref_timers_get(double* t4, double* step9) { for (int i = 0; i < 4; ++i) t4[i] = timer.timing(i); for (int i = 0; i < 9; ++i) step9[i] = timer_step.timing(i); }
void ref_set_output_folder(const char* path) { outputFolderPath = path; if (logFile.is_open()) logFile.close(); logFile.open(outputFolderPath + "log.txt"); }

}  // extern "C"
