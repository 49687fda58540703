// Lock-step evaluation of the topology step's local problems on the GPU (see shim/CudaCandidates.hpp).
#include "CudaCandidates.hpp"
#include "OcbThreadPool.hpp"
#include "SymDirichletEnergy.hpp"
#include "optcuts_b200.h"

#include <igl/triangle/triangulate.h>
#include <tbb/tbb.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>

namespace OptCuts {

namespace {

enum Mode { OFF = 0, RECORD = 1, REPLAY = 2 };

struct Key {
    unsigned long long a, b;
    bool operator<(const Key& o) const { return a < o.a || (a == o.a && b < o.b); }
};
struct Hasher {
    unsigned long long a = 1469598103934665603ull, b = 0x9e3779b97f4a7c15ull;
    void add(const void* data, size_t bytes) {
        const unsigned char* p = static_cast<const unsigned char*>(data);
        for (size_t i = 0; i < bytes; ++i) { a ^= p[i]; a *= 1099511628211ull; b = (b ^ p[i]) * 0xff51afd7ed558ccdull + 0x2545f4914f6cdd1dull; }
    }
    template <typename T> void pod(const T& v) { add(&v, sizeof(T)); }
};

struct Problem {
    // the local problem as the reference built it
    Eigen::MatrixXd V_rest, V; Eigen::MatrixXi F;
    std::vector<unsigned char> isFree;
    bool bij = false;
    Eigen::MatrixXd UV_bnds; Eigen::MatrixXi E; Eigen::VectorXi bnd;
    double areaThres = 0.0, targetGRes = 0.0;
    int maxIter = 100;
    // state of the lock-step solve
    Eigen::MatrixXd airV; Eigen::MatrixXi airF;       // this round's air mesh (Triangle output)
    double Esd = 0.0;
    int iters = 0, status = 0;                         // status: 0 running, 1 finished, < 0 not evaluated on the device
    TriMesh mesh;                                      // OCB_CANDIDATES_SELFCHECK only: the local mesh itself
};

struct Batch {
    std::atomic<int> mode{OFF};
    std::mutex mu;
    std::map<Key, int> index;
    std::vector<Problem> problems;
    ocb_ctx* ctx = NULL;
    // statistics
    long queries = 0, solved = 0, rounds = 0, passThroughs = 0;
    double tSolve = 0.0, tTriangle = 0.0, tDevice = 0.0, tPass1 = 0.0, tPass2 = 0.0, tMark = 0.0, tKey = 0.0, tCopy = 0.0;
};
Batch& batch(void) { static Batch B; return B; }

bool enabled(void)
{
    static const bool on = []() { const char* e = std::getenv("OCB_DEVICE_CANDIDATES"); return !(e && std::atoi(e) == 0); }();
    return on;
}
// OCB_CANDIDATES_SELFCHECK=1 (tests/test_host_logic.py, no GPU needed): the recorded problems are solved by the reference's
// own nested Optimizer instead of the device, so that the record / replay machinery alone is under test -- the run must
// then reproduce the reference's trace bit for bit.  Never set in production: it is the slow path by construction.
bool selfCheck(void)
{
    static const bool on = []() { const char* e = std::getenv("OCB_CANDIDATES_SELFCHECK"); return e && std::atoi(e) != 0; }();
    return on;
}
// OCB_CANDIDATES_VERIFY=<file> (tests/test_gpu_candidates.py): after the device solve, every recorded problem is ALSO solved
// by the reference's own nested Optimizer (host threads) and the two results are compared: one line per query in <file>
// (#problems, worst relative difference of the final E_SD, worst difference of the final UVs relative to the stencil's
// extent, #problems whose Newton iteration count differs), totals at exit.  The device results are the ones used.
const char* verifyPath(void)
{
    static const char* p = []() { const char* e = std::getenv("OCB_CANDIDATES_VERIFY"); return (e && *e) ? e : (const char*)NULL; }();
    return p;
}
double now(void) { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void report(void)
{
    Batch& B = batch();
    if (B.queries && (std::getenv("OCB_HOST_TIMING") || std::getenv("OCB_CANDIDATES_REPORT")))
        std::fprintf(stderr, "[ocb candidates] %ld queries, %ld local problems on the device in %ld lock-step rounds (%ld passed to the CPU), "
                             "%.3f s: Triangle %.3f s, device + packing %.3f s; the reference's query code around it: recording pass %.3f s, replay pass %.3f s\n",
                     B.queries, B.solved, B.rounds, B.passThroughs, B.tSolve, B.tTriangle, B.tDevice, B.tPass1, B.tPass2);
}

Key keyOf(const TriMesh& m, bool bij, const Eigen::MatrixXd& UV_bnds, const Eigen::MatrixXi& E, const Eigen::VectorXi& bnd, double tol, int maxIter)
{
    Hasher h;
    h.pod(m.V_rest.rows()); h.pod(m.F.rows()); h.pod(bij); h.pod(tol); h.pod(maxIter);
    h.add(m.V_rest.data(), sizeof(double) * m.V_rest.size());
    h.add(m.V.data(), sizeof(double) * m.V.size());
    h.add(m.F.data(), sizeof(int) * m.F.size());
    for (int v : m.fixedVert) h.pod(v);
    if (bij) {
        h.pod(bnd.size()); h.pod(UV_bnds.rows());
        h.add(bnd.data(), sizeof(int) * bnd.size());
        h.add(E.data(), sizeof(int) * E.size());
        // rows [0, bnd.size()) of UV_bnds are filled from the mesh when the air mesh is built (Scaffold.cpp:159-162) and
        // are uninitialised until then: only the outer loop is part of the problem
        for (int r = static_cast<int>(bnd.size()); r < UV_bnds.rows(); ++r) { h.pod(UV_bnds(r, 0)); h.pod(UV_bnds(r, 1)); }
    }
    Key k; k.a = h.a; k.b = h.b;
    return k;
}

// the local air mesh of one problem from its current UVs: Scaffold::Scaffold, local branch (Scaffold.cpp:153-169)
void buildAir(Problem& p)
{
    Eigen::MatrixXd UV = p.UV_bnds, H;
    for (int i = 0; i < p.bnd.size(); ++i) UV.row(i) = p.V.row(p.bnd[i]);
    igl::triangle::triangulate(UV, p.E, H, "qYQ", p.airV, p.airF);
}

void solveAll(Batch& B)
{
    const double t0 = now();
    std::vector<Problem>& P = B.problems;
    if (P.empty()) return;
    if (selfCheck()) {
        static SymDirichletEnergy SD;
        tbb::parallel_for(0, static_cast<int>(P.size()), 1, [&](int k) {
            Problem& p = P[k];
            std::vector<OptCuts::Energy*> terms(1, &SD);
            std::vector<double> params(1, 1.0);
            Optimizer opt(p.mesh, terms, params, 0, true, p.bij, p.UV_bnds, p.E, p.bnd, true);
            opt.precompute();
            opt.setRelGL2Tol(p.targetGRes / (static_cast<double>(p.mesh.V_rest.rows() - p.mesh.fixedVert.size()) / static_cast<double>(p.mesh.V_rest.rows())));
            opt.solve(p.maxIter);
            p.V = opt.getResult().V;
            SD.computeEnergyVal(opt.getResult(), p.Esd);
            p.status = 1;
        });
        B.solved += static_cast<long>(P.size());
        return;
    }
    if (!B.ctx) {
        if (ocb_create(&B.ctx, 0) != OCB_OK) { std::fprintf(stderr, "optcuts_b200: ocb_create failed\n"); std::exit(-1); }
        std::atexit(report);
    }
    std::vector<int> active(P.size());
    for (size_t i = 0; i < P.size(); ++i) active[i] = static_cast<int>(i);
    std::vector<int32_t> vp, tp, nvm, ntm, F, res;
    std::vector<double> Vr, UV, th, tg, UVo, out6;
    std::vector<unsigned char> fr;
    for (int round = 0; !active.empty(); ++round) {
        // ---- host: this round's air meshes (Triangle), all threads
        const double t1 = now();
        OcbThreadPool::get().run(static_cast<int>(active.size()), [&](int k) { Problem& p = P[active[k]]; if (p.bij) buildAir(p); });
        const double t2 = now();
        B.tTriangle += t2 - t1;
        // ---- pack: per problem the mesh's vertices, then the air mesh's own; the mesh's triangles, then the air mesh's
        const int nS = static_cast<int>(active.size());
        vp.assign(1, 0); tp.assign(1, 0); nvm.clear(); ntm.clear(); F.clear(); Vr.clear(); UV.clear(); fr.clear(); th.clear(); tg.clear();
        for (int k = 0; k < nS; ++k) {
            const Problem& p = P[active[k]];
            const int nVm = static_cast<int>(p.V.rows()), nTm = static_cast<int>(p.F.rows()), nB = static_cast<int>(p.bnd.size());
            const int nVa = p.bij ? static_cast<int>(p.airV.rows()) : 0, nTa = p.bij ? static_cast<int>(p.airF.rows()) : 0;
            const int nLoop = p.bij ? static_cast<int>(p.UV_bnds.rows()) : 0;
            for (int v = 0; v < nVm; ++v) {
                Vr.push_back(p.V_rest(v, 0)); Vr.push_back(p.V_rest(v, 1)); Vr.push_back(p.V_rest(v, 2));
                UV.push_back(p.V(v, 0)); UV.push_back(p.V(v, 1)); fr.push_back(p.isFree[v]);
            }
            for (int a = nB; a < nVa; ++a) {               // outer loop: fixed (fixAMBoundary, Scaffold.cpp:194-199); Steiner points: free
                Vr.push_back(0.0); Vr.push_back(0.0); Vr.push_back(0.0);
                UV.push_back(p.airV(a, 0)); UV.push_back(p.airV(a, 1)); fr.push_back(a >= nLoop ? 1 : 0);
            }
            for (int t = 0; t < nTm; ++t) for (int c = 0; c < 3; ++c) F.push_back(p.F(t, c));
            for (int t = 0; t < nTa; ++t) for (int c = 0; c < 3; ++c) { const int a = p.airF(t, c); F.push_back(a < nB ? p.bnd[a] : nVm + a - nB); }
            vp.push_back(vp.back() + nVm + (nVa > nB ? nVa - nB : 0)); tp.push_back(tp.back() + nTm + nTa);
            nvm.push_back(nVm); ntm.push_back(nTm); th.push_back(p.areaThres); tg.push_back(p.targetGRes);
        }
        ocb_stencil_step_batch sb;
        sb.nStencil = nS; sb.vert_ptr = vp.data(); sb.tri_ptr = tp.data(); sb.n_mesh_vert = nvm.data(); sb.n_mesh_tri = ntm.data();
        sb.V_rest = Vr.data(); sb.UV = UV.data(); sb.F = F.data(); sb.is_free = fr.data(); sb.area_thres = th.data(); sb.target_gres = tg.data();
        sb.w_scaf = 0.01;                                   // Optimizer.cpp:87 with energyParams = {1}
        UVo.resize(UV.size()); out6.resize(6 * static_cast<size_t>(nS)); res.resize(nS);
        const int rc = ocb_stencil_newton_step(B.ctx, &sb, UVo.data(), out6.data(), res.data());
        if (rc < 0) { std::fprintf(stderr, "optcuts_b200: ocb_stencil_newton_step failed (%d): %s\n", rc, ocb_last_error(B.ctx)); std::exit(-1); }
        // ---- unpack; a problem leaves the lock step when it converged, stopped, ran out of iterations or cannot run on the device
        std::vector<int> next;
        for (int k = 0; k < nS; ++k) {
            Problem& p = P[active[k]];
            if (res[k] < 0) { p.status = res[k]; continue; }
            const int nVm = static_cast<int>(p.V.rows());
            for (int v = 0; v < nVm; ++v) { p.V(v, 0) = UVo[2 * (static_cast<size_t>(vp[k]) + v)]; p.V(v, 1) = UVo[2 * (static_cast<size_t>(vp[k]) + v) + 1]; }
            p.Esd = out6[6 * static_cast<size_t>(k)];
            p.iters = round + 1;
            if (res[k] == 0 && round + 1 < p.maxIter) next.push_back(active[k]); else p.status = 1;
        }
        active.swap(next);
        B.rounds++;
        B.tDevice += now() - t2;
    }
    B.solved += static_cast<long>(P.size());
    B.tSolve += now() - t0;
    if (verifyPath()) {
        static SymDirichletEnergy SD;
        std::vector<double> dE(P.size(), 0.0), dV(P.size(), 0.0);
        std::vector<int> dIt(P.size(), 0), onDev(P.size(), 0);
        tbb::parallel_for(0, static_cast<int>(P.size()), 1, [&](int k) {
            Problem& p = P[k];
            if (p.status < 0) return;
            onDev[k] = 1;
            std::vector<OptCuts::Energy*> terms(1, &SD);
            std::vector<double> params(1, 1.0);
            Optimizer opt(p.mesh, terms, params, 0, true, p.bij, p.UV_bnds, p.E, p.bnd, true);
            opt.precompute();
            opt.setRelGL2Tol(p.targetGRes / (static_cast<double>(p.mesh.V_rest.rows() - p.mesh.fixedVert.size()) / static_cast<double>(p.mesh.V_rest.rows())));
            opt.solve(p.maxIter);
            double e = 0.0;
            SD.computeEnergyVal(opt.getResult(), e);
            dE[k] = std::abs(e - p.Esd) / std::abs(e);
            const Eigen::MatrixXd& Vr = opt.getResult().V;
            const double ext = (Vr.colwise().maxCoeff() - Vr.colwise().minCoeff()).maxCoeff();
            dV[k] = (Vr - p.V).cwiseAbs().maxCoeff() / ext;
            dIt[k] = (opt.getIterNum() != p.iters) ? 1 : 0;
        });
        double wE = 0.0, wV = 0.0; int nIt = 0, nDev = 0, nBij = 0;
        for (size_t k = 0; k < P.size(); ++k) { wE = std::max(wE, dE[k]); wV = std::max(wV, dV[k]); nIt += dIt[k]; nDev += onDev[k]; nBij += P[k].bij ? 1 : 0; }
        FILE* f = std::fopen(verifyPath(), "a");
        if (f) { std::fprintf(f, "query=%ld problems=%d on_device=%d bijective=%d worst_rel_Esd=%.3e worst_rel_UV=%.3e iter_count_differs=%d\n",
                              B.queries, static_cast<int>(P.size()), nDev, nBij, wE, wV, nIt); std::fclose(f); }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ the stand-in
OcbLocalOptimizer::OcbLocalOptimizer(const TriMesh& p_data0, const std::vector<Energy*>& p_energyTerms, const std::vector<double>& p_energyParams,
                                     int, bool, bool p_scaffolding, const Eigen::MatrixXd& p_UV_bnds, const Eigen::MatrixXi& p_E,
                                     const Eigen::VectorXi& p_bnd, bool)
    : data0(p_data0), energyTerms(p_energyTerms), energyParams(p_energyParams), scaffolding(p_scaffolding), UV_bnds(p_UV_bnds), E(p_E), bnd(p_bnd),
      tol(1.0e-12), real(NULL), Esd(0.0)
{
}
OcbLocalOptimizer::~OcbLocalOptimizer(void) { delete real; }

void OcbLocalOptimizer::precompute(void) {}
void OcbLocalOptimizer::setRelGL2Tol(double p_tol) { tol = p_tol; }

void OcbLocalOptimizer::passThrough(int maxIter)
{
    real = new Optimizer(data0, energyTerms, energyParams, 0, true, scaffolding, UV_bnds, E, bnd, true);
    real->precompute();
    real->setRelGL2Tol(tol);
    real->solve(maxIter);
    SymDirichletEnergy SD;
    SD.computeEnergyVal(real->getResult(), Esd);           // = Optimizer::computeEnergyVal(..., excludeScaffold = true) with energyParams = {1}
}

int OcbLocalOptimizer::solve(int maxIter)
{
    Batch& B = batch();
    const int mode = B.mode.load();
    if (mode == OFF) { passThrough(maxIter); return 0; }
    const Key key = keyOf(data0, scaffolding, UV_bnds, E, bnd, tol, maxIter);
    if (mode == RECORD) {
        std::lock_guard<std::mutex> lock(B.mu);
        if (B.index.find(key) == B.index.end()) {
            B.index[key] = static_cast<int>(B.problems.size());
            B.problems.emplace_back();
            Problem& p = B.problems.back();
            p.V_rest = data0.V_rest; p.V = data0.V; p.F = data0.F;
            p.isFree.assign(data0.V.rows(), 1);
            for (int v : data0.fixedVert) p.isFree[v] = 0;
            p.bij = scaffolding;
            if (scaffolding) { p.UV_bnds = UV_bnds; p.E = E; p.bnd = bnd; }
            const double edgeLen_eps = data0.avgEdgeLen * 0.5 * 0.1;                           // Scaffold.cpp:34, 156
            p.areaThres = std::sqrt(3.0) / 4.0 * edgeLen_eps * edgeLen_eps;                     // Scaffold.cpp:176
            p.targetGRes = 1.0 * static_cast<double>(data0.V_rest.rows() - data0.fixedVert.size()) / static_cast<double>(data0.V_rest.rows()) * tol;   // Optimizer.cpp:675-678
            p.maxIter = maxIter;
            if (selfCheck() || verifyPath()) p.mesh = data0;
        }
        result = data0;                                      // placeholder: pass 1's numbers are discarded
        Esd = 0.0;
        return 0;
    }
    // REPLAY
    const Problem* p = NULL;
    {
        std::lock_guard<std::mutex> lock(B.mu);
        const auto f = B.index.find(key);
        if (f != B.index.end()) p = &B.problems[f->second];
    }
    if (!p || p->status < 0) {                              // not recorded (cannot happen) or over the kernel's limits: the reference's own solve
        { std::lock_guard<std::mutex> lock(B.mu); B.passThroughs++; }
        passThrough(maxIter);
        return 0;
    }
    result = data0;
    result.V = p->V;
    Esd = p->Esd;
    return 0;
}

void OcbLocalOptimizer::computeEnergyVal(const TriMesh&, const Scaffold&, double& energyVal, bool) { energyVal = Esd; }
TriMesh& OcbLocalOptimizer::getResult(void) { return real ? real->getResult() : result; }
const Scaffold& OcbLocalOptimizer::getScaffold(void) const { return real ? real->getScaffold() : noScaffold; }

// ------------------------------------------------------------------------------------------------ the two passes
OcbBatchScope::OcbBatchScope(void) : outer(false)
{
    Batch& B = batch();
    if (enabled() && B.mode.load() == OFF) {
        outer = true;
        B.index.clear(); B.problems.clear();
        B.mode.store(RECORD);
        B.tMark = now();
    }
}
void OcbBatchScope::solve(void)
{
    Batch& B = batch();
    B.queries++;
    B.tPass1 += now() - B.tMark;
    solveAll(B);
    B.mode.store(REPLAY);
    B.tMark = now();
}
OcbBatchScope::~OcbBatchScope(void)
{
    if (outer) {
        Batch& B = batch();
        B.tPass2 += now() - B.tMark;
        B.mode.store(OFF);
        B.index.clear(); B.problems.clear();
    }
}

}  // namespace OptCuts
