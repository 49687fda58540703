// TEST INFRASTRUCTURE — not on the product path.
//
// Per-iteration dump hook for the reference program.  oracle/Makefile generates
// _ref/main_probe.cpp = reference main.cpp + `#include "probe_hook.hpp"` after the global
// timers (main.cpp:104) + one call `oracle_probe_iteration();` after the per-iteration solve
// (main.cpp:116).  Nothing else of the reference is touched.
//
//   ORACLE_TRACE=<file>        one text line per Newton iteration, every double at %.17g
//                              (the reference's own logs keep 6 digits: Optimizer.cpp:713-723)
//   ORACLE_DUMP_DIR=<dir>      + ORACLE_DUMP_ITERS="1,2,40"  full binary state after those iterations
//   ORACLE_MAX_ITERS=<n>       exit(0) after n Newton iterations (bounded CPU-baseline samples)
//
// State file = sequence of records {char name[32]; char type ('d'|'i'); int64 rows, cols; data col-major}.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

// a crash in a test run leaves a backtrace on stderr (the probe binaries are linked -rdynamic)
static void oracle_segv_handler(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "\n[oracle probe] fatal signal, backtrace:\n";
    if (write(2, msg, sizeof(msg) - 1) < 0) {}
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}
static const bool oracle_segv_installed = []() { signal(SIGSEGV, oracle_segv_handler); signal(SIGABRT, oracle_segv_handler); return true; }();

static void oracle_write_rec(FILE* f, const char* name, char type, long rows, long cols, const void* data)
{
    char nm[32]; std::memset(nm, 0, sizeof nm); std::strncpy(nm, name, 31);
    std::fwrite(nm, 1, 32, f); std::fwrite(&type, 1, 1, f);
    long long rc[2] = {rows, cols}; std::fwrite(rc, sizeof(long long), 2, f);
    std::fwrite(data, type == 'd' ? 8 : 4, (size_t)(rows * cols), f);
}
static void oracle_write_mat(FILE* f, const char* name, const Eigen::MatrixXd& m) { oracle_write_rec(f, name, 'd', m.rows(), m.cols(), m.data()); }
static void oracle_write_mat(FILE* f, const char* name, const Eigen::MatrixXi& m) { oracle_write_rec(f, name, 'i', m.rows(), m.cols(), m.data()); }
static void oracle_write_vec(FILE* f, const char* name, const Eigen::VectorXd& m) { oracle_write_rec(f, name, 'd', m.size(), 1, m.data()); }
static void oracle_write_vec(FILE* f, const char* name, const Eigen::VectorXi& m) { oracle_write_rec(f, name, 'i', m.size(), 1, m.data()); }

static void oracle_dump_state(const std::string& path)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return;
    const OptCuts::TriMesh& r = optimizer->getResult();
    oracle_write_mat(f, "V_rest", r.V_rest); oracle_write_mat(f, "F", r.F); oracle_write_mat(f, "V", r.V);
    oracle_write_mat(f, "cohE", r.cohE); oracle_write_vec(f, "boundaryEdge", r.boundaryEdge); oracle_write_vec(f, "edgeLen", r.edgeLen);
    oracle_write_vec(f, "vertWeight", r.vertWeight);
    Eigen::VectorXi fx(r.fixedVert.size()); { int k = 0; for (int v : r.fixedVert) fx[k++] = v; }
    oracle_write_vec(f, "fixedVert", fx);
    double sc[8] = {r.surfaceArea, r.avgEdgeLen, r.virtualRadius, r.initSeamLen, optimizer->getLastEnergyVal(),
                    optimizer->getLastEnergyVal(true), energyParams[0], (double)iterNum};
    oracle_write_rec(f, "scalars", 'd', 8, 1, sc);
    if (optimizer->isScaffolding()) {
        const OptCuts::Scaffold& s = optimizer->getScaffold();
        const OptCuts::TriMesh& am = s.airMesh;
        oracle_write_mat(f, "air_V", am.V); oracle_write_mat(f, "air_F", am.F);
        oracle_write_vec(f, "air_bnd", s.bnd); oracle_write_vec(f, "air_localVI2Global", s.localVI2Global);
        Eigen::VectorXi afx(am.fixedVert.size()); { int k = 0; for (int v : am.fixedVert) afx[k++] = v; }
        oracle_write_vec(f, "air_fixedVert", afx);
        double asc[4] = {am.surfaceArea, am.avgEdgeLen, am.areaThres_AM, (double)s.wholeMeshSize};
        oracle_write_rec(f, "air_scalars", 'd', 4, 1, asc);
        oracle_write_vec(f, "air_triArea", am.triArea);
    }
    std::fclose(f);
}

// FNV-1a over the raw bytes: identical mesh connectivity <=> identical hash, iteration by iteration, so two traces with
// equal Fhash / cohEhash columns went through the SAME sequence of topology operations (type AND path of every split / merge)
static unsigned long long oracle_fnv(const void* data, size_t bytes)
{
    const unsigned char* p = static_cast<const unsigned char*>(data);
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

static void oracle_probe_iteration(void)
{
    static FILE* trace = NULL;
    static bool init = false;
    static std::set<int> dumpIters;
    static std::string dumpDir;
    static int maxIters = -1;
    if (!init) {
        init = true;
        if (const char* t = std::getenv("ORACLE_TRACE")) trace = std::fopen(t, "w");
        if (const char* d = std::getenv("ORACLE_DUMP_DIR")) dumpDir = d;
        if (const char* s = std::getenv("ORACLE_DUMP_ITERS")) {
            std::string str(s); size_t pos = 0;
            while (pos < str.size()) { size_t c = str.find(',', pos); if (c == std::string::npos) c = str.size();
                dumpIters.insert(std::atoi(str.substr(pos, c - pos).c_str())); pos = c + 1; }
        }
        if (const char* m = std::getenv("ORACLE_MAX_ITERS")) maxIters = std::atoi(m);
    }
    if (trace) {
        const OptCuts::TriMesh& r = optimizer->getResult();
        double E_se; r.computeSeamSparsity(E_se, !fractureMode); E_se /= r.virtualRadius;
        const bool sc = optimizer->isScaffolding();
        std::fprintf(trace, "it=%d conv=%d topo=%d Fhash=%016llx cohEhash=%016llx F=%d V=%d cohE=%d amF=%d amV=%d bnd=%d E=%.17g Enoscaf=%.17g Ese=%.17g p0=%.17g\n",
                     iterNum, converged, optimizer->getTopoIter(), oracle_fnv(r.F.data(), sizeof(int) * r.F.size()),
                     oracle_fnv(r.cohE.data(), sizeof(int) * r.cohE.size()), (int)r.F.rows(), (int)r.V.rows(), (int)r.cohE.rows(),
                     sc ? (int)optimizer->getAirMesh().F.rows() : 0, sc ? (int)optimizer->getAirMesh().V.rows() : 0,
                     sc ? (int)optimizer->getScaffold().bnd.size() : 0,
                     optimizer->getLastEnergyVal(), optimizer->getLastEnergyVal(true), E_se, energyParams[0]);
        std::fflush(trace);
    }
    if (!dumpDir.empty() && dumpIters.count(iterNum)) {
        char nm[64]; std::snprintf(nm, sizeof nm, "/state_%06d.bin", iterNum);
        oracle_dump_state(dumpDir + nm);
    }
    if (maxIters >= 0 && iterNum >= maxIters) {
        double t4[4], s9[9];
        for (int i = 0; i < 4; ++i) t4[i] = timer.timing(i);
        for (int i = 0; i < 9; ++i) s9[i] = timer_step.timing(i);
        std::printf("ORACLE_TIMERS iters=%d topo=%.6f desc=%.6f scaf=%.6f enUp=%.6f mtrComp=%.6f mtrAssem=%.6f symFac=%.6f numFac=%.6f backSolve=%.6f lineSearch=%.6f\n",
                    iterNum, t4[0], t4[1], t4[2], t4[3], s9[0], s9[1], s9[2], s9[3], s9[4], s9[5]);
        std::fflush(stdout);
        std::exit(0);
    }
}
