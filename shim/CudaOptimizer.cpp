// Implementation of the Optimizer hooks (shim/CudaOptimizerHooks.hpp): the reference's global optimizer with its Newton
// iteration resident on the B200 through the C-ABI of liboptcuts_b200.so.
#include "CudaOptimizerHooks.hpp"
#include "CudaLinSysSolver.hpp"
#include "Timer.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>

extern Timer timer, timer_step;              // main.cpp:104

namespace OptCuts {

typedef CudaLinSysSolver<Eigen::VectorXi, Eigen::VectorXd> Solver;

static bool enabled()
{
    static const bool on = []() { const char* e = std::getenv("OCB_DEVICE_NEWTON"); return !(e && std::atoi(e) == 0); }();
    return on;
}
static double pcgTol() { static const double t = []() { const char* e = std::getenv("OCB_PCG_TOL"); return e ? std::atof(e) : 1.0e-12; }(); return t; }

static void die(ocb_ctx* ctx, const char* what, int rc)
{
    // the reference's own failure convention: report and exit(-1) (Optimizer.cpp:186-189, 548-554)
    std::fprintf(stderr, "optcuts_b200: %s failed (%d): %s\n", what, rc, ocb_last_error(ctx));
    std::exit(-1);
}
#define OCB_DO(ctx, call, what) do { const int rc_ = (call); if (rc_ < 0 && rc_ != OCB_ERR_INVERTED && rc_ != OCB_ERR_NOT_CONVERGED) die(ctx, what, rc_); } while (0)

template <typename M> static bool sameBits(const M& a, const M& b)
{
    return a.rows() == b.rows() && a.cols() == b.cols() && (a.size() == 0 || std::memcmp(a.data(), b.data(), sizeof(typename M::Scalar) * a.size()) == 0);
}

static void rest8Of(const TriMesh& m, Eigen::MatrixXd& rest8)
{
    const int nF = static_cast<int>(m.F.rows());
    rest8.resize(nF, 8);                          // column k = feature k == 8 x nF SoA in memory (TriMesh.hpp:44-52)
    rest8.col(0) = m.triArea; rest8.col(1) = m.triAreaSq; rest8.col(2) = m.e0SqLen; rest8.col(3) = m.e1SqLen; rest8.col(4) = m.e0dote1;
    rest8.col(5) = m.e0SqLen_div_dbAreaSq; rest8.col(6) = m.e1SqLen_div_dbAreaSq; rest8.col(7) = m.e0dote1_div_dbAreaSq;
}

// make the device hold what the Optimizer's `result` / `scaffold` hold now; uploads only what changed
static Solver* bind(OcbOptView& v)
{
    if (!enabled()) return NULL;
    Solver* S = dynamic_cast<Solver*>(v.solver);
    if (!S) return NULL;
    CudaNewtonState& st = S->newton;
    ocb_ctx* ctx = S->context();
    const TriMesh& m = v.result;
    const int nF = static_cast<int>(m.F.rows()), nV = static_cast<int>(m.V.rows());
    std::vector<int> fixed(m.fixedVert.begin(), m.fixedVert.end());
    const bool sameMesh = st.meshBound && st.surfaceArea == m.surfaceArea && st.fixed == fixed && st.V.rows() == nV &&
                          sameBits(st.F, m.F) && sameBits(st.triArea, m.triArea);
    if (!sameMesh) {
        Eigen::MatrixXd rest8; rest8Of(m, rest8);
        OCB_DO(ctx, ocb_set_mesh(ctx, nV, nF, m.F.data(), rest8.data(), m.surfaceArea, fixed.data(), static_cast<int>(fixed.size())), "ocb_set_mesh");
        st.F = m.F; st.triArea = m.triArea; st.surfaceArea = m.surfaceArea; st.fixed = fixed;
        st.meshBound = true; st.uvBound = false; st.airBound = false;
    }
    if (!st.uvBound || !sameBits(st.V, m.V)) {
        OCB_DO(ctx, ocb_set_uv(ctx, m.V.data(), NULL), "ocb_set_uv");
        st.V = m.V; st.uvBound = true;
    }
    if (v.scaffolding) {
        const Scaffold& sc = v.scaffold;
        const TriMesh& am = sc.airMesh;
        const int nFa = static_cast<int>(am.F.rows()), nVa = static_cast<int>(am.V.rows());
        const double w = v.w_scaf / nFa;                           // Optimizer.cpp:776, 795, 839
        std::vector<int> fixedAir(am.fixedVert.begin(), am.fixedVert.end());
        const bool sameAir = st.airBound && st.wScaf == w && st.fixedAir == fixedAir && st.nBnd == sc.bnd.size() && st.Va.rows() == nVa &&
                             sameBits(st.Fa, am.F) && sameBits(st.l2g, sc.localVI2Global) && sameBits(st.airArea, am.triArea);
        if (!sameAir) {
            Eigen::MatrixXd rest8; rest8Of(am, rest8);
            OCB_DO(ctx, ocb_set_air(ctx, nVa, nFa, am.F.data(), rest8.data(), sc.localVI2Global.data(), static_cast<int>(sc.bnd.size()),
                                    fixedAir.data(), static_cast<int>(fixedAir.size()), w), "ocb_set_air");
            st.Fa = am.F; st.l2g = sc.localVI2Global; st.airArea = am.triArea; st.nBnd = sc.bnd.size(); st.wScaf = w; st.fixedAir = fixedAir;
            st.airBound = true; st.airUvBound = false;
        }
        if (!st.airUvBound || !sameBits(st.Va, am.V)) {
            OCB_DO(ctx, ocb_set_uv(ctx, NULL, am.V.data()), "ocb_set_uv(air)");
            st.Va = am.V; st.airUvBound = true;
        }
    } else if (st.airBound) {
        OCB_DO(ctx, ocb_set_air(ctx, 0, 0, NULL, NULL, NULL, 0, NULL, 0, 0.0), "ocb_set_air(none)");
        st.airBound = false;
    }
    st.deviceResident = true;                    // from here on the solver's set_pattern / update_a are no-ops
    return S;
}

bool ocbDeviceResident(LinSysSolver<Eigen::VectorXi, Eigen::VectorXd>* solver)
{
    if (!enabled()) return false;
    Solver* S = dynamic_cast<Solver*>(solver);
    return S && S->newton.deviceResident;
}

bool ocbHookEnergy(OcbOptView& v, double& energyVal, bool excludeScaffold)
{
    Solver* S = bind(v);
    if (!S) return false;
    double tot = 0.0, esd = 0.0, escaf = 0.0;
    OCB_DO(S->context(), ocb_energy(S->context(), v.energyParams[0], &tot, &esd, &escaf), "ocb_energy");
    v.energyVal_ET[0] = esd;
    v.energyVal_scaffold = (v.scaffolding && !excludeScaffold) ? escaf : 0.0;
    energyVal = v.energyParams[0] * esd + v.energyVal_scaffold;                 // Optimizer.cpp:766-781
    return true;
}

bool ocbHookGradient(OcbOptView& v, Eigen::VectorXd& gradient, bool excludeScaffold)
{
    Solver* S = bind(v);
    if (!S) return false;
    ocb_ctx* ctx = S->context();
    int64_t sz[8];
    ocb_get_sizes(ctx, sz);
    gradient.resize(sz[5]);
    double sqn = 0.0, info[4];
    OCB_DO(ctx, ocb_gradient(ctx, v.energyParams[0], gradient.data(), &sqn), "ocb_gradient");
    OCB_DO(ctx, ocb_gradient_info(ctx, info), "ocb_gradient_info");
    // gradient_ET is only read for its norm (Optimizer::writeGradL2NormToFile, Optimizer.cpp:729-740)
    v.gradient_ET[0].resize(1);
    v.gradient_ET[0][0] = std::sqrt(info[1]);
    return true;
}

bool ocbHookHessian(OcbOptView& v)
{
    Solver* S = bind(v);
    if (!S) return false;
    ocb_ctx* ctx = S->context();
    if (!v.mute) timer_step.start(0);
    OCB_DO(ctx, ocb_hessian_assemble(ctx, v.energyParams[0]), "ocb_hessian_assemble");
    if (!v.mute) { ocb_synchronize(ctx); timer_step.stop(); }
    return true;
}

bool ocbHookStep(OcbOptView& v, bool& stopped)
{
    Solver* S = bind(v);
    if (!S) return false;
    ocb_ctx* ctx = S->context();
    CudaNewtonState& st = S->newton;
    ocb_newton_result r;
    const int flags = OCB_STEP_REUSE_GRADIENT | OCB_STEP_SKIP_CONVERGENCE_TEST | (v.fractureInitiated ? OCB_STEP_REUSE_MATRIX : 0);
    if (!v.mute) timer_step.start(4);
    OCB_DO(ctx, ocb_newton_step_ex(ctx, v.energyParams[0], v.targetGRes, pcgTol(), 0, v.allowEDecRelTol ? 1 : 0, flags, &r), "ocb_newton_step_ex");
    // the host program needs the new UVs: the scaffold is re-triangulated from them and the topology step edits them
    OCB_DO(ctx, ocb_get_uv(ctx, v.result.V.data(), v.scaffolding ? v.scaffold.airMesh.V.data() : NULL), "ocb_get_uv");
    if (!v.mute) timer_step.stop();
    st.V = v.result.V;
    if (v.scaffolding) st.Va = v.scaffold.airMesh.V;
    v.lastEDec = r.lastEDec;                                                     // Optimizer.cpp:631-639
    v.lastEnergyVal = r.E_new;
    v.energyVal_scaffold = r.E_scaf_new;
    v.energyVal_ET[0] = r.E_sd_new;
    st.lastIters = r.pcg_iters; st.totalIters += r.pcg_iters; st.steps++;
    if (!v.mute) std::printf("stepSize: %g -> %g (device: %d CG iterations, rel. residual %.2e%s%s)\n", r.alpha_init, r.alpha, r.pcg_iters, r.pcg_rel_res,
                             r.pcg_status == OCB_ERR_BREAKDOWN ? ", truncated at non-positive curvature" : (r.pcg_status == OCB_ERR_NOT_CONVERGED ? ", iteration cap" : ""),
                             r.reserved ? ", diagonal lifted after a breakdown" : "");
    stopped = r.stopped != 0;
    return true;
}

}  // namespace OptCuts
