// CudaLinSysSolver — OptCuts::LinSysSolver subclass over the optcuts_b200 C-ABI (include/optcuts_b200.h).
//
// Drop-in for OptCuts::EigenLibSolver (src/LinSysSolver/EigenLibSolver.{hpp,cpp}): same virtuals, called by
// the UNMODIFIED OptCuts::Optimizer exactly as before (Optimizer.cpp:173-183, 522-544, 563):
//     set_type -> set_pattern(vNeighbor, fixedVert) -> update_a(I,J,S) -> analyze_pattern -> factorize -> solve
// The sparse LDL^T is replaced by the device block-Jacobi PCG; the std::map-based pattern/assembly of the
// base class (LinSysSolver.hpp:37-159) is replaced by the device BSR pattern + triplet scatter.
//
// Constructing the object does no CUDA work (ocb_create is lazy): the reference creates one solver per
// Optimizer, including thousands of nested dense-mode optimizers that never call it.
// Errors: the C-ABI returns status codes; they are turned into std::runtime_error, which is what the
// reference's call sites catch (Optimizer.cpp:181-189, 538-554: dump the matrix, exit(-1)).
#ifndef CudaLinSysSolver_hpp
#define CudaLinSysSolver_hpp

#include "LinSysSolver.hpp"
#include "optcuts_b200.h"

#include <stdexcept>
#include <string>
#include <vector>

#include "CudaCoordinateHint.hpp"
namespace OptCuts {

// What the device holds for the Optimizer that owns this solver (written by shim/CudaOptimizer.cpp, the hooks of
// shim/CudaOptimizerHooks.hpp): copies of the host data last uploaded, so only what changed travels again.
struct CudaNewtonState {
    bool deviceResident = false;             // the Optimizer hooks drive this context: set_pattern / update_a are no-ops
    bool meshBound = false, uvBound = false, airBound = false, airUvBound = false;
    Eigen::MatrixXi F, Fa; Eigen::MatrixXd V, Va; Eigen::VectorXd triArea, airArea; Eigen::VectorXi l2g;
    std::vector<int> fixed, fixedAir;
    double surfaceArea = 0.0, wScaf = 0.0; long nBnd = 0;
    int lastIters = 0; long totalIters = 0, steps = 0;
};


template <typename vectorTypeI, typename vectorTypeS>
class CudaLinSysSolver : public LinSysSolver<vectorTypeI, vectorTypeS>
{
    typedef LinSysSolver<vectorTypeI, vectorTypeS> Base;

protected:
    ocb_ctx* ctx;
    double relTol;
    int maxIter, lastIters;
    double lastRelRes;
    long long downloadedVersion = -1;

    void check(int rc, const char* what) const {
        if (rc == OCB_ERR_NOT_CONVERGED) return;      // the solution is still written; Optimizer's line search decides
        if (rc < 0) throw std::runtime_error(std::string(what) + ": " + ocb_last_error(ctx));
    }

public:
    CudaNewtonState newton;
    ocb_ctx* context(void) const { return ctx; }

    CudaLinSysSolver(void) : ctx(NULL), relTol(1.0e-12), maxIter(0), lastIters(0), lastRelRes(0.0) {
        if (ocb_create(&ctx, 0) != OCB_OK) throw std::runtime_error("ocb_create failed");
        Base::numRows = 0;
    }
    ~CudaLinSysSolver(void) { ocb_destroy(ctx); }

    void set_type(int threadAmt, int _mtype, bool is_upper_half = false) {}   // SPD only, like EigenLibSolver::set_type

    // LinSysSolver::set_pattern (LinSysSolver.hpp:37-135): the vNeighbor sets go over as CSR, ascending like std::set
    void set_pattern(const std::vector<std::set<int>>& vNeighbor, const std::set<int>& fixedVert) {
        const int nV = static_cast<int>(vNeighbor.size());
        Base::numRows = nV * DIM;
        if (newton.deviceResident) return;        // the pattern is derived on the device from the uploaded element lists
        std::vector<int32_t> ptr(nV + 1, 0), idx, fixed(fixedVert.begin(), fixedVert.end());
        size_t tot = 0;
        for (const auto& s : vNeighbor) tot += s.size();
        idx.reserve(tot);
        for (int v = 0; v < nV; ++v) {
            for (int nb : vNeighbor[v]) idx.push_back(nb);
            ptr[v + 1] = static_cast<int32_t>(idx.size());
        }
        // geometry hint for the preconditioner hierarchy: the UVs CudaSymDirichletEnergy saw last (the mesh being optimised;
        // the air mesh's interior vertices follow it in the system and are placed by the library)
        {
            std::vector<double> xy; int n = 0;
            cudaCoordinateHint(false, n, xy);
            if (n > 0 && n <= nV) ocb_set_coordinate_hint(ctx, n, xy.data());
        }
        check(ocb_set_pattern(ctx, nV, ptr.data(), idx.data(), fixed.data(), static_cast<int>(fixed.size())), "ocb_set_pattern");
    }
    // EigenLibSolver::set_pattern(const SparseMatrix&) (EigenLibSolver.cpp:46-69): the matrix itself is handed over (pattern AND
    // values, SPD, no fixed vertices known).  No call site in the reference; implemented for the completeness of the plugin surface:
    // the vertex adjacency is read off the block structure (vertex = index / DIM), the upper-triangular entries go through the
    // same triplet route as update_a.
    void set_pattern(const Eigen::SparseMatrix<double>& mtr) {
        if (mtr.rows() != mtr.cols() || mtr.rows() % DIM) throw std::runtime_error("CudaLinSysSolver::set_pattern(SparseMatrix): not a square matrix of DIM x DIM blocks");
        const int nV = static_cast<int>(mtr.rows() / DIM);
        Base::numRows = nV * DIM;
        std::vector<std::set<int>> nb(nV);
        std::vector<int32_t> I, J; std::vector<double> S;
        for (int k = 0; k < mtr.outerSize(); ++k)
            for (Eigen::SparseMatrix<double>::InnerIterator it(mtr, k); it; ++it) {
                const int r = static_cast<int>(it.row()), c = static_cast<int>(it.col());
                if (r / DIM != c / DIM) { nb[r / DIM].insert(c / DIM); nb[c / DIM].insert(r / DIM); }
                if (r <= c) { I.push_back(r); J.push_back(c); S.push_back(it.value()); }
            }
        const bool wasResident = newton.deviceResident;
        newton.deviceResident = false;
        set_pattern(nb, std::set<int>());
        newton.deviceResident = wasResident;
        check(ocb_update_values_triplets(ctx, static_cast<int64_t>(S.size()), I.data(), J.data(), S.data()), "ocb_update_values_triplets");
    }

    // LinSysSolver::update_a (LinSysSolver.hpp:138-159): zero, accumulate triplets with i <= j
    void update_a(const vectorTypeI& II, const vectorTypeI& JJ, const vectorTypeS& SS) {
        if (newton.deviceResident) return;        // the element blocks were assembled on the device (ocb_hessian_assemble)
        check(ocb_update_values_triplets(ctx, static_cast<int64_t>(SS.size()), II.data(), JJ.data(), SS.data()), "ocb_update_values_triplets");
    }

    void analyze_pattern(void) {}                                   // nothing symbolic for PCG

    bool factorize(void) {                                          // EigenLibSolver::factorize: false / throw on a non-SPD matrix
        const int rc = ocb_factorize(ctx);
        if (rc == OCB_ERR_BREAKDOWN) return false;
        check(rc, "ocb_factorize");
        return true;
    }

    void solve(Eigen::VectorXd& rhs, Eigen::VectorXd& result) {
        result.resize(rhs.size());
        check(ocb_solve(ctx, rhs.data(), result.data(), relTol, maxIter, &lastIters, &lastRelRes), "ocb_solve");
    }

    void multiply(const Eigen::VectorXd& x, Eigen::VectorXd& Ax) {
        Ax.resize(x.size());
        check(ocb_multiply(ctx, x.data(), Ax.data()), "ocb_multiply");
    }

    // read-back in the reference layout (1-based upper-triangular CSR)
    void download(void) {
        int64_t sz[8];
        ocb_get_sizes(ctx, sz);
        Base::ia.resize(sz[5] + 1); Base::ja.resize(sz[6]); Base::a.resize(sz[6]);
        check(ocb_download_csr(ctx, Base::ia.data(), Base::ja.data(), Base::a.data()), "ocb_download_csr");
    }
    // LinSysSolver::coeffMtr (LinSysSolver.hpp:183-196): one entry of the upper-triangular CSR.  The device matrix is read back
    // once per assembly (a 64-bit change counter of the context tells whether the host copy is current), not once per call.
    double coeffMtr(int rowI, int colI) const {
        if (rowI > colI) std::swap(rowI, colI);
        CudaLinSysSolver* self = const_cast<CudaLinSysSolver*>(this);
        const long long ver = ocb_matrix_version(ctx);
        if (ver != self->downloadedVersion || Base::ia.size() == 0) { self->download(); self->downloadedVersion = ver; }
        for (int k = Base::ia[rowI] - 1; k < Base::ia[rowI + 1] - 1; ++k) if (Base::ja[k] - 1 == colI) return Base::a[k];
        return 0.0;
    }
    int getNumNonzeros(void) const { int64_t sz[8]; ocb_get_sizes(ctx, sz); return static_cast<int>(sz[6]); }

    void setTolerance(double tol, int maxIt) { relTol = tol; maxIter = maxIt; }
    int lastIterations(void) const { return lastIters; }
    double lastRelativeResidual(void) const { return lastRelRes; }
};

}  // namespace OptCuts
#endif
