// optcuts_b200 — internal declarations shared by the translation units of liboptcuts_b200.so.
// sm_100a only; fp64 everywhere; no tensor cores (nothing on this path is a dense contraction).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/optcuts_b200.h"
#include "ocb_mas.cuh"

namespace ocb {

// ---------------------------------------------------------------------------------------------
// growable device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n, cudaStream_t s, bool keep = false) {
        if (n <= cap) return cudaSuccess;
        size_t ncap = n + n / 4 + 64;
        T* q = nullptr;
        cudaError_t e = cudaMalloc((void**)&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(q, 0, ncap * sizeof(T), s);      // growth slack is defined memory (compute-sanitizer initcheck)
        if (e != cudaSuccess) { cudaFree(q); return e; }
        if (keep && p && cap) {
            e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) { cudaFree(q); return e; }
            cudaStreamSynchronize(s);
        }
        if (p) cudaFree(p);
        p = q; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// one set of triangles (mesh or air mesh) on the device, SoA, GLOBAL vertex ids
struct ElemSet {
    int n = 0;
    DevBuf<int32_t> v;      // 3 x n  (v0 | v1 | v2), stride = n
    DevBuf<double> rest;    // 8 x n  (TriMesh.hpp:44-52 order, see optcuts_b200.h), stride = n
    DevBuf<int32_t> slot;   // 9 x n  BSR block slot of (k,l), -1 if either vertex is fixed
    DevBuf<double> rec;     // 8 x n  one 64-byte record per triangle {v0 v1 v2 - | area areaSq e0 e1 d -} for the vertex-gather
                            //        kernels, which visit triangles in no particular order (2 sectors instead of 8)
    DevBuf<int32_t> vcSlot; // 4 x 3n per corner of the vertex->corner list: {element << 2 | corner, slot(k,0), slot(k,1), slot(k,2)}
};

// kernel-side view of an ElemSet
struct ElemView {
    int n;
    const int32_t* v0; const int32_t* v1; const int32_t* v2;
    const double* area; const double* areaSq; const double* e0; const double* e1; const double* d;
    const double* k0; const double* k1; const double* kd;
    const int32_t* slot;     // 9 x n or nullptr
    const double* rec;       // 8 x n records (see ElemSet)
    const int32_t* vcSlot;   // 4 per corner (see ElemSet)
    double surfaceArea;      // normaliser; weight = uniform ? 1 : area / surfaceArea
    int uniform;
    double scale;            // energyParam0 (mesh) or w_scaf/|Fa| (air): applied after projection
};

// kernel classes for the optional per-kernel CUDA-event profile (ocb_profile_*)
enum KernelClass {
    K_ENERGY = 0, K_GRADIENT, K_HESSIAN, K_PCG, K_STEP_BOUND, K_STEP_FORWARD, K_JACOBI_SETUP, K_SPMV,
    K_FEATURES, K_PATTERN, K_MISC, K_STENCILS, K_MAS_SETUP, K_HESSIAN_ROWS, K_COUNT
};

enum ScalarSlot {            // layout of the device/pinned scalar block
    S_E_MESH = 0, S_E_AIR, S_N_INVERTED, S_SQN_G, S_STEP_BOUND, S_PCG_ITERS, S_PCG_RELRES,
    S_PCG_STATUS, S_PCG_BNORM, S_MISC0, S_MISC1, S_MISC2, S_JACOBI_BAD, S_SQN_G_MESH, S_COUNT = 16
};

// MAS preconditioner: host-side hierarchy (built at pattern time) and its device mirror (ocb_mas.cu)
struct MasHost {
    bool enabled = false;
    int L = 0, Lloc = 0, grid = 0, topNodes = 0, maxLocalNodes = 0;
    struct Level {
        std::vector<int32_t> childBeg, parent, ctaBeg;   // see MasLevel
        std::vector<double> geom;                        // 4 per node
        std::vector<int32_t> rowPtr, colIdx;             // node adjacency (pattern of the Galerkin matrix A_l, 6x6 blocks)
    };
    std::vector<Level> lv;
    std::vector<float> vinfo;                            // 4 per row
};
struct MasDev {
    DevBuf<int32_t> ints, tabI; DevBuf<double> geom, val, rcCta, tabD, dense; DevBuf<float> inv, vinfo, cinv;
    std::vector<MasLevel> lv;
    std::vector<const int32_t*> lvRowPtr, lvColIdx; std::vector<double*> lvVal; std::vector<int> lvNnz;
    size_t valTotal = 0; int groupTotal = 0; bool denseAttr = false;
    MasView view = {};
};

}  // namespace ocb

struct ocb_ctx {
    int device = 0;
    bool inited = false;
    bool own_stream = false;
    bool streamGiven = false;                // ocb_set_stream was called: use that handle, NULL included (legacy default stream)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    int64_t launches = 0;
    int numSMs = 148;

    // sizes
    int nV = 0, nF = 0, nVa = 0, nFa = 0, nBnd = 0, nVtot = 0;
    double surfaceArea = 1.0, wScafOverFa = 0.0;
    bool haveUV = false, patternValid = false, matrixValid = false, precondValid = false, slotsValid = false;

    ocb::ElemSet mesh, air;
    std::vector<int32_t> hF, hFa;            // host copies of the element lists (INTERNAL global ids), 3 x n SoA
    std::vector<int32_t> hFuser;             // mesh element list in the caller's vertex numbering
    std::vector<int32_t> hPerm, hInv;        // caller vertex id -> internal id (locality order) and back; size nVtot
    ocb::DevBuf<int32_t> perm;
    std::vector<int32_t> hL2G;               // localVI2Global
    std::vector<uint8_t> hFixed;             // per global vertex
    ocb::DevBuf<int32_t> l2g;
    ocb::DevBuf<uint8_t> fixedMask;          // nVtot
    // vertex -> corner incidence (deterministic gather assembly: every vertex sums its incident element contributions in
    // ascending element order, which is the order the reference's serial loops add them in, SymDirichletEnergy.cpp:264-298,
    // LinSysSolver.hpp:147-157).  Entries are (element << 2 | corner).  Mesh list by INTERNAL vertex id; air list by
    // air-LOCAL vertex id; g2l: internal mesh vertex -> air-local id of the air vertex aliasing it (-1: none).
    ocb::DevBuf<int32_t> vcPtrM, vcIdxM, vcPtrA, vcIdxA, g2l;
    ocb::DevBuf<double> hel;                 // projected element Hessians, 6 upper 2x2 blocks x (nF + nFa), block-major

    // state vectors, all nSys doubles (interleaved u,v per global vertex)
    ocb::DevBuf<double> x, x0, g, p;
    // PCG work vectors
    ocb::DevBuf<double> pr, pz, pd, pd2, pAp, pb, px, minv;   // minv: 4 per block row; px: x in solver order
    // BSR(2x2), full symmetric storage
    int nnzb = 0;
    std::vector<int32_t> hRowPtr, hColIdx;   // pattern in INTERNAL vertex order (host only: download_csr, nnz)
    // the solver's own row order (Hilbert-curve order of the UVs, ocb_mas.cu): the device BSR, the PCG vectors and
    // the preconditioner live in it; identity when no UV was known at pattern time
    std::vector<int32_t> hRowOf, hVertOf;    // internal vertex -> solver row and back
    std::vector<int32_t> hSRowPtr, hSColIdx; // the device pattern (solver order)
    std::vector<int32_t> hBlkMap;            // internal-order block -> device block
    ocb::DevBuf<int32_t> rowOf, vertOf, userRow;   // device: internal vertex -> row, row -> internal vertex, caller vertex -> row
    ocb::MasHost masH; ocb::MasDev masD;
    int planGrid = 0;                        // persistent-CTA count the hierarchy was built for
    // what the last ocb_gradient / fused gradient pass left (ocb_newton_step_ex(OCB_STEP_REUSE_GRADIENT) continues from it)
    bool gradValid = false; double gradP0 = 0.0, gradSqn = 0.0, gradEMesh = 0.0, gradEAir = 0.0, gradSqnMesh = 0.0;
    bool pcgPlainNorm = true;                // false (ocb_set_option("pcg_scaled_norm", 1)): stop on the D-scaled norm instead of ||r|| / ||b||
    bool scaleSystem = false;                // ocb_newton_step_ex scales the assembled system symmetrically (option "scale_system")
    bool systemScaled = false;               // c->val currently holds S A S (only between the assembly and the end of ocb_newton_step_ex)
    ocb::DevBuf<double> rowScale;            // S: 2 per solver row
    bool tolerateIndefinite = false;         // ocb_newton_step_ex: a matrix that is SPD only up to rounding is handled, not reported
    bool deferFactorCheck = false;           // ocb_newton_step: the block-Jacobi verdict is read together with the PCG status
    int64_t precondFallbacks = 0;            // solves repeated with block-Jacobi after the two-level preconditioner failed
    std::vector<int32_t> hStamp;             // scratch of the pattern builders
    std::vector<int32_t> hMeshAdjPtr, hMeshAdj;   // de-duplicated vertex adjacency of the MESH (internal ids), kept until the next ocb_set_mesh
    bool meshAdjValid = false;
    void* directHandle = nullptr;             // cuSOLVER handle of the dense safety net (ocb_direct.cu), created at first use
    void* directBlas = nullptr;               // cuBLAS handle of the block-tridiagonal path
    ocb::DevBuf<double> directA, directB, directWork; ocb::DevBuf<int> directInfo; ocb::DevBuf<int32_t> directI; ocb::DevBuf<long long> directL;
    int lastDirectLifts = 0;
    int directSkip = 0, directBackoff = 0;    // Newton solves that still go to the safety net first / length of the current back-off
    bool forceDirect = false;                 // option force_direct: ocb_solve uses the direct safety net instead of CG (tests)
    long long directSolves = 0;               // solves that went through the dense Cholesky
    bool masEquilibrate = true;               // option mas_equilibrate (OCB_MAS_EQUILIBRATE=0 turns it off): diagonal equilibration inside the group / coarse inversions
    long long matrixVersion = 0;              // bumped whenever the device matrix (pattern or values) changes: ocb_matrix_version
    std::vector<double> hHint;               // ocb_set_coordinate_hint: 2 per vertex (interleaved), caller numbering
    std::vector<double> hXY;                 // host mirror of x (INTERNAL numbering) as last written by ocb_set_uv: spares the
    bool hXYMesh = false, hXYAir = false;    // row-order code its download; a device-side update of x invalidates it
    std::vector<double> hCoords;             // positions of all nVtot vertices for a solver-only context (INTERNAL numbering)
    ocb::DevBuf<int32_t> rowPtr, colIdx;
    ocb::DevBuf<double> val;                 // 4 per block, row-major
    // reductions
    ocb::DevBuf<double> partials;            // per-block partial sums
    ocb::DevBuf<unsigned> sync;              // tickets / grid barrier words
    double* dScal = nullptr;                 // S_COUNT device scalars
    double* hScal = nullptr;                 // pinned mirror
    // scratch for uploads
    ocb::DevBuf<double> scratchD, scratchV, stD;
    ocb::DevBuf<int32_t> stI;
    unsigned char* stPinned = nullptr; size_t stPinnedCap = 0;   // pinned staging arena of ocb_stencil_newton_step
    ocb::DevBuf<int32_t> scratchI;
    int pcgGrid = 0, pcgBlock = 0;
    ocb::DevBuf<int32_t> pcgHalo;            // cluster-mode PCG: halo slot table (see spmv_fused_ell)
    size_t pcgSmemAttr = 0;
    int clusterOk = -1;                      // -1 unknown, 0 a 16-CTA cluster cannot be scheduled, 1 ok
    ocb::DevBuf<double> xSaved;              // ocb_save_uv / ocb_restore_uv snapshot
    int xSavedN = 0;

    // per-kernel-class event profile
    bool prof = false;
    struct ProfRec { int cls; cudaEvent_t a, b; };
    std::vector<ProfRec> profRecs;
    std::vector<cudaEvent_t> profPool;
    double profMs[ocb::K_COUNT] = {0};
    int64_t profCnt[ocb::K_COUNT] = {0};

    int nSys() const { return 2 * nVtot; }
};

namespace ocb {

int set_err(ocb_ctx* c, int code, const char* what);
int cuda_fail(ocb_ctx* c, cudaError_t e, const char* where);
#define OCB_CUDA(c, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return ocb::cuda_fail((c), _e, #call); } while (0)
#define OCB_TRY(call) do { int _r = (call); if (_r < 0) return _r; } while (0)

int ensure_init(ocb_ctx* c);
// OCB_HOST_TIMING=1: wall-clock per host-side section, printed by ocb_destroy (e2e tuning aid)
struct HostTimer {
    const char* name; double t0;
    explicit HostTimer(const char* n);
    ~HostTimer();
};
void host_timing_report();
// RAII: brackets the launches of one kernel class with CUDA events when profiling is enabled
struct ProfScope {
    ocb_ctx* c; int cls; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(ocb_ctx* ctx, int k);
    ~ProfScope();
};
ElemView view_of(const ocb_ctx* c, const ElemSet& s, bool isAir, double scale, int uniform);
int fetch_scalars(ocb_ctx* c);   // D2H of the scalar block + stream sync

// launchers implemented in ocb_kernels.cu / ocb_pcg.cu
int launch_energy(ocb_ctx* c, double p0, bool stepped, double alpha, bool alphaFromStepBound = false);   // alphaFromStepBound: alpha * scal[S_STEP_BOUND], read on the device
int launch_energy_per_elem(ocb_ctx* c, int uniform, double* d_out);
int launch_energy_one_elem(ocb_ctx* c, int t, int uniform, double* d_out);
int launch_gradient(ocb_ctx* c, double p0);
int launch_sqnorm(ocb_ctx* c, const double* v, int n, int slot);
int launch_g2l(ocb_ctx* c);                          // mesh-vertex -> air-local alias table after a new air mesh
int launch_build_slots(ocb_ctx* c);
int launch_hessian(ocb_ctx* c, double p0);
int launch_hessian_blocks(ocb_ctx* c, int uniform, double* d_out36);
int launch_dense_hessian(ocb_ctx* c, const double* d_blocks36, const int32_t* d_inv, double* d_out);
int launch_step_bound(ocb_ctx* c, const double* d_dir, double alpha0);
int launch_step_forward(ocb_ctx* c, double alpha);
int launch_triplet_scatter(ocb_ctx* c, int64_t nT, const int32_t* dI, const int32_t* dJ, const double* dS);
int launch_set_uv(ocb_ctx* c, const double* dV, const double* dVa);
int launch_get_uv(ocb_ctx* c, double* dV, double* dVa);
int launch_permute_vec(ocb_ctx* c, const double* in, double* out, bool toInternal);
int launch_permute_scalar(ocb_ctx* c, int n, const double* in, double* out);
int launch_rest_features(ocb_ctx* c, int nV, int nF, const double* dVrest, const int32_t* dF, double thres, double* dRest8);
int launch_seam(ocb_ctx* c, int nCoh, const int32_t* dCoh, const double* dLen, const int32_t* dBnd, double thresLen, int triSoup);
int launch_divgrad(ocb_ctx* c, double* d_out);
struct StencilHost {     // device pointers of one uploaded stencil batch
    int nStencil; const int32_t* vertPtr; const int32_t* triPtr; const double* Vrest; const double* UV; const int32_t* F;
    const uint8_t* isFree; const double* scoreScale; const double* scoreOffset; int maxIter; double relGL2Tol;
    double* Einit; double* Efinal; double* UVout; int32_t* iters; double* score; int32_t* status; int* argmax;
};
int launch_stencils(ocb_ctx* c, const StencilHost& h);
struct StencilStepHost {  // device pointers of one uploaded batch of bijective stencils (one Newton iteration each)
    int nStencil; const int32_t* vertPtr; const int32_t* triPtr; const int32_t* nVm; const int32_t* nTm; const double* Vrest; const double* UV;
    const int32_t* F; const uint8_t* isFree; const double* areaThres; const double* targetGRes; double wScaf;
    double* UVout; double* out6; int32_t* result;
};
int launch_stencil_step(ocb_ctx* c, const StencilStepHost& h);
int launch_spmv(ocb_ctx* c, const double* dx, double* dy);
static constexpr int kDirectMaxDof = 40000;               // largest system (scalar unknowns) the dense safety net takes: 12.8 GB of fp64
bool direct_solver_available(const ocb_ctx* c);
int launch_direct_solve(ocb_ctx* c, const double* d_rhs, bool negate, int* liftsUsed);   // 0 solved, 1 not available / not factorisable, < 0 error
void direct_release(ocb_ctx* c);
int direct_level_blocks_host(int n, const int32_t* rowPtr, const int32_t* colIdx, int target, int32_t* pos, int32_t* blkOf, int32_t* blkBeg);
int launch_diag_shift(ocb_ctx* c, double delta);          // diagonal entries *= (1 + delta)
int launch_scale_system(ocb_ctx* c);                       // val <- S val S, S = diag^-1/2; sets c->systemScaled
int launch_jacobi_setup(ocb_ctx* c, bool check = true);   // check = false: no host round trip, the verdict stays in scal[S_JACOBI_BAD]
int launch_pcg(ocb_ctx* c, const double* d_rhs, bool negate_rhs, double rel_tol, int max_it, bool allowMas = true);
int pcg_plan_grid(ocb_ctx* c, int nRows);            // CTA count launch_pcg will use for this system size
int mas_build_hierarchy(ocb_ctx* c, const double* xy, int grid);
int mas_install(ocb_ctx* c);
int launch_mas_setup(ocb_ctx* c);
int launch_gather_rows(ocb_ctx* c, const double* in_internal, double* out_rows, bool toRows);

}  // namespace ocb
