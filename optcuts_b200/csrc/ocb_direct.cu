// optcuts_b200 — safety net of the linear solve: dense Cholesky of the assembled matrix on the device.
//
// The reference solves every Newton system with a sparse LDL^T (EigenLibSolver.cpp:71-107), which returns a usable direction
// even when the matrix spans 30 orders of magnitude (the Tutte starts of 9 of the 71 benchmark meshes: ||g||^2 up to 1e51).
// Conjugate gradients cannot: on those systems the two-level preconditioner comes out indefinite, block-Jacobi CG stalls at
// its iteration cap, and the truncated iterate is a poor direction (cat_noUV: E 2.6e9 -> 3.5e8 after 30 iterations where the
// reference reaches 5.1e5, profiles/r2_cat_noUV_before_direct.txt).  For systems of at most kDirectMaxDof unknowns such a solve
// is repeated here directly: the BSR matrix is expanded to a dense lower triangle, factorised with cuSOLVER's potrf and solved
// with potrs (library calls: a plain dense factorisation, not a kernel of this path; ~n^3/3 flops, 15k unknowns ~ 50 ms, i.e.
// 25x the cost of a healthy PCG solve and several times cheaper than a stalled one).  cuSOLVER is loaded with dlopen at first
// use, so the library itself has no link-time dependency on it; when it is missing, or the system is larger, the block-Jacobi
// retry stays the fallback.  A matrix that is indefinite by rounding makes potrf fail: the diagonal is then lifted relatively
// (1e-8, 1e-5, 1e-2) as after a CG breakdown.
#include "ocb_internal.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#define KCHECK(c) do { (c)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return cuda_fail((c), _e, __func__); } while (0)

namespace ocb {

namespace {

typedef void* cusolverDnHandle_t;
typedef int (*fnCreate)(cusolverDnHandle_t*);
typedef int (*fnSetStream)(cusolverDnHandle_t, cudaStream_t);
typedef int (*fnPotrfBuf)(cusolverDnHandle_t, int, int, double*, int, int*);
typedef int (*fnPotrf)(cusolverDnHandle_t, int, int, double*, int, double*, int, int*);
typedef int (*fnPotrs)(cusolverDnHandle_t, int, int, int, const double*, int, double*, int, int*);

struct Cusolver {
    bool tried = false, ok = false;
    void* lib = nullptr;
    fnCreate create = nullptr; fnSetStream setStream = nullptr; fnPotrfBuf potrfBuf = nullptr; fnPotrf potrf = nullptr; fnPotrs potrs = nullptr;
};
Cusolver& cusolver()
{
    static Cusolver S;
    if (S.tried) return S;
    S.tried = true;
    const char* off = getenv("OCB_NO_DIRECT");
    if (off && atoi(off)) return S;
    const char* names[] = {"libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so.12", "libcusolver.so"};
    for (const char* nme : names) { S.lib = dlopen(nme, RTLD_NOW | RTLD_LOCAL); if (S.lib) break; }
    if (!S.lib) return S;
    S.create = (fnCreate)dlsym(S.lib, "cusolverDnCreate"); S.setStream = (fnSetStream)dlsym(S.lib, "cusolverDnSetStream");
    S.potrfBuf = (fnPotrfBuf)dlsym(S.lib, "cusolverDnDpotrf_bufferSize"); S.potrf = (fnPotrf)dlsym(S.lib, "cusolverDnDpotrf");
    S.potrs = (fnPotrs)dlsym(S.lib, "cusolverDnDpotrs");
    S.ok = S.create && S.setStream && S.potrfBuf && S.potrf && S.potrs;
    return S;
}
constexpr int kFillLower = 0;                 // CUBLAS_FILL_MODE_LOWER

// dense column-major lower triangle (both triangles are written: the matrix is stored in full) from the BSR blocks;
// lift: diagonal entries *= 1 + lift
__global__ void __launch_bounds__(256)
dense_from_bsr_kernel(int nRows, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, const double* __restrict__ val,
                      double* __restrict__ D, size_t ld, double lift)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256)
        for (int b = rowPtr[row]; b < rowPtr[row + 1]; ++b) {
            const int col = colIdx[b];
            const double* v = val + 4 * (size_t)b;             // [a00 a01 a10 a11] of block (row, col)
            const double f = col == row ? 1.0 + lift : 1.0;
            D[(size_t)(2 * col) * ld + 2 * row] = v[0] * f;
            D[(size_t)(2 * col + 1) * ld + 2 * row] = v[1];
            D[(size_t)(2 * col) * ld + 2 * row + 1] = v[2];
            D[(size_t)(2 * col + 1) * ld + 2 * row + 1] = v[3] * f;
        }
}
// b (solver order) = +-rhs (internal vertex order)
__global__ void __launch_bounds__(256)
direct_gather_rhs_kernel(int nRows, const int32_t* __restrict__ vertOf, const double* __restrict__ rhs, int negate, double* __restrict__ b)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        const size_t src = vertOf ? (size_t)vertOf[row] : (size_t)row;
        const double b0 = rhs[2 * src], b1 = rhs[2 * src + 1];
        b[2 * (size_t)row] = negate ? -b0 : b0; b[2 * (size_t)row + 1] = negate ? -b1 : b1;
    }
}
__global__ void __launch_bounds__(256)
direct_scatter_x_kernel(int nRows, const int32_t* __restrict__ vertOf, const double* __restrict__ x, double* __restrict__ xOut, double* __restrict__ scal, int devInfo0,
                        const int* __restrict__ devInfo)
{
    const bool ok = devInfo[0] == 0 && devInfo[1] == 0;
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        const size_t dst = vertOf ? (size_t)vertOf[row] : (size_t)row;
        if (ok) { xOut[2 * dst] = x[2 * (size_t)row]; xOut[2 * dst + 1] = x[2 * (size_t)row + 1]; }
    }
    (void)scal; (void)devInfo0;
}

}  // namespace

bool direct_solver_available(const ocb_ctx* c)
{
    return !c->systemScaled && 2 * (size_t)c->nVtot <= (size_t)kDirectMaxDof && cusolver().ok;
}

// x (internal order, c->p) = A^-1 (+-rhs).  Returns 0 on success, 1 when the factorisation failed for every lift (the caller
// keeps its iterative fallback), < 0 on a CUDA / library error.
int launch_direct_solve(ocb_ctx* c, const double* d_rhs, bool negate, int* liftsUsed)
{
    Cusolver& S = cusolver();
    if (!S.ok) return 1;
    ProfScope prof(c, K_PCG);
    const int nRows = c->nVtot, n = 2 * nRows;
    if (!c->directHandle) {
        if (S.create(&c->directHandle) != 0) { c->directHandle = nullptr; return 1; }
    }
    if (S.setStream(c->directHandle, c->stream) != 0) return 1;
    OCB_CUDA(c, c->directA.reserve((size_t)n * n + 8, c->stream));
    OCB_CUDA(c, c->directB.reserve((size_t)n + 8, c->stream));
    OCB_CUDA(c, c->directInfo.reserve(4, c->stream));
    int lwork = 0;
    if (S.potrfBuf(c->directHandle, kFillLower, n, c->directA.p, n, &lwork) != 0) return 1;
    OCB_CUDA(c, c->directWork.reserve((size_t)lwork + 8, c->stream));
    const int grid = std::max(1, std::min((nRows + 255) / 256, c->numSMs * 8));
    static const double lifts[4] = {0.0, 1.0e-8, 1.0e-5, 1.0e-2};
    for (int t = 0; t < 4; ++t) {
        OCB_CUDA(c, cudaMemsetAsync(c->directA.p, 0, sizeof(double) * (size_t)n * n, c->stream));
        dense_from_bsr_kernel<<<grid, 256, 0, c->stream>>>(nRows, c->rowPtr.p, c->colIdx.p, c->val.p, c->directA.p, (size_t)n, lifts[t]);
        KCHECK(c);
        direct_gather_rhs_kernel<<<grid, 256, 0, c->stream>>>(nRows, c->vertOf.p, d_rhs, negate ? 1 : 0, c->directB.p);
        KCHECK(c);
        OCB_CUDA(c, cudaMemsetAsync(c->directInfo.p, 0, sizeof(int) * 4, c->stream));
        if (S.potrf(c->directHandle, kFillLower, n, c->directA.p, n, c->directWork.p, lwork, c->directInfo.p) != 0) return 1;
        int hInfo[2] = {0, 0};
        OCB_CUDA(c, cudaMemcpyAsync(hInfo, c->directInfo.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        OCB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (hInfo[0] != 0) continue;                          // not positive definite to working precision: lift the diagonal and repeat
        if (S.potrs(c->directHandle, kFillLower, n, 1, c->directA.p, n, c->directB.p, n, c->directInfo.p + 1) != 0) return 1;
        direct_scatter_x_kernel<<<grid, 256, 0, c->stream>>>(nRows, c->vertOf.p, c->directB.p, c->p.p, c->dScal, 0, c->directInfo.p);
        KCHECK(c);
        if (liftsUsed) *liftsUsed = t;
        c->directSolves++;
        return 0;
    }
    return 1;
}

void direct_release(ocb_ctx* c)
{
    c->directA.release(); c->directB.release(); c->directWork.release(); c->directInfo.release();
    // the cuSOLVER handle is left to process teardown (destroying it needs the library, which may already be unloading)
}

}  // namespace ocb
