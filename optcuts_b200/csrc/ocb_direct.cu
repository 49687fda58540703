// optcuts_b200 — safety net of the linear solve: dense Cholesky of the assembled matrix on the device.
//
// The reference solves every Newton system with a sparse LDL^T (EigenLibSolver.cpp:71-107), which returns a usable direction
// even when the matrix spans 30 orders of magnitude (the Tutte starts of 9 of the 71 benchmark meshes: ||g||^2 up to 1e51).
// Conjugate gradients cannot: on those systems the two-level preconditioner comes out indefinite, block-Jacobi CG stalls at
// its iteration cap, and the truncated iterate is a poor direction (cat_noUV: E 2.6e9 -> 3.5e8 after 30 iterations where the
// reference reaches 5.1e5, profiles/r2_cat_noUV_before_direct.txt).  For systems of at most kDirectMaxDof unknowns such a solve
// is repeated here directly, as a BLOCK-TRIDIAGONAL Cholesky: the rows are ordered by breadth-first levels of the matrix graph
// (a vertex of level l is coupled to levels l-1, l, l+1 only), consecutive levels are merged into blocks of ~200 vertices, and
// the factorisation walks the block chain
//     D_k <- D_k - L_{k,k-1} L_{k,k-1}^T ;  L_kk = chol(D_k) ;  L_{k+1,k} = B_k L_kk^-T
// with dense library calls per block (cuSOLVER potrf, cuBLAS syrk / trsm, then trsv / gemv for the two substitutions): plain
// dense factorisations of 400 x 400 blocks, not kernels of this path.  ~n b^2 flops instead of n^3 / 3 (15k unknowns: 1e10 against
// 1e12).  A mesh whose levels are too wide for that (a block over 8192 unknowns) takes one dense potrf of the whole matrix.
// Both libraries are loaded with dlopen at first use, so this library has no link-time dependency on them; when they are missing,
// or the system is larger, the block-Jacobi retry stays the fallback.  A matrix that is indefinite by rounding makes potrf fail:
// the diagonal is then lifted relatively (1e-8, 1e-5, 1e-2) as after a CG breakdown.
// The libraries' host side uses OpenMP: under OMP_NUM_THREADS=1 (what torchrun exports to its ranks) a solve costs ~7x
// (male_2: 70-93 s for 150 Newton iterations instead of 12 s, profiles/r2_bench_n2_before_omp.json).
#include "ocb_internal.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

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
typedef int (*fnDestroy)(void*);
Cusolver load_cusolver()
{
    Cusolver S;
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
Cusolver& cusolver() { static Cusolver S = load_cusolver(); return S; }      // (thread-safe: initialised once, on first use)
constexpr int kFillLower = 0;                 // CUBLAS_FILL_MODE_LOWER
constexpr int kOpN = 0, kOpT = 1, kSideRight = 1, kDiagNonUnit = 0;

typedef void* cublasHandle_t;
typedef int (*fnBCreate)(cublasHandle_t*);
typedef int (*fnBSetStream)(cublasHandle_t, cudaStream_t);
typedef int (*fnSyrk)(cublasHandle_t, int, int, int, int, const double*, const double*, int, const double*, double*, int);
typedef int (*fnTrsm)(cublasHandle_t, int, int, int, int, int, int, const double*, const double*, int, double*, int);
typedef int (*fnGemv)(cublasHandle_t, int, int, int, const double*, const double*, int, const double*, int, const double*, double*, int);
typedef int (*fnTrsv)(cublasHandle_t, int, int, int, int, const double*, int, double*, int);
struct Cublas {
    bool tried = false, ok = false;
    void* lib = nullptr;
    fnBCreate create = nullptr; fnBSetStream setStream = nullptr; fnSyrk syrk = nullptr; fnTrsm trsm = nullptr; fnGemv gemv = nullptr; fnTrsv trsv = nullptr;
};
Cublas load_cublas()
{
    Cublas S;
    S.tried = true;
    const char* names[] = {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so.13", "libcublas.so"};
    for (const char* nme : names) { S.lib = dlopen(nme, RTLD_NOW | RTLD_LOCAL); if (S.lib) break; }
    if (!S.lib) return S;
    S.create = (fnBCreate)dlsym(S.lib, "cublasCreate_v2"); S.setStream = (fnBSetStream)dlsym(S.lib, "cublasSetStream_v2");
    S.syrk = (fnSyrk)dlsym(S.lib, "cublasDsyrk_v2"); S.trsm = (fnTrsm)dlsym(S.lib, "cublasDtrsm_v2");
    S.gemv = (fnGemv)dlsym(S.lib, "cublasDgemv_v2"); S.trsv = (fnTrsv)dlsym(S.lib, "cublasDtrsv_v2");
    S.ok = S.create && S.setStream && S.syrk && S.trsm && S.gemv && S.trsv;
    return S;
}
Cublas& cublas() { static Cublas S = load_cublas(); return S; }

// block-tridiagonal layout: row (solver order) -> position in the level order; per block its first position and the offsets of
// D_k (n_k x n_k, column-major) and B_k (n_{k+1} x n_k: rows of block k+1, columns of block k) in one buffer
__global__ void __launch_bounds__(256)
tridiag_from_bsr_kernel(int nRows, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, const double* __restrict__ val,
                        const int32_t* __restrict__ pos, const int32_t* __restrict__ blkOf, const int32_t* __restrict__ blkBeg,
                        const long long* __restrict__ offD, const long long* __restrict__ offB, double* __restrict__ buf, double lift)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        const int pr = pos[row], kr = blkOf[row];
        for (int b = rowPtr[row]; b < rowPtr[row + 1]; ++b) {
            const int col = colIdx[b], pc = pos[col], kc = blkOf[col];
            const double* v = val + 4 * (size_t)b;             // [a00 a01 a10 a11] of block (row, col)
            double* dst; long long ld; int r0, c0;
            if (kr == kc) { ld = 2LL * (blkBeg[kr + 1] - blkBeg[kr]); dst = buf + offD[kr]; r0 = 2 * (pr - blkBeg[kr]); c0 = 2 * (pc - blkBeg[kr]); }
            else if (kr == kc + 1) { ld = 2LL * (blkBeg[kr + 1] - blkBeg[kr]); dst = buf + offB[kc]; r0 = 2 * (pr - blkBeg[kr]); c0 = 2 * (pc - blkBeg[kc]); }
            else continue;                                     // the upper triangle's copy (or, never: a coupling across two blocks)
            const double f = col == row ? 1.0 + lift : 1.0;
            dst[(long long)c0 * ld + r0] = v[0] * f;
            dst[(long long)(c0 + 1) * ld + r0] = v[1];
            dst[(long long)c0 * ld + r0 + 1] = v[2];
            dst[(long long)(c0 + 1) * ld + r0 + 1] = v[3] * f;
        }
    }
}
__global__ void __launch_bounds__(256)
tridiag_gather_rhs_kernel(int nRows, const int32_t* __restrict__ vertOf, const int32_t* __restrict__ pos, const double* __restrict__ rhs, int negate, double* __restrict__ b)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        const size_t src = vertOf ? (size_t)vertOf[row] : (size_t)row;
        const double b0 = rhs[2 * src], b1 = rhs[2 * src + 1];
        b[2 * (size_t)pos[row]] = negate ? -b0 : b0; b[2 * (size_t)pos[row] + 1] = negate ? -b1 : b1;
    }
}
__global__ void __launch_bounds__(256)
tridiag_scatter_x_kernel(int nRows, const int32_t* __restrict__ vertOf, const int32_t* __restrict__ pos, const double* __restrict__ x, double* __restrict__ xOut)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        const size_t dst = vertOf ? (size_t)vertOf[row] : (size_t)row;
        xOut[2 * dst] = x[2 * (size_t)pos[row]]; xOut[2 * dst + 1] = x[2 * (size_t)pos[row] + 1];
    }
}
__global__ void tridiag_info_sum_kernel(int nb, int* __restrict__ info) { int s = 0; for (int k = 0; k < nb; ++k) s |= info[1 + k] != 0 ? (k + 1) : 0; info[0] = s; }

// breadth-first level order of the solver-order pattern (all components), levels merged into blocks of >= target vertices
void level_blocks(const std::vector<int32_t>& rp, const std::vector<int32_t>& ci, int n, int target, std::vector<int32_t>& pos, std::vector<int32_t>& blkOf,
                  std::vector<int32_t>& blkBeg)
{
    std::vector<int32_t> order; order.reserve((size_t)n);
    std::vector<int32_t> levelBeg;                            // positions where a level starts
    std::vector<int32_t> dist((size_t)n, -1);
    auto bfs = [&](int s0, std::vector<int32_t>& out, std::vector<int32_t>* lv) {
        const size_t first = out.size();
        dist[s0] = 0; out.push_back(s0);
        if (lv) lv->push_back((int32_t)first);
        size_t head = first, levelEnd = out.size();
        while (head < out.size()) {
            if (head == levelEnd) { if (lv) lv->push_back((int32_t)head); levelEnd = out.size(); }
            const int v = out[head++];
            for (int q = rp[v]; q < rp[v + 1]; ++q) { const int u = ci[q]; if (dist[u] < 0) { dist[u] = dist[v] + 1; out.push_back(u); } }
        }
    };
    std::vector<int32_t> tmp;
    for (int s0 = 0; s0 < n; ++s0) {
        if (dist[s0] >= 0) continue;
        // pseudo-peripheral start: the last vertex of a first sweep (narrower levels than from an arbitrary start)
        tmp.clear();
        bfs(s0, tmp, nullptr);
        const int far = tmp.back();
        for (int v : tmp) dist[v] = -1;
        bfs(far, order, &levelBeg);
    }
    levelBeg.push_back(n);
    pos.assign((size_t)n, 0);
    for (int i = 0; i < n; ++i) pos[order[i]] = i;
    blkBeg.assign(1, 0);
    for (size_t l = 0; l + 1 < levelBeg.size(); ++l)
        if (levelBeg[l + 1] - blkBeg.back() >= target || l + 2 == levelBeg.size()) blkBeg.push_back(levelBeg[l + 1]);
    if (blkBeg.back() != n) blkBeg.push_back(n);
    blkOf.assign((size_t)n, 0);
    for (size_t k = 0; k + 1 < blkBeg.size(); ++k) for (int i = blkBeg[k]; i < blkBeg[k + 1]; ++i) blkOf[order[i]] = (int32_t)k;
}

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
direct_scatter_x_kernel(int nRows, const int32_t* __restrict__ vertOf, const double* __restrict__ x, double* __restrict__ xOut, const int* __restrict__ devInfo)
{
    const bool ok = devInfo[0] == 0 && devInfo[1] == 0;       // potrf and potrs both went through
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        const size_t dst = vertOf ? (size_t)vertOf[row] : (size_t)row;
        if (ok) { xOut[2 * dst] = x[2 * (size_t)row]; xOut[2 * dst + 1] = x[2 * (size_t)row + 1]; }
    }
}

}  // namespace

// host only (tests): the level order and blocks the block-tridiagonal path would use for a vertex pattern given as CSR (rows may be
// unsorted, the diagonal may be present).  pos / blkOf: n entries each; blkBeg: up to n + 1 entries.  Returns the number of blocks.
int direct_level_blocks_host(int n, const int32_t* rowPtr, const int32_t* colIdx, int target, int32_t* pos, int32_t* blkOf, int32_t* blkBeg)
{
    std::vector<int32_t> rp(rowPtr, rowPtr + n + 1), ci(colIdx, colIdx + rowPtr[n]), p, b, bb;
    level_blocks(rp, ci, n, target, p, b, bb);
    std::copy(p.begin(), p.end(), pos); std::copy(b.begin(), b.end(), blkOf); std::copy(bb.begin(), bb.end(), blkBeg);
    return (int)bb.size() - 1;
}

bool direct_solver_available(const ocb_ctx* c)
{
    return !c->systemScaled && 2 * (size_t)c->nVtot <= (size_t)kDirectMaxDof && cusolver().ok;
}

// ---- block-tridiagonal path.  Returns 0 solved, 1 not applicable / not factorisable (the caller tries the dense path), < 0 error
static int tridiag_solve(ocb_ctx* c, const double* d_rhs, bool negate, int* liftsUsed)
{
    Cusolver& S = cusolver();
    Cublas& B = cublas();
    const int nRows = c->nVtot;
    if (!S.ok || !B.ok || (int)c->hSRowPtr.size() != nRows + 1) return 1;
    static const int target = []() { const char* e = getenv("OCB_DIRECT_BLOCK"); const int v = e ? atoi(e) : 192; return v < 16 ? 16 : v; }();
    std::vector<int32_t> pos, blkOf, blkBeg;
    level_blocks(c->hSRowPtr, c->hSColIdx, nRows, target, pos, blkOf, blkBeg);
    const int nb = (int)blkBeg.size() - 1;
    std::vector<long long> offD((size_t)nb), offB((size_t)nb);
    long long tot = 0; int maxN = 0;
    for (int k = 0; k < nb; ++k) {
        const long long nk = 2LL * (blkBeg[k + 1] - blkBeg[k]);
        maxN = std::max(maxN, (int)nk);
        offD[k] = tot; tot += nk * nk;
        offB[k] = tot; if (k + 1 < nb) tot += 2LL * (blkBeg[k + 2] - blkBeg[k + 1]) * nk;
    }
    if (maxN > 8192) return 1;                                 // levels too wide: the dense path is as good
    if (!c->directHandle && S.create(&c->directHandle) != 0) { c->directHandle = nullptr; return 1; }
    if (!c->directBlas && B.create(&c->directBlas) != 0) { c->directBlas = nullptr; return 1; }
    if (S.setStream(c->directHandle, c->stream) != 0 || B.setStream(c->directBlas, c->stream) != 0) return 1;
    const int n = 2 * nRows;
    OCB_CUDA(c, c->directA.reserve((size_t)tot + 8, c->stream));
    OCB_CUDA(c, c->directB.reserve((size_t)n + 8, c->stream));
    OCB_CUDA(c, c->directInfo.reserve((size_t)nb + 8, c->stream));
    OCB_CUDA(c, c->directI.reserve(2 * (size_t)nRows + (size_t)nb + 8, c->stream));
    OCB_CUDA(c, c->directL.reserve(2 * (size_t)nb + 8, c->stream));
    int32_t* dPos = c->directI.p; int32_t* dBlkOf = dPos + nRows; int32_t* dBlkBeg = dBlkOf + nRows;
    long long* dOffD = c->directL.p; long long* dOffB = dOffD + nb;
    OCB_CUDA(c, cudaMemcpyAsync(dPos, pos.data(), sizeof(int32_t) * nRows, cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(dBlkOf, blkOf.data(), sizeof(int32_t) * nRows, cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(dBlkBeg, blkBeg.data(), sizeof(int32_t) * (nb + 1), cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(dOffD, offD.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(dOffB, offB.data(), sizeof(long long) * nb, cudaMemcpyHostToDevice, c->stream));
    int lwork = 0;
    if (S.potrfBuf(c->directHandle, kFillLower, maxN, c->directA.p, maxN, &lwork) != 0) return 1;
    OCB_CUDA(c, c->directWork.reserve((size_t)lwork + 8, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));            // the host vectors above go out of scope at return; pageable copies are staged, but be explicit
    const int grid = std::max(1, std::min((nRows + 255) / 256, c->numSMs * 8));
    const double one = 1.0, mone = -1.0;
    static const double lifts[4] = {0.0, 1.0e-8, 1.0e-5, 1.0e-2};
    for (int t = 0; t < 4; ++t) {
        OCB_CUDA(c, cudaMemsetAsync(c->directA.p, 0, sizeof(double) * (size_t)tot, c->stream));
        OCB_CUDA(c, cudaMemsetAsync(c->directInfo.p, 0, sizeof(int) * ((size_t)nb + 1), c->stream));
        tridiag_from_bsr_kernel<<<grid, 256, 0, c->stream>>>(nRows, c->rowPtr.p, c->colIdx.p, c->val.p, dPos, dBlkOf, dBlkBeg, dOffD, dOffB, c->directA.p, lifts[t]);
        KCHECK(c);
        tridiag_gather_rhs_kernel<<<grid, 256, 0, c->stream>>>(nRows, c->vertOf.p, dPos, d_rhs, negate ? 1 : 0, c->directB.p);
        KCHECK(c);
        double* A = c->directA.p; double* x = c->directB.p;
        // factorisation fused with the forward substitution, block by block
        for (int k = 0; k < nb; ++k) {
            const int nk = 2 * (blkBeg[k + 1] - blkBeg[k]);
            double* Dk = A + offD[k]; double* xk = x + 2 * (size_t)blkBeg[k];
            if (k > 0) {
                const int np = 2 * (blkBeg[k] - blkBeg[k - 1]);
                const double* Lp = A + offB[k - 1];                 // L_{k,k-1}: nk x np
                if (B.syrk(c->directBlas, kFillLower, kOpN, nk, np, &mone, Lp, nk, &one, Dk, nk) != 0) return 1;
                if (B.gemv(c->directBlas, kOpN, nk, np, &mone, Lp, nk, x + 2 * (size_t)blkBeg[k - 1], 1, &one, xk, 1) != 0) return 1;
            }
            if (S.potrf(c->directHandle, kFillLower, nk, Dk, nk, c->directWork.p, lwork, c->directInfo.p + 1 + k) != 0) return 1;
            if (B.trsv(c->directBlas, kFillLower, kOpN, kDiagNonUnit, nk, Dk, nk, xk, 1) != 0) return 1;
            if (k + 1 < nb) {
                const int nn = 2 * (blkBeg[k + 2] - blkBeg[k + 1]);
                if (B.trsm(c->directBlas, kSideRight, kFillLower, kOpT, kDiagNonUnit, nn, nk, &one, Dk, nk, A + offB[k], nn) != 0) return 1;
            }
        }
        tridiag_info_sum_kernel<<<1, 1, 0, c->stream>>>(nb, c->directInfo.p);
        KCHECK(c);
        int hInfo = 0;
        OCB_CUDA(c, cudaMemcpyAsync(&hInfo, c->directInfo.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        OCB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (hInfo != 0) continue;                                  // a block was not positive definite to working precision: lift and repeat
        for (int k = nb - 1; k >= 0; --k) {                        // backward substitution
            const int nk = 2 * (blkBeg[k + 1] - blkBeg[k]);
            double* xk = x + 2 * (size_t)blkBeg[k];
            if (k + 1 < nb) {
                const int nn = 2 * (blkBeg[k + 2] - blkBeg[k + 1]);
                if (B.gemv(c->directBlas, kOpT, nn, nk, &mone, A + offB[k], nn, x + 2 * (size_t)blkBeg[k + 1], 1, &one, xk, 1) != 0) return 1;
            }
            if (B.trsv(c->directBlas, kFillLower, kOpT, kDiagNonUnit, nk, A + offD[k], nk, xk, 1) != 0) return 1;
        }
        tridiag_scatter_x_kernel<<<grid, 256, 0, c->stream>>>(nRows, c->vertOf.p, dPos, x, c->p.p);
        KCHECK(c);
        if (liftsUsed) *liftsUsed = t;
        c->directSolves++;
        static const bool dbg = []() { const char* e = getenv("OCB_PCG_DEBUG"); return e && atoi(e); }();
        if (dbg) fprintf(stderr, "[ocb direct] block-tridiagonal Cholesky: %d unknowns, %d blocks (largest %d), %.1f MB, diagonal lift %g\n", n, nb, maxN, 8e-6 * (double)tot, lifts[t]);
        return 0;
    }
    return 1;
}

// x (internal order, c->p) = A^-1 (+-rhs).  Returns 0 on success, 1 when the factorisation failed for every lift (the caller
// keeps its iterative fallback), < 0 on a CUDA / library error.
int launch_direct_solve(ocb_ctx* c, const double* d_rhs, bool negate, int* liftsUsed)
{
    Cusolver& S = cusolver();
    if (!S.ok) return 1;
    ProfScope prof(c, K_PCG);
    {
        static const bool denseOnly = []() { const char* e = getenv("OCB_DIRECT_DENSE"); return e && atoi(e); }();
        const int rt = denseOnly ? 1 : tridiag_solve(c, d_rhs, negate, liftsUsed);
        if (rt <= 0) return rt;
    }
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
        direct_scatter_x_kernel<<<grid, 256, 0, c->stream>>>(nRows, c->vertOf.p, c->directB.p, c->p.p, c->directInfo.p);
        KCHECK(c);
        if (liftsUsed) *liftsUsed = t;
        c->directSolves++;
        return 0;
    }
    return 1;
}

void direct_release(ocb_ctx* c)
{
    c->directA.release(); c->directB.release(); c->directWork.release(); c->directInfo.release(); c->directI.release(); c->directL.release();
    // the library handles (the libraries are never dlclose()d, so their entry points are still there)
    if (c->directHandle && cusolver().lib) { fnDestroy d = (fnDestroy)dlsym(cusolver().lib, "cusolverDnDestroy"); if (d) d(c->directHandle); }
    if (c->directBlas && cublas().lib) { fnDestroy d = (fnDestroy)dlsym(cublas().lib, "cublasDestroy_v2"); if (d) d(c->directBlas); }
    c->directHandle = nullptr; c->directBlas = nullptr;
}

}  // namespace ocb
