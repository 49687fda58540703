// optcuts_b200 — linear solve: block-Jacobi preconditioned CG on the BSR(2x2) Hessian, replacing
// Eigen::SimplicialLDLT (EigenLibSolver.cpp:71-107).  sm_100a, fp64.
//
// ONE persistent cooperative kernel runs the whole solve: every CTA owns a contiguous range of block
// rows; the three phases of a CG iteration are separated by a hand-written grid barrier that also
// carries the dot products (each CTA publishes its partial, everybody sums all partials in the same
// order -> bitwise identical scalars in every CTA, deterministic, no host round trip, no atomics on
// doubles).  Convergence is decided on the device.
//
// SpMV mapping: 16 lanes per block row; lane pair (2k, 2k+1) owns block k of the row, the even lane
// the top row of the 2x2 block and the odd lane the bottom row, so one warp-wide 16-byte load reads
// 512 contiguous bytes of the value array (fully coalesced) and x is gathered as double2.
// HBM bytes per CG iteration and block row (7 blocks on average): 7*(32+4) matrix + 4 rowPtr +
// vectors (d, Ap, x, r, z, minv) 144+48 = ~480 B  -> ~228 B per mesh face.
#include "ocb_internal.cuh"
#include <cooperative_groups.h>

namespace ocb {

static constexpr int kPcgBlock = 512;

struct PcgParams {
    int nRows;                 // block rows (= global vertices)
    const int32_t* rowPtr; const int32_t* colIdx; const double* val; const double* minv;
    const double* rhs; int negate;
    double* x; double* r; double* z; double* d; double* Ap;
    double* partials;          // 3 x gridDim (double buffered by phase: 2 sets)
    unsigned* bar;             // [0] arrival counter, [1] generation
    double* scal;              // device scalar block
    double relTol; int maxIt;
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nBlocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned* gen = bar + 1;
        const unsigned g = *gen;
        __threadfence();
        const unsigned t = atomicAdd(bar, 1u);
        if (t == nBlocks - 1) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { __nanosleep(20); }
        }
        __threadfence();
    }
    __syncthreads();
}

// all CTAs sum the published partials in the same order
template <int NV>
__device__ __forceinline__ void gather_partials(const double* partials, int nBlocks, double (&out)[NV])
{
    __shared__ double sm[NV][kPcgBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int b = threadIdx.x; b < nBlocks; b += kPcgBlock)
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] += __ldcg(&partials[(size_t)b * NV + k]);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sm[k][warp] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (int w = 0; w < kPcgBlock / 32; ++w) t += sm[k][w];
        out[k] = t;
    }
    __syncthreads();
}

template <int NV>
__device__ __forceinline__ void publish_partials(double (&v)[NV], double* partials)
{
    __shared__ double sm[NV][kPcgBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sm[k][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = 0.0;
            for (int w = 0; w < kPcgBlock / 32; ++w) t += sm[k][w];
            partials[(size_t)blockIdx.x * NV + k] = t;
        }
    }
    __syncthreads();
}

// y[rows of this CTA] = A * v ; returns this thread's share of v . y.
// v may have been written by other CTAs before the last grid barrier: plain (coherent) loads, no
// __restrict__/__ldg on it.
__device__ __forceinline__ double spmv_rows(const PcgParams& P, int rowBeg, int rowEnd, const double* v, double* y)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 15, kblk = sub >> 1, half = sub & 1;
    const double2* __restrict__ val2 = reinterpret_cast<const double2*>(P.val);
    const double2* v2 = reinterpret_cast<const double2*>(v);
    double dotAcc = 0.0;
    for (int rowBase = rowBeg + warp * 2; rowBase < rowEnd; rowBase += (kPcgBlock / 32) * 2) {
        const int row = rowBase + (lane >> 4);
        const bool active = row < rowEnd;
        const int beg = active ? __ldg(P.rowPtr + row) : 0, end = active ? __ldg(P.rowPtr + row + 1) : 0;
        double acc = 0.0;
        for (int b = beg + kblk; b < end; b += 8) {
            const int col = __ldg(P.colIdx + b);
            const double2 a = __ldg(val2 + 2 * (size_t)b + half);
            const double2 xv = v2[col];
            acc += a.x * xv.x + a.y * xv.y;
        }
        __syncwarp();
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        if (active && sub < 2) {
            y[2 * (size_t)row + half] = acc;
            dotAcc += acc * v[2 * (size_t)row + half];
        }
    }
    return dotAcc;
}

__global__ void __launch_bounds__(kPcgBlock, 1)
pcg_kernel(PcgParams P)
{
    const int nB = gridDim.x;
    const int rowsPer = (P.nRows + nB - 1) / nB;
    const int rowBeg = min(P.nRows, (int)blockIdx.x * rowsPer), rowEnd = min(P.nRows, rowBeg + rowsPer);
    const int sBeg = 2 * rowBeg, sEnd = 2 * rowEnd;
    double* part0 = P.partials;                 // two partial buffers, alternated between barriers
    double* part1 = P.partials + 3 * (size_t)nB;

    // ---- init: x = 0, r = b, z = Minv r, d = z ; rz = r.z, bb = b.b
    double loc[3] = {0.0, 0.0, 0.0};
    for (int row = rowBeg + threadIdx.x; row < rowEnd; row += kPcgBlock) {
        const double b0 = P.negate ? -P.rhs[2 * row] : P.rhs[2 * row];
        const double b1 = P.negate ? -P.rhs[2 * row + 1] : P.rhs[2 * row + 1];
        const double m00 = P.minv[4 * (size_t)row], m01 = P.minv[4 * (size_t)row + 1], m11 = P.minv[4 * (size_t)row + 3];
        const double z0 = m00 * b0 + m01 * b1, z1 = m01 * b0 + m11 * b1;
        P.x[2 * row] = 0.0; P.x[2 * row + 1] = 0.0;
        P.r[2 * row] = b0; P.r[2 * row + 1] = b1;
        P.z[2 * row] = z0; P.z[2 * row + 1] = z1;
        P.d[2 * row] = z0; P.d[2 * row + 1] = z1;
        loc[0] += b0 * z0 + b1 * z1;
        loc[1] += b0 * b0 + b1 * b1;
    }
    publish_partials<3>(loc, part0);
    grid_barrier(P.bar, nB);
    double red[3];
    gather_partials<3>(part0, nB, red);
    double rz = red[0];
    const double bb = red[1];
    const double tol2 = P.relTol * P.relTol * bb;
    double rr = bb;
    int it = 0, status = 0;
    if (bb == 0.0) { status = 0; }
    else {
        for (it = 0; it < P.maxIt; ) {
            // ---- phase A: Ap = A d ; pAp
            loc[0] = spmv_rows(P, rowBeg, rowEnd, P.d, P.Ap); loc[1] = 0.0; loc[2] = 0.0;
            publish_partials<3>(loc, part1);
            grid_barrier(P.bar, nB);
            gather_partials<3>(part1, nB, red);
            const double dAd = red[0];
            if (!(dAd > 0.0)) { status = 2; break; }
            const double alpha = rz / dAd;
            // ---- phase B: x += alpha d ; r -= alpha Ap ; z = Minv r ; rz', rr
            loc[0] = 0.0; loc[1] = 0.0; loc[2] = 0.0;
            for (int row = rowBeg + threadIdx.x; row < rowEnd; row += kPcgBlock) {
                const double2 dd = reinterpret_cast<const double2*>(P.d)[row];
                const double2 ap = reinterpret_cast<const double2*>(P.Ap)[row];
                double2 xx = reinterpret_cast<double2*>(P.x)[row];
                double2 r2 = reinterpret_cast<double2*>(P.r)[row];
                xx.x += alpha * dd.x; xx.y += alpha * dd.y;
                r2.x -= alpha * ap.x; r2.y -= alpha * ap.y;
                const double4 m = reinterpret_cast<const double4*>(P.minv)[row];
                double2 zz; zz.x = m.x * r2.x + m.y * r2.y; zz.y = m.y * r2.x + m.w * r2.y;
                reinterpret_cast<double2*>(P.x)[row] = xx;
                reinterpret_cast<double2*>(P.r)[row] = r2;
                reinterpret_cast<double2*>(P.z)[row] = zz;
                loc[0] += r2.x * zz.x + r2.y * zz.y;
                loc[1] += r2.x * r2.x + r2.y * r2.y;
            }
            publish_partials<3>(loc, part0);
            grid_barrier(P.bar, nB);
            gather_partials<3>(part0, nB, red);
            const double rzNew = red[0];
            rr = red[1];
            ++it;
            if (rr <= tol2) { status = 0; break; }
            if (it >= P.maxIt) { status = 1; break; }
            const double beta = rzNew / rz;
            rz = rzNew;
            // ---- phase C: d = z + beta d
            for (int i = sBeg + threadIdx.x; i < sEnd; i += kPcgBlock) P.d[i] = P.z[i] + beta * P.d[i];
            grid_barrier(P.bar, nB);
        }
        if (it >= P.maxIt && status == 0 && rr > tol2) status = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.scal[S_PCG_ITERS] = (double)it;
        P.scal[S_PCG_RELRES] = bb > 0.0 ? sqrt(rr / bb) : 0.0;
        P.scal[S_PCG_STATUS] = (double)status;
        P.scal[S_PCG_BNORM] = sqrt(bb);
    }
}

// stand-alone y = A x (ocb_multiply; also used by tests)
__global__ void __launch_bounds__(kPcgBlock)
spmv_kernel(PcgParams P, const double* __restrict__ v, double* __restrict__ y)
{
    const int nB = gridDim.x;
    const int rowsPer = (P.nRows + nB - 1) / nB;
    const int rowBeg = min(P.nRows, (int)blockIdx.x * rowsPer), rowEnd = min(P.nRows, rowBeg + rowsPer);
    spmv_rows(P, rowBeg, rowEnd, v, y);
}

// block-Jacobi preconditioner: inverse of every 2x2 diagonal block
__global__ void __launch_bounds__(256)
jacobi_setup_kernel(int nRows, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx,
                    const double* __restrict__ val, double* __restrict__ minv, int* __restrict__ bad)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        double a00 = 0.0, a01 = 0.0, a11 = 0.0;
        for (int b = rowPtr[row]; b < rowPtr[row + 1]; ++b)
            if (colIdx[b] == row) { a00 = val[4 * (size_t)b]; a01 = 0.5 * (val[4 * (size_t)b + 1] + val[4 * (size_t)b + 2]); a11 = val[4 * (size_t)b + 3]; }
        const double det = a00 * a11 - a01 * a01;
        if (!(a00 > 0.0) || !(det > 0.0)) { atomicAdd(bad, 1); minv[4 * (size_t)row] = 1.0; minv[4 * (size_t)row + 1] = 0.0; minv[4 * (size_t)row + 2] = 0.0; minv[4 * (size_t)row + 3] = 1.0; continue; }
        minv[4 * (size_t)row] = a11 / det; minv[4 * (size_t)row + 1] = -a01 / det;
        minv[4 * (size_t)row + 2] = -a01 / det; minv[4 * (size_t)row + 3] = a00 / det;
    }
}

#define KCHECK(c) do { (c)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return cuda_fail((c), _e, __func__); } while (0)

static PcgParams make_params(ocb_ctx* c)
{
    PcgParams P;
    P.nRows = c->nVtot; P.rowPtr = c->rowPtr.p; P.colIdx = c->colIdx.p; P.val = c->val.p; P.minv = c->minv.p;
    P.rhs = nullptr; P.negate = 0; P.x = c->p.p; P.r = c->pr.p; P.z = c->pz.p; P.d = c->pd.p; P.Ap = c->pAp.p;
    P.partials = c->partials.p; P.bar = c->sync.p + 16; P.scal = c->dScal; P.relTol = 1e-12; P.maxIt = 1;
    return P;
}

static int pcg_grid(ocb_ctx* c)
{
    if (c->pcgGrid == 0) {
        int perSM = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, pcg_kernel, kPcgBlock, 0);
        if (perSM < 1) perSM = 1;
        if (perSM > 2) perSM = 2;
        c->pcgGrid = c->numSMs * perSM;
    }
    // small systems: do not spread 2 rows per warp thinner than one warp-iteration per CTA
    const int rowsPerCtaPass = (kPcgBlock / 32) * 2;
    int need = (c->nVtot + rowsPerCtaPass - 1) / rowsPerCtaPass;
    int g = c->pcgGrid < need ? c->pcgGrid : need;
    return g < 1 ? 1 : g;
}

int launch_spmv(ocb_ctx* c, const double* dx, double* dy)
{
    ProfScope prof(c, K_SPMV);
    PcgParams P = make_params(c);
    spmv_kernel<<<pcg_grid(c), kPcgBlock, 0, c->stream>>>(P, dx, dy);
    KCHECK(c);
    return 0;
}

int launch_jacobi_setup(ocb_ctx* c)
{
    int* bad = reinterpret_cast<int*>(c->sync.p + 8);
    OCB_CUDA(c, cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    OCB_CUDA(c, c->minv.reserve(4 * (size_t)c->nVtot, c->stream));
    int grid = (c->nVtot + 255) / 256; if (grid > c->numSMs * 8) grid = c->numSMs * 8; if (grid < 1) grid = 1;
    {
        ProfScope prof(c, K_JACOBI_SETUP);
        jacobi_setup_kernel<<<grid, 256, 0, c->stream>>>(c->nVtot, c->rowPtr.p, c->colIdx.p, c->val.p, c->minv.p, bad);
        KCHECK(c);
    }
    int hBad = 0;
    OCB_CUDA(c, cudaMemcpyAsync(&hBad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (hBad) return set_err(c, OCB_ERR_BREAKDOWN, "a diagonal 2x2 block of the matrix is not positive definite");
    return 0;
}

int launch_pcg(ocb_ctx* c, const double* d_rhs, bool negate_rhs, double rel_tol, int max_it)
{
    ProfScope prof(c, K_PCG);
    const size_t n = c->nSys();
    OCB_CUDA(c, c->pr.reserve(n, c->stream)); OCB_CUDA(c, c->pz.reserve(n, c->stream));
    OCB_CUDA(c, c->pd.reserve(n, c->stream)); OCB_CUDA(c, c->pAp.reserve(n, c->stream));
    OCB_CUDA(c, c->p.reserve(n, c->stream));
    const int grid = pcg_grid(c);
    OCB_CUDA(c, c->partials.reserve((size_t)grid * 6 + 64, c->stream));
    PcgParams P = make_params(c);
    P.rhs = d_rhs; P.negate = negate_rhs ? 1 : 0; P.relTol = rel_tol; P.maxIt = max_it;
    OCB_CUDA(c, cudaMemsetAsync(c->sync.p + 16, 0, 2 * sizeof(unsigned), c->stream));
    void* args[] = {&P};
    OCB_CUDA(c, cudaLaunchCooperativeKernel((void*)pcg_kernel, dim3(grid), dim3(kPcgBlock), args, 0, c->stream));
    c->launches++;
    return 0;
}

}  // namespace ocb
