// Micro-benchmark: cycles per pivot of the register-resident 48x48 Gauss-Jordan tile inversion (tile_invert_smem of
// ocb_mas.cu, copied verbatim below) on one CTA.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tile_invert_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
static constexpr double kMasPivotTol = 1e-6;
static constexpr int kCB = 48, kCBs = 48 + 2;
static constexpr int kDenseThreads = 576;
__device__ __forceinline__ void tile_invert_smem(double (*T)[kCBs], double* buf, const double* d0s)
{
    __syncthreads();
    const int t = threadIdx.x;
    if (t < 160) {
        const bool act = t < 144;
        const int r = act ? t / 3 : 0, c0 = act ? 16 * (t % 3) : 0;
        double d[16];
#pragma unroll
        for (int e = 0; e < 16; e += 2) { const double2 v = *reinterpret_cast<const double2*>(&T[r][c0 + e]); d[e] = v.x; d[e + 1] = v.y; }
        // The pivot loop is unrolled over the 16 columns of a segment so that every register index is a compile-time
        // constant (a run-time index would push d[] into local memory; rotating the registers instead was measured 2x
        // slower).  What bounds a pivot is the chain of DEPENDENT fp64 operations between two barriers (~700 cycles with
        // the IEEE reciprocal), so the reciprocal is the hardware approximation plus ONE Newton step (relative error
        // ~1e-12: this is a preconditioner, stored in fp32 anyway).
        for (int kq = 0; kq < kCB / 16; ++kq) {
            const bool mine = act && c0 == 16 * kq;           // this thread's segment holds the pivot columns of this round
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                const int k = 16 * kq + kk;
                double* rowb = buf + (k & 1) * 2 * kCB;
                double* colb = rowb + kCB;
                if (act && r == k) {
#pragma unroll
                    for (int e = 0; e < 16; e += 2) *reinterpret_cast<double2*>(rowb + c0 + e) = make_double2(d[e], d[e + 1]);
                }
                if (mine) colb[r] = d[kk];
                asm volatile("bar.sync 1, 160;" ::: "memory");
                // Everything below is straight-line code ordered for the in-order issue: the reciprocal goes first, the
                // collapsed-pivot test runs in its shadow and is applied with selects, and the two candidates for the row
                // coefficient are formed in parallel.  The reciprocal is the hardware approximation plus ONE Newton step
                // (relative error ~1e-12).  The bare approximation (1e-6) is NOT enough: the coarse Galerkin matrices are
                // ill-conditioned, the inverse came out indefinite on the third bimba iteration and CG stalled
                // (tools/gpu_freerun_check.py).
                const double p = rowb[k], dk0 = d0s[k], f = colb[r];
                double ip;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ip) : "d"(p));
                ip = fma(ip, fma(-p, ip, 1.0), ip);
                const bool bad = !(dk0 > 0.0) || !(p > kMasPivotTol * dk0);
                // one FMA per element for every row: the pivot row's own values ARE the row buffer, so its scaling d * g is
                // d + (g - 1) * rowb; a collapsed pivot zeroes its row (coefficient -1) and leaves the other rows alone
                const double gOther = -f * ip, gPivot = ip - 1.0;
                const double coef = r != k ? (bad ? 0.0 : gOther) : (bad ? -1.0 : gPivot);
                const double g = bad ? 0.0 : (r != k ? gOther : ip);       // new value of the pivot-column element
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const double2 rj = *reinterpret_cast<const double2*>(rowb + c0 + e);
                    d[e] += coef * rj.x; d[e + 1] += coef * rj.y;
                }
                if (mine) d[kk] = g;
            }
        }
        if (act) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) *reinterpret_cast<double2*>(&T[r][c0 + e]) = make_double2(d[e], d[e + 1]);
        }
    }
    __syncthreads();
}


// ---- 2x2 block pivots: pivots (k, k+1) are eliminated together, so the chain of dependent steps is 24 long instead of
// 48.  buf: 2 x (2 x 48 row values + 2 x 48 column values).
template <int MAP>
__device__ __forceinline__ void tile_invert_smem2(double (*T)[kCBs], double* buf, const double* d0s)
{
    __syncthreads();
    const int t = threadIdx.x;
    if (t < 160) {
        const bool act = t < 144;
        const int r = act ? (MAP ? t % kCB : t / 3) : 0, c0 = act ? 16 * (MAP ? t / kCB : t % 3) : 0;
        double d[16];
#pragma unroll
        for (int e = 0; e < 16; e += 2) { const double2 v = *reinterpret_cast<const double2*>(&T[r][c0 + e]); d[e] = v.x; d[e + 1] = v.y; }
        for (int kq = 0; kq < kCB / 16; ++kq) {
            const bool mine = act && c0 == 16 * kq;
#pragma unroll
            for (int kk = 0; kk < 16; kk += 2) {
                const int k = 16 * kq + kk;
                double* row0 = buf + ((k >> 1) & 1) * 4 * kCB;
                double* row1 = row0 + kCB;
                double* col0 = row1 + kCB;
                double* col1 = col0 + kCB;
                if (act && (r == k || r == k + 1)) {
                    double* dst = (r == k ? row0 : row1) + c0;
#pragma unroll
                    for (int e = 0; e < 16; e += 2) *reinterpret_cast<double2*>(dst + e) = make_double2(d[e], d[e + 1]);
                }
                if (mine) { col0[r] = d[kk]; col1[r] = d[kk + 1]; }
                asm volatile("bar.sync 1, 160;" ::: "memory");
                const double a = row0[k], b = 0.5 * (row0[k + 1] + row1[k]), cc = row1[k + 1];
                const double da = d0s[k], dc = d0s[k + 1], f0 = col0[r], f1 = col1[r];
                const double det = fma(a, cc, -b * b);
                double idet;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(idet) : "d"(det));
                idet = fma(idet, fma(-det, idet, 1.0), idet);
                const bool badA = !(da > 0.0) || !(a > kMasPivotTol * da);
                // second pivot after the first: cc - b^2 / a = det / a
                const bool badC = badA ? (!(dc > 0.0) || !(cc > kMasPivotTol * dc)) : (!(dc > 0.0) || !(det > kMasPivotTol * dc * a));
                double p00, p01, p11;
                if (!badA && !badC) { p00 = cc * idet; p01 = -b * idet; p11 = a * idet; }
                else {                                        // rare, warp-uniform: a dropped DOF leaves a scalar pivot (or none)
                    p00 = p01 = p11 = 0.0;
                    if (!badA) { double ia; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ia) : "d"(a)); p00 = fma(ia, fma(-a, ia, 1.0), ia); }
                    else if (!badC) { double ic; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ic) : "d"(cc)); p11 = fma(ic, fma(-cc, ic, 1.0), ic); }
                }
                // row update d += g0 * row_k + g1 * row_k1: other rows (g0, g1) = -(f0, f1) Pinv; the pivot rows become
                // Pinv [row_k; row_k1], i.e. their own buffer line enters with (Pinv - I)
                double g0 = -(f0 * p00 + f1 * p01), g1 = -(f0 * p01 + f1 * p11);
                double n0 = g0, n1 = g1;                      // new values of the two pivot-column elements
                if (r == k) { g0 = p00 - 1.0; g1 = p01; n0 = p00; n1 = p01; }
                if (r == k + 1) { g0 = p01; g1 = p11 - 1.0; n0 = p01; n1 = p11; }
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const double2 ra = *reinterpret_cast<const double2*>(row0 + c0 + e);
                    const double2 rb = *reinterpret_cast<const double2*>(row1 + c0 + e);
                    d[e] = fma(g1, rb.x, fma(g0, ra.x, d[e])); d[e + 1] = fma(g1, rb.y, fma(g0, ra.y, d[e + 1]));
                }
                if (mine) { d[kk] = n0; d[kk + 1] = n1; }
            }
        }
        if (act) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) *reinterpret_cast<double2*>(&T[r][c0 + e]) = make_double2(d[e], d[e + 1]);
        }
    }
    __syncthreads();
}

// ---- scalar pivots, W columns per thread: 48 * 48 / W threads share the rank-1 update (W = 16: 5 warps, 8: 9 warps, 4: 18 warps)
template <int W, int XMODE = 0, int MAP = 0>
__device__ __forceinline__ void tile_invert_smemW(double (*T)[kCBs], double* buf, const double* d0s)
{
    // MAP 0: row-major (thread t: row t / SEG, segment t % SEG); 1: segment-major, 64 threads per segment (a warp reads ONE
    // pivot-row address: pure broadcast, no bank conflict); 2: segment-major, 48 threads per segment
    constexpr int SEG = kCB / W, NACT = MAP == 1 ? 64 * SEG : kCB * SEG, NBAR = (NACT + 31) / 32 * 32;
    __syncthreads();
    const int t = threadIdx.x;
    if (t < NBAR) {
        const bool act = MAP == 1 ? (t < NACT && (t & 63) < kCB) : t < NACT;
        const int r = !act ? 0 : (MAP == 0 ? t / SEG : (MAP == 1 ? (t & 63) : t % kCB));
        const int c0 = !act ? 0 : W * (MAP == 0 ? t % SEG : (MAP == 1 ? t >> 6 : t / kCB));
        double d[W];
#pragma unroll
        for (int e = 0; e < W; e += 2) { const double2 v = *reinterpret_cast<const double2*>(&T[r][c0 + e]); d[e] = v.x; d[e + 1] = v.y; }
        for (int kq = 0; kq < SEG; ++kq) {
            const bool mine = act && c0 == W * kq;
#pragma unroll
            for (int kk = 0; kk < W; ++kk) {
                const int k = W * kq + kk;
                double* rowb = buf + (k & 1) * 2 * kCB;
                double* colb = rowb + kCB;
                if (XMODE != 2) {
                    if (act && r == k) {
#pragma unroll
                        for (int e = 0; e < W; e += 2) *reinterpret_cast<double2*>(rowb + c0 + e) = make_double2(d[e], d[e + 1]);
                    }
                    if (mine) colb[r] = d[kk];
                }
                if (XMODE != 1) asm volatile("bar.sync 1, %0;" ::"n"(NBAR) : "memory");
                const double p = XMODE == 2 ? d[kk] + 2.0 : rowb[k], dk0 = d0s[k], f = XMODE == 2 ? d[(kk + 1) % W] : colb[r];
                double ip;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ip) : "d"(p));
                ip = fma(ip, fma(-p, ip, 1.0), ip);
                const bool bad = !(dk0 > 0.0) || !(p > kMasPivotTol * dk0);
                const double gOther = -f * ip, gPivot = ip - 1.0;
                const double coef = r != k ? (bad ? 0.0 : gOther) : (bad ? -1.0 : gPivot);
                const double g = bad ? 0.0 : (r != k ? gOther : ip);
#pragma unroll
                for (int e = 0; e < W; e += 2) {
                    const double2 rj = XMODE == 2 ? make_double2(d[(e + 3) % W] * 1e-3, d[(e + 5) % W] * 1e-3) : *reinterpret_cast<const double2*>(rowb + c0 + e);
                    d[e] += coef * rj.x; d[e + 1] += coef * rj.y;
                }
                if (mine) d[kk] = g;
            }
        }
        if (act) {
#pragma unroll
            for (int e = 0; e < W; e += 2) *reinterpret_cast<double2*>(&T[r][c0 + e]) = make_double2(d[e], d[e + 1]);
        }
    }
    __syncthreads();
}

template <int VAR>
__global__ void __launch_bounds__(kDenseThreads, 1) bench(const double* A, double* out, long long* cyc, int reps)
{
    __shared__ double T[kCB][kCBs];
    __shared__ double buf[8 * kCB];
    __shared__ double d0s[kCB];
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int i = threadIdx.x; i < kCB * kCB; i += blockDim.x) T[i / kCB][i % kCB] = A[i];
        if (threadIdx.x < kCB) d0s[threadIdx.x] = A[threadIdx.x * kCB + threadIdx.x];
        __syncthreads();
        const long long t0 = clock64();
        if (VAR == 0) tile_invert_smem(T, buf, d0s); else if (VAR == 1) tile_invert_smem2<0>(T, buf, d0s); else if (VAR == 2) tile_invert_smemW<8>(T, buf, d0s); else if (VAR == 3) tile_invert_smemW<4>(T, buf, d0s); else if (VAR == 4) tile_invert_smemW<16, 0, 1>(T, buf, d0s); else if (VAR == 5) tile_invert_smem2<1>(T, buf, d0s); else if (VAR == 6) tile_invert_smemW<4, 0, 2>(T, buf, d0s); else tile_invert_smemW<16, 0, 2>(T, buf, d0s);
        total += clock64() - t0;
    }
    for (int i = threadIdx.x; i < kCB * kCB; i += blockDim.x) out[i] = T[i / kCB][i % kCB];
    if (threadIdx.x == 0) *cyc = total / reps;
}
int main()
{
    std::vector<double> B(kCB * kCB), A(kCB * kCB, 0.0);
    srand(3);
    for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < kCB; ++i) for (int j = 0; j < kCB; ++j) { double s = i == j ? 1.0 : 0.0; for (int k = 0; k < kCB; ++k) s += B[i * kCB + k] * B[j * kCB + k]; A[i * kCB + j] = s; }
    double *dA, *dO; long long* dC;
    cudaMalloc(&dA, sizeof(double) * kCB * kCB); cudaMalloc(&dO, sizeof(double) * kCB * kCB); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), sizeof(double) * kCB * kCB, cudaMemcpyHostToDevice);
    std::vector<double> O0(kCB * kCB), O1(kCB * kCB), O2(kCB * kCB);
    const char* names[8] = {"scalar, 16 cols/thread (5 warps)", "2x2 block, 16 cols/thread (5 warps)", "scalar, 8 cols/thread (9 warps)", "scalar, 4 cols/thread (18 warps)", "segment-major 16 cols (6 warps, padded)", "2x2 block, segment-major 16 cols (5 warps)", "segment-major 4 cols (18 warps)", "segment-major 16 cols (5 warps, packed)"};
    for (int pass = 0; pass < 2; ++pass) {                 // pass 1: DOFs 5 and 20 are dependent on others (collapsed pivots -> dropped)
        if (pass == 1) {
            for (int j = 0; j < kCB; ++j) { A[5 * kCB + j] = A[4 * kCB + j]; A[20 * kCB + j] = 0.0; }
            for (int i = 0; i < kCB; ++i) { A[i * kCB + 5] = A[i * kCB + 4]; A[i * kCB + 20] = 0.0; }
            A[5 * kCB + 5] = A[4 * kCB + 4];
            cudaMemcpy(dA, A.data(), sizeof(double) * kCB * kCB, cudaMemcpyHostToDevice);
        }
        for (int var = 0; var < 8; ++var) {
            if (var == 0) bench<0><<<1, kDenseThreads>>>(dA, dO, dC, 20); else if (var == 1) bench<1><<<1, kDenseThreads>>>(dA, dO, dC, 20); else if (var == 2) bench<2><<<1, kDenseThreads>>>(dA, dO, dC, 20); else if (var == 3) bench<3><<<1, kDenseThreads>>>(dA, dO, dC, 20); else if (var == 4) bench<4><<<1, kDenseThreads>>>(dA, dO, dC, 20); else if (var == 5) bench<5><<<1, kDenseThreads>>>(dA, dO, dC, 20); else if (var == 6) bench<6><<<1, kDenseThreads>>>(dA, dO, dC, 20); else bench<7><<<1, kDenseThreads>>>(dA, dO, dC, 20);
            cudaDeviceSynchronize();
            long long c; std::vector<double>& O = var == 1 ? O1 : (var == 0 ? O0 : O2);
            cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(O.data(), dO, sizeof(double) * kCB * kCB, cudaMemcpyDeviceToHost);
            double err = 0.0;
            for (int i = 0; i < kCB; ++i) for (int j = 0; j < kCB; ++j) { double s = 0.0; for (int k = 0; k < kCB; ++k) s += A[i * kCB + k] * O[k * kCB + j]; err = fmax(err, fabs(s - (i == j ? 1.0 : 0.0))); }
            printf("pass %d %-38s: %lld cycles per 48x48 tile = %.0f per pivot; |A inv(A) - I|_max = %.2e (%s)\n", pass, names[var], c, c / 48.0, err, cudaGetErrorString(cudaGetLastError()));
        }
        double dif = 0.0, mx = 0.0;
        for (int i = 0; i < kCB * kCB; ++i) { dif = fmax(dif, fabs(O0[i] - O1[i])); mx = fmax(mx, fabs(O0[i])); }
        printf("pass %d: max |scalar - block| = %.3e (max entry %.3e); zero rows of dropped DOFs: row5 %.1e row20 %.1e\n", pass, dif, mx, fabs(O1[5 * kCB + 7]) + fabs(O1[5 * kCB + 5]), fabs(O1[20 * kCB + 3]) + fabs(O1[20 * kCB + 20]));
    }
    return 0;
}
