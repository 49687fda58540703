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

__global__ void __launch_bounds__(kDenseThreads, 1) bench(const double* A, double* out, long long* cyc, int reps)
{
    __shared__ double T[kCB][kCBs];
    __shared__ double buf[4 * kCB];
    __shared__ double d0s[kCB];
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int i = threadIdx.x; i < kCB * kCB; i += blockDim.x) T[i / kCB][i % kCB] = A[i];
        if (threadIdx.x < kCB) d0s[threadIdx.x] = A[threadIdx.x * kCB + threadIdx.x];
        __syncthreads();
        const long long t0 = clock64();
        tile_invert_smem(T, buf, d0s);
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
    bench<<<1, kDenseThreads>>>(dA, dO, dC, 20);
    cudaDeviceSynchronize();
    long long c; std::vector<double> O(kCB * kCB);
    cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(O.data(), dO, sizeof(double) * kCB * kCB, cudaMemcpyDeviceToHost);
    double err = 0.0;
    for (int i = 0; i < kCB; ++i) for (int j = 0; j < kCB; ++j) { double s = 0.0; for (int k = 0; k < kCB; ++k) s += A[i * kCB + k] * O[k * kCB + j]; err = fmax(err, fabs(s - (i == j ? 1.0 : 0.0))); }
    printf("tile inversion: %lld cycles per 48x48 tile = %.0f per pivot; |A inv(A) - I|_max = %.2e (%s)\n", c, c / 48.0, err, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
