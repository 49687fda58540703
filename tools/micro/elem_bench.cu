// Micro-benchmark: variants of the per-element energy pass (60 B of streamed data + 3 gathered UVs + ~100 fp64
// instructions per triangle) on a synthetic 1M-triangle grid mesh, to find what bounds the short element kernels.
//   nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a -I../../optcuts_b200/csrc elem_bench.cu -o elem_bench
// Every launch is timed alone with CUDA events after an L2 flush (256 MB memset); the median of 15 is printed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "ocb_element.cuh"
using namespace ocb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Elems { int n; const int* v0; const int* v1; const int* v2; const double* area; const double* A2; const double* e0; const double* e1; const double* d; double surf; };

template <int BLOCK>
__device__ __forceinline__ void finalize(double acc, double* partials, unsigned* ticket, double* out)
{
    __shared__ double sm[BLOCK / 32];
    __shared__ bool isLast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0; for (int w = 0; w < BLOCK / 32; ++w) t += sm[w];
        partials[blockIdx.x] = t;
        __threadfence();
        isLast = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    double a = 0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += BLOCK) a += __ldcg(&partials[b]);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __syncthreads();
    if (lane == 0) sm[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < BLOCK / 32; ++w) t += sm[w]; *out = t; *ticket = 0u; }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp4(void* s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(s)), "l"(g) : "memory"); }
__device__ __forceinline__ void cp8(void* s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(s)), "l"(g) : "memory"); }
__device__ __forceinline__ void cp16(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(s)), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ double elem_energy(const Elems& S, int i0, int i1, int i2, double area, double A2, double e0, double e1, double d, const double* x)
{
    const Vec2 U1 = ld2(x, i0), U2 = ld2(x, i1), U3 = ld2(x, i2);
    double db;
    return sd_energy(U2 - U1, U3 - U1, A2, e0, e1, d, area / S.surf, db);
}

// A: plain grid-stride register loop
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) varA(Elems S, const double* __restrict__ x, double* partials, unsigned* ticket, double* out)
{
    double acc = 0;
    for (int t = blockIdx.x * BLOCK + threadIdx.x; t < S.n; t += gridDim.x * BLOCK)
        acc += elem_energy(S, S.v0[t], S.v1[t], S.v2[t], S.area[t], S.A2[t], S.e0[t], S.e1[t], S.d[t], x);
    finalize<BLOCK>(acc, partials, ticket, out);
}

// B: cp.async queue of the streamed data, DEPTH stages
template <int BLOCK, int DEPTH, bool MATH>
__global__ void __launch_bounds__(BLOCK) varB(Elems S, const double* __restrict__ x, double* partials, unsigned* ticket, double* out)
{
    __shared__ int qi[DEPTH][3][BLOCK];
    __shared__ double qd[DEPTH][5][BLOCK];
    const int tid = threadIdx.x, stride = gridDim.x * BLOCK;
    auto issue = [&](int t, int st) {
        if (t < S.n) {
            cp4(&qi[st][0][tid], S.v0 + t); cp4(&qi[st][1][tid], S.v1 + t); cp4(&qi[st][2][tid], S.v2 + t);
            cp8(&qd[st][0][tid], S.area + t); cp8(&qd[st][1][tid], S.A2 + t); cp8(&qd[st][2][tid], S.e0 + t);
            cp8(&qd[st][3][tid], S.e1 + t); cp8(&qd[st][4][tid], S.d + t);
        }
        cp_commit();
    };
    double acc = 0;
    int t = blockIdx.x * BLOCK + tid, st = 0;
#pragma unroll
    for (int k = 0; k < DEPTH - 1; ++k) issue(t + k * stride, k);
    for (; t < S.n; t += stride, st = (st + 1 == DEPTH) ? 0 : st + 1) {
        issue(t + (DEPTH - 1) * stride, (st + DEPTH - 1) % DEPTH);
        cp_wait<DEPTH - 1>();
        const int i0 = qi[st][0][tid], i1 = qi[st][1][tid], i2 = qi[st][2][tid];
        if (MATH) acc += elem_energy(S, i0, i1, i2, qd[st][0][tid], qd[st][1][tid], qd[st][2][tid], qd[st][3][tid], qd[st][4][tid], x);
        else {
            const Vec2 U1 = ld2(x, i0), U2 = ld2(x, i1), U3 = ld2(x, i2);
            acc += qd[st][0][tid] + qd[st][1][tid] + qd[st][2][tid] + qd[st][3][tid] + qd[st][4][tid] + U1.x + U2.y + U3.x;
        }
    }
    finalize<BLOCK>(acc, partials, ticket, out);
}

// C: queue + the UVs of the NEXT element gathered into registers before the current one is computed
template <int BLOCK, int DEPTH>
__global__ void __launch_bounds__(BLOCK) varC(Elems S, const double* __restrict__ x, double* partials, unsigned* ticket, double* out)
{
    __shared__ int qi[DEPTH][3][BLOCK];
    __shared__ double qd[DEPTH][5][BLOCK];
    const int tid = threadIdx.x, stride = gridDim.x * BLOCK;
    auto issue = [&](int t, int st) {
        if (t < S.n) {
            cp4(&qi[st][0][tid], S.v0 + t); cp4(&qi[st][1][tid], S.v1 + t); cp4(&qi[st][2][tid], S.v2 + t);
            cp8(&qd[st][0][tid], S.area + t); cp8(&qd[st][1][tid], S.A2 + t); cp8(&qd[st][2][tid], S.e0 + t);
            cp8(&qd[st][3][tid], S.e1 + t); cp8(&qd[st][4][tid], S.d + t);
        }
        cp_commit();
    };
    double acc = 0;
    int t = blockIdx.x * BLOCK + tid, st = 0;
#pragma unroll
    for (int k = 0; k < DEPTH - 1; ++k) issue(t + k * stride, k);
    Vec2 U1 = mk(0, 0), U2 = U1, U3 = U1;
    cp_wait<DEPTH - 2>();
    if (t < S.n) { U1 = ld2(x, qi[0][0][tid]); U2 = ld2(x, qi[0][1][tid]); U3 = ld2(x, qi[0][2][tid]); }
    for (; t < S.n; t += stride, st = (st + 1 == DEPTH) ? 0 : st + 1) {
        issue(t + (DEPTH - 1) * stride, (st + DEPTH - 1) % DEPTH);
        cp_wait<DEPTH - 2>();                     // stage st+1 has landed too
        const int sn = (st + 1 == DEPTH) ? 0 : st + 1;
        Vec2 N1 = U1, N2 = U2, N3 = U3;
        if (t + stride < S.n) { N1 = ld2(x, qi[sn][0][tid]); N2 = ld2(x, qi[sn][1][tid]); N3 = ld2(x, qi[sn][2][tid]); }
        double db;
        acc += sd_energy(U2 - U1, U3 - U1, qd[st][1][tid], qd[st][2][tid], qd[st][3][tid], qd[st][4][tid], qd[st][0][tid] / S.surf, db);
        U1 = N1; U2 = N2; U3 = N3;
    }
    finalize<BLOCK>(acc, partials, ticket, out);
}

// D: one element per thread, no loop
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) varD(Elems S, const double* __restrict__ x, double* partials, unsigned* ticket, double* out)
{
    double acc = 0;
    const int t = blockIdx.x * BLOCK + threadIdx.x;
    if (t < S.n) acc = elem_energy(S, S.v0[t], S.v1[t], S.v2[t], S.area[t], S.A2[t], S.e0[t], S.e1[t], S.d[t], x);
    finalize<BLOCK>(acc, partials, ticket, out);
}

// E: the streamed arrays only (ceiling of a 60 MB read in one short launch)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) varE(Elems S, const double* __restrict__ x, double* partials, unsigned* ticket, double* out)
{
    double acc = 0;
    for (int t = blockIdx.x * BLOCK + threadIdx.x; t < S.n; t += gridDim.x * BLOCK)
        acc += S.area[t] + S.A2[t] + S.e0[t] + S.e1[t] + S.d[t] + (double)(S.v0[t] + S.v1[t] + S.v2[t]);
    finalize<BLOCK>(acc, partials, ticket, out);
}

// F: register loop, two elements in flight
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) varF(Elems S, const double* __restrict__ x, double* partials, unsigned* ticket, double* out)
{
    double acc = 0;
    const int stride = gridDim.x * BLOCK;
    int t = blockIdx.x * BLOCK + threadIdx.x;
    for (; t + stride < S.n; t += 2 * stride) {
        const int u = t + stride;
        const int a0 = S.v0[t], a1 = S.v1[t], a2 = S.v2[t], b0 = S.v0[u], b1 = S.v1[u], b2 = S.v2[u];
        const double ar = S.area[t], aA = S.A2[t], ae0 = S.e0[t], ae1 = S.e1[t], ad = S.d[t];
        const double br = S.area[u], bA = S.A2[u], be0 = S.e0[u], be1 = S.e1[u], bd = S.d[u];
        const Vec2 P1 = ld2(x, a0), P2 = ld2(x, a1), P3 = ld2(x, a2), Q1 = ld2(x, b0), Q2 = ld2(x, b1), Q3 = ld2(x, b2);
        double db;
        acc += sd_energy(P2 - P1, P3 - P1, aA, ae0, ae1, ad, ar / S.surf, db);
        acc += sd_energy(Q2 - Q1, Q3 - Q1, bA, be0, be1, bd, br / S.surf, db);
    }
    if (t < S.n) acc += elem_energy(S, S.v0[t], S.v1[t], S.v2[t], S.area[t], S.A2[t], S.e0[t], S.e1[t], S.d[t], x);
    finalize<BLOCK>(acc, partials, ticket, out);
}

template <typename K>
static void run(const char* name, K kern, int block, int grid, const Elems& S, const double* x, double* partials, unsigned* ticket, double* out, void* flush, size_t flushBytes)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    std::vector<float> ms;
    for (int it = 0; it < 18; ++it) {
        CK(cudaMemsetAsync(flush, it, flushBytes));
        CK(cudaEventRecord(a));
        kern<<<grid, block>>>(S, x, partials, ticket, out);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float t; CK(cudaEventElapsedTime(&t, a, b));
        if (it >= 3) ms.push_back(t);
    }
    CK(cudaGetLastError());
    std::sort(ms.begin(), ms.end());
    double h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    const double bytes = 60.0 * S.n;
    printf("%-44s grid %5d x %3d : median %7.2f us  min %7.2f us  -> %6.0f GB/s (60 B/tri)   sum %.10g\n", name, grid, block,
           ms[ms.size() / 2] * 1e3, ms[0] * 1e3, bytes / (ms[ms.size() / 2] * 1e-3) * 1e-9, h);
}

template <typename K>
static int resident(K kern, int block) { int b = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, block, 0)); return b; }

int main(int argc, char** argv)
{
    const int NX = argc > 1 ? atoi(argv[1]) : 1000, NY = argc > 2 ? atoi(argv[2]) : 500;
    const int nV = (NX + 1) * (NY + 1), nF = 2 * NX * NY;
    std::vector<int> v0(nF), v1(nF), v2(nF);
    std::vector<double> rest(5 * (size_t)nF), x(2 * (size_t)nV);
    srand(1);
    auto rnd = []() { return rand() / (double)RAND_MAX; };
    for (int j = 0; j <= NY; ++j) for (int i = 0; i <= NX; ++i) { const int v = j * (NX + 1) + i; x[2 * v] = i + 0.2 * rnd(); x[2 * v + 1] = j + 0.2 * rnd(); }
    for (int j = 0; j < NY; ++j) for (int i = 0; i < NX; ++i) {
        const int a = j * (NX + 1) + i, b = a + 1, c = a + NX + 1, d = c + 1, t = 2 * (j * NX + i);
        v0[t] = a; v1[t] = b; v2[t] = d; v0[t + 1] = a; v1[t + 1] = d; v2[t + 1] = c;
    }
    for (int t = 0; t < nF; ++t) {
        rest[t] = 0.5; rest[nF + t] = 0.25; rest[2 * (size_t)nF + t] = 1.0 + 0.1 * rnd(); rest[3 * (size_t)nF + t] = 2.0 + 0.1 * rnd(); rest[4 * (size_t)nF + t] = 1.0;
    }
    int* dI; double* dR; double* dX; double* partials; unsigned* ticket; double* out; void* flush;
    const size_t flushBytes = 256u << 20;
    CK(cudaMalloc(&dI, 3 * (size_t)nF * 4)); CK(cudaMalloc(&dR, 5 * (size_t)nF * 8)); CK(cudaMalloc(&dX, 2 * (size_t)nV * 8));
    CK(cudaMalloc(&partials, 8 * 65536)); CK(cudaMalloc(&ticket, 4)); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&flush, flushBytes));
    CK(cudaMemset(ticket, 0, 4));
    CK(cudaMemcpy(dI, v0.data(), (size_t)nF * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dI + nF, v1.data(), (size_t)nF * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dI + 2 * (size_t)nF, v2.data(), (size_t)nF * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR, rest.data(), rest.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, x.data(), x.size() * 8, cudaMemcpyHostToDevice));
    Elems S{nF, dI, dI + nF, dI + 2 * (size_t)nF, dR, dR + nF, dR + 2 * (size_t)nF, dR + 3 * (size_t)nF, dR + 4 * (size_t)nF, 0.5 * nF};
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs, %d triangles, %d vertices\n", prop.name, sms, nF, nV);
#define RUN(name, kern, block, grid) run(name, kern, block, grid, S, dX, partials, ticket, out, flush, flushBytes)
    RUN("A register loop, 8 CTAs/SM", varA<256>, 256, sms * 8);
    RUN("A register loop, resident wave", varA<256>, 256, sms * resident(varA<256>, 256));
    RUN("A register loop, 128 thr, resident wave", varA<128>, 128, sms * resident(varA<128>, 128));
    RUN("B queue depth 3, resident wave", (varB<256, 3, true>), 256, sms * resident(varB<256, 3, true>, 256));
    RUN("B queue depth 4, 128 thr, resident wave", (varB<128, 4, true>), 128, sms * resident(varB<128, 4, true>, 128));
    RUN("B queue depth 6, 128 thr, resident wave", (varB<128, 6, true>), 128, sms * resident(varB<128, 6, true>, 128));
    RUN("B queue depth 3, no math", (varB<256, 3, false>), 256, sms * resident(varB<256, 3, false>, 256));
    RUN("C queue 3 + next UVs in registers", (varC<256, 3>), 256, sms * resident(varC<256, 3>, 256));
    RUN("C queue 4 + next UVs, 128 thr", (varC<128, 4>), 128, sms * resident(varC<128, 4>, 128));
    RUN("D one element per thread", varD<256>, 256, (nF + 255) / 256);
    RUN("D one element per thread, 128 thr", varD<128>, 128, (nF + 127) / 128);
    RUN("E streamed arrays only, 8 CTAs/SM", varE<256>, 256, sms * 8);
    RUN("E streamed arrays only, one per thread", varE<256>, 256, (nF + 255) / 256);
    RUN("F register loop, 2 elements in flight", varF<256>, 256, sms * resident(varF<256>, 256));
    return 0;
}
