// micro-benchmark: latency of the grid-wide all-reduce used by the persistent PCG kernel, per variant
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
struct __align__(16) SyncPacket { double v; unsigned long long epoch; };
struct __align__(64) SyncSlot { SyncPacket p[2]; unsigned long long pad[4]; };
__device__ __forceinline__ SyncPacket ld_packet(const SyncPacket* p) {
    SyncPacket r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<unsigned long long*>(&r.v)), "=l"(r.epoch) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_packet(SyncPacket* p, double v, unsigned long long epoch) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(__double_as_longlong(v)), "l"(epoch) : "memory");
}
template <int FENCE>   // 0 none, 1 __threadfence, 2 fence.acq_rel.gpu
__global__ void k_allreduce(SyncSlot* slots, int iters, double* out, double* work, int workPerThread)
{
    __shared__ double bc;
    const int nB = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc_total = 0.0;
    for (int it = 1; it <= iters; ++it) {
        // some global stores per thread, as the solver's phases do
        for (int w = 0; w < workPerThread; ++w) work[((size_t)blockIdx.x * blockDim.x + threadIdx.x) * workPerThread + w] = it + w;
        __syncthreads();
        if (warp == 0) {
            SyncSlot* base = slots + (size_t)(it & 1) * nB;
            if (lane == 0) {
                if (FENCE == 1) __threadfence();
                if (FENCE == 2) asm volatile("fence.acq_rel.gpu;" ::: "memory");
                st_packet(&base[blockIdx.x].p[0], 1.0, it);
            }
            double acc = 0.0;
            for (int b = lane; b < nB; b += 32) {
                SyncPacket q = ld_packet(&base[b].p[0]);
                while (q.epoch < (unsigned long long)it) q = ld_packet(&base[b].p[0]);
                acc += q.v;
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) bc = acc;
        }
        __syncthreads();
        acc_total += bc;
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = acc_total;
}
// variant Y: atomic arrival; the LAST arriver sums all partials in fixed order and publishes the result packets
template <int NV>
__global__ void k_lastarriver(SyncSlot* slots, unsigned long long* counter, SyncPacket* result, int iters, double* out, double* work, int workPerThread)
{
    __shared__ double bc[NV];
    const int nB = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc_total = 0.0;
    for (int it = 1; it <= iters; ++it) {
        for (int w = 0; w < workPerThread; ++w) work[((size_t)blockIdx.x * blockDim.x + threadIdx.x) * workPerThread + w] = it + w;
        __syncthreads();
        if (warp == 0) {
            SyncSlot* base = slots + (size_t)(it & 1) * nB;
            SyncPacket* res = result + (size_t)(it & 1) * NV;
            unsigned long long old = 0;
            if (lane == 0) {
                for (int k = 0; k < NV; ++k) base[blockIdx.x].p[k].v = 1.0 + k;
                __threadfence();
                old = atomicAdd(counter, 1ull);
            }
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == (unsigned long long)nB * it - 1) {      // last arriver
                __threadfence();
                double acc[NV];
                for (int k = 0; k < NV; ++k) acc[k] = 0.0;
                for (int b = lane; b < nB; b += 32)
                    for (int k = 0; k < NV; ++k) acc[k] += __ldcg(&base[b].p[k].v);
                for (int k = 0; k < NV; ++k) {
                    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
                    if (lane == 0) st_packet(&res[k], acc[k], it);
                }
            }
            if (lane < NV) {
                SyncPacket q = ld_packet(&res[lane]);
                while (q.epoch < (unsigned long long)it) q = ld_packet(&res[lane]);
                bc[lane] = q.v;
            }
        }
        __syncthreads();
        acc_total += bc[0];
        __syncthreads();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = acc_total;
}
__global__ void k_cg_sync(int iters, double* out)
{
    namespace cg = cooperative_groups;
    cg::grid_group g = cg::this_grid();
    double a = 0;
    for (int it = 0; it < iters; ++it) { g.sync(); a += 1.0; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = a;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    SyncSlot* slots; double* out; double* work;
    cudaMalloc(&slots, sizeof(SyncSlot) * 2 * 1024); cudaMalloc(&out, 8); cudaMalloc(&work, 8ull * 148 * 1024 * 8);
    const int iters = 2000;
    const int grids[] = {16, 41, 74, 148}; const int blocks[] = {256, 1024};
    for (int g : grids) for (int bs : blocks) for (int wp : {0, 2}) {
        float t[3];
        for (int f = 0; f < 3; ++f) {
            auto run = [&]() {
                cudaMemset(slots, 0, sizeof(SyncSlot) * 2 * 1024);
                void* args[] = {&slots, (void*)&iters, &out, &work, (void*)&wp};
                if (f == 0) cudaLaunchCooperativeKernel((void*)k_allreduce<0>, dim3(g), dim3(bs), args, 0, 0);
                if (f == 1) cudaLaunchCooperativeKernel((void*)k_allreduce<1>, dim3(g), dim3(bs), args, 0, 0);
                if (f == 2) cudaLaunchCooperativeKernel((void*)k_allreduce<2>, dim3(g), dim3(bs), args, 0, 0);
            };
            t[f] = timeit(run);
        }
        unsigned long long* counter; SyncPacket* result; cudaMalloc(&counter, 8); cudaMalloc(&result, 64 * 4);
        auto runY = [&]() {
            cudaMemset(counter, 0, 8); cudaMemset(result, 0, 256); cudaMemset(slots, 0, sizeof(SyncSlot) * 2 * 1024);
            void* args[] = {&slots, &counter, &result, (void*)&iters, &out, &work, (void*)&wp};
            cudaLaunchCooperativeKernel((void*)k_lastarriver<2>, dim3(g), dim3(bs), args, 0, 0);
        };
        float tY = timeit(runY);
        double hout = 0; cudaMemcpy(&hout, out, 8, cudaMemcpyDeviceToHost);
        printf("   variant Y (last arriver reduces, NV=2): %.2f us   check %.1f (want %.1f)\n", 1e3 * tY / iters, hout, (double)iters * g);
        cudaFree(counter); cudaFree(result);
        auto runcg = [&]() { void* args[] = {(void*)&iters, &out}; cudaLaunchCooperativeKernel((void*)k_cg_sync, dim3(g), dim3(bs), args, 0, 0); };
        float tcg = timeit(runcg);
        printf("grid %3d block %4d stores/thread %d : us per allreduce  nofence %.2f  threadfence %.2f  acq_rel %.2f   | cg grid.sync %.2f  (%s)\n",
               g, bs, wp, 1e3 * t[0] / iters, 1e3 * t[1] / iters, 1e3 * t[2] / iters, 1e3 * tcg / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
