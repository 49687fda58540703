// Micro-benchmark: latency (dependent chain, one warp) and throughput (32 warps x 4 independent chains) per SM of the
// fp64 instructions the solver leans on.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_bench.cu -o fp64_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ void step(double& a, double& b, float& f, double x, double y)
{
    if (OP == 0) a = fma(a, x, y);                                      // DFMA
    if (OP == 1) f = fmaf(f, (float)x, (float)y);                       // FFMA
    if (OP == 2) a = (double)(float)a * x;                              // F2F.F32.F64 + F2F.F64.F32 + DMUL
    if (OP == 3) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(a), "+d"(b) : "d"(x), "d"(y));
    if (OP == 4) { double t; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(t) : "d"(a)); a = t + y; }   // MUFU.RCP64H + DADD
    if (OP == 5) a = a + x;                                             // DADD
}

template <int OP, int CHAINS>
__global__ void bench(double* out, long long* cyc, int iters, double x, double y)
{
    double a[CHAINS], b[CHAINS]; float f[CHAINS];
    for (int c = 0; c < CHAINS; ++c) { a[c] = 1.0 + threadIdx.x * 1e-9 + c; b[c] = 0.5; f[c] = 1.0f + c; }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) step<OP>(a[c], b[c], f[c], x, y);
    }
    const long long t1 = clock64();
    double s = 0.0;
    for (int c = 0; c < CHAINS; ++c) s += a[c] + b[c] + f[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double flopsPerInstrLane)
{
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    long long h;
    bench<OP, 1><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double lat = (double)h / (iters * 8);
    bench<OP, 4><<<1, 1024>>>(out, cyc, iters, 0.999999, 1e-9); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double perClk = 1024.0 * 4 * iters * 8 / (double)h;           // thread-instructions per clock per SM
    printf("%-28s latency %7.1f cycles   throughput %7.2f thread-instr/clk/SM  (%.1f useful flop-lanes/clk/SM)\n", name, lat, perClk, perClk * flopsPerInstrLane);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("DFMA", 1);
    run<5>("DADD", 1);
    run<1>("FFMA", 1);
    run<2>("F2F.f32<-f64 + F2F.f64<-f32 + DMUL", 1);
    run<3>("DMMA m8n8k4 (256 FMA/warp)", 8);
    run<4>("MUFU.RCP64H + DADD", 1);
    return 0;
}
