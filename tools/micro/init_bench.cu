// fixed cost of a process that touches the GPU: context creation alone, then the first call into liboptcuts_b200.so
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <dlfcn.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv)
{
    const double t0 = now();
    cudaFree(0);
    const double t1 = now();
    void* p = nullptr; cudaMalloc(&p, 1 << 20); cudaMemset(p, 0, 1 << 20); cudaDeviceSynchronize();
    const double t2 = now();
    std::printf("context creation (cudaFree(0)) %.3f s, first malloc + memset + sync %.3f s\n", t1 - t0, t2 - t1);
    if (argc > 1) {
        void* h = dlopen(argv[1], RTLD_NOW);
        const double t3 = now();
        typedef int (*create_t)(void**, int); typedef int (*sync_t)(void*);
        create_t cr = (create_t)dlsym(h, "ocb_create"); sync_t sy = (sync_t)dlsym(h, "ocb_synchronize");
        void* ctx = nullptr;
        if (cr && sy) { cr(&ctx, 0); sy(ctx); }
        std::printf("dlopen %.3f s, ocb_create + first ocb_synchronize (library's own runtime: second context handle, stream, pinned block) %.3f s\n", t3 - t2, now() - t3);
    }
    return 0;
}
