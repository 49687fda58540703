// optcuts_b200 — the C-ABI (include/optcuts_b200.h): host-side orchestration of the device kernels.
// Mirrors, call for call, what the reference's Optimizer does around its Energy / LinSysSolver
// plugins (Optimizer.cpp:154-201, 203-261, 505-704, 764-843); every entry point cites its
// reference counterpart in the header.
#include "ocb_internal.cuh"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <unistd.h>
#include <algorithm>
#include <cmath>
#include <new>

namespace ocb {

int set_err(ocb_ctx* c, int code, const char* what)
{
    if (c) c->err = what ? what : "";
    return code;
}
int cuda_fail(ocb_ctx* c, cudaError_t e, const char* where)
{
    if (c) { c->err = std::string(where) + ": " + cudaGetErrorString(e); }
    cudaGetLastError();
    return OCB_ERR_CUDA;
}

static cudaEvent_t prof_event(ocb_ctx* c)
{
    cudaEvent_t e = nullptr;
    if (!c->profPool.empty()) { e = c->profPool.back(); c->profPool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
ProfScope::ProfScope(ocb_ctx* ctx, int k) : c(ctx), cls(k)
{
    if (!c->prof) return;
    a = prof_event(c); b = prof_event(c);
    cudaEventRecord(a, c->stream);
}
ProfScope::~ProfScope()
{
    if (!a) return;
    cudaEventRecord(b, c->stream);
    ocb_ctx::ProfRec r; r.cls = cls; r.a = a; r.b = b;
    c->profRecs.push_back(r);
}
static void prof_collect(ocb_ctx* c)
{
    if (c->profRecs.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& r : c->profRecs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->profMs[r.cls] += ms; c->profCnt[r.cls]++; }
        c->profPool.push_back(r.a); c->profPool.push_back(r.b);
    }
    c->profRecs.clear();
}

static bool host_timing_on() { static const bool on = []() { const char* e = getenv("OCB_HOST_TIMING"); return e && atoi(e); }(); return on; }
struct HostTimingRec { const char* name; double total; long count; };
static std::vector<HostTimingRec>& host_timing_table() { static std::vector<HostTimingRec>* t = new std::vector<HostTimingRec>(); return *t; }   // never destroyed: the report runs from atexit
static double wall_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
HostTimer::HostTimer(const char* n) : name(n), t0(host_timing_on() ? wall_now() : 0.0)
{
    static const bool reg = []() { if (host_timing_on()) atexit(host_timing_report); return true; }();     // host programs that exit() without destroying their contexts
    (void)reg;
}
HostTimer::~HostTimer()
{
    if (!host_timing_on()) return;
    const double dt = wall_now() - t0;
    for (auto& r : host_timing_table()) if (r.name == name) { r.total += dt; r.count++; return; }
    host_timing_table().push_back(HostTimingRec{name, dt, 1});
}
// seconds since this process started (Linux: field 22 of /proc/self/stat against the boot-time clock)
static double process_age_s()
{
    FILE* f = fopen("/proc/self/stat", "r");
    if (!f) return -1.0;
    char buf[2048]; const size_t n = fread(buf, 1, sizeof(buf) - 1, f); fclose(f); buf[n] = 0;
    const char* p = strrchr(buf, ')');                        // the command name may contain spaces
    if (!p) return -1.0;
    unsigned long long start = 0; int field = 2;
    for (p += 2; *p && field < 21; ++p) if (*p == ' ') ++field;
    sscanf(p, "%llu", &start);
    struct timespec ts; clock_gettime(CLOCK_BOOTTIME, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec - (double)start / (double)sysconf(_SC_CLK_TCK);
}
void host_timing_report()
{
    if (!host_timing_on()) return;
    fprintf(stderr, "[ocb host] process age at the report: %.3f s\n", process_age_s());
    for (auto& r : host_timing_table()) fprintf(stderr, "[ocb host] %-28s %8ld calls  %10.3f ms total  %8.3f us/call\n", r.name, r.count, 1e3 * r.total, 1e6 * r.total / (r.count ? r.count : 1));
    host_timing_table().clear();
}

int ensure_init(ocb_ctx* c)
{
    if (c->inited) { cudaSetDevice(c->device); return 0; }
    HostTimer _hc("cuda context + stream (first call)");
    if (host_timing_on()) fprintf(stderr, "[ocb host] process age at the first CUDA call: %.3f s\n", process_age_s());
    OCB_CUDA(c, cudaSetDevice(c->device));
    OCB_CUDA(c, cudaFree(0));                                // creates the primary context here (timed above), not in the first allocation
    if (!c->streamGiven) { OCB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    OCB_CUDA(c, cudaEventCreate(&c->ev0));
    OCB_CUDA(c, cudaEventCreate(&c->ev1));
    cudaDeviceProp prop;
    OCB_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
    c->numSMs = prop.multiProcessorCount;
    OCB_CUDA(c, cudaMalloc((void**)&c->dScal, sizeof(double) * S_COUNT));
    OCB_CUDA(c, cudaMemset(c->dScal, 0, sizeof(double) * S_COUNT));
    OCB_CUDA(c, cudaMallocHost((void**)&c->hScal, sizeof(double) * S_COUNT));
    OCB_CUDA(c, c->sync.reserve(64, c->stream));
    OCB_CUDA(c, cudaMemset(c->sync.p, 0, sizeof(unsigned) * 64));
    OCB_CUDA(c, c->partials.reserve(4096, c->stream));
    c->inited = true;
    return 0;
}

ElemView view_of(const ocb_ctx* c, const ElemSet& s, bool isAir, double scale, int uniform)
{
    ElemView v;
    const size_t n = (size_t)s.n;
    v.n = s.n;
    v.v0 = s.v.p; v.v1 = s.v.p + n; v.v2 = s.v.p + 2 * n;
    const double* r = s.rest.p;
    v.area = r; v.areaSq = r + n; v.e0 = r + 2 * n; v.e1 = r + 3 * n; v.d = r + 4 * n;
    v.k0 = r + 5 * n; v.k1 = r + 6 * n; v.kd = r + 7 * n;
    v.slot = s.slot.p; v.rec = s.rec.p; v.vcSlot = s.vcSlot.p;
    v.surfaceArea = isAir ? 1.0 : c->surfaceArea;
    v.uniform = uniform;
    v.scale = scale;
    return v;
}

int fetch_scalars(ocb_ctx* c)
{
    OCB_CUDA(c, cudaMemcpyAsync(c->hScal, c->dScal, sizeof(double) * S_COUNT, cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

static int upload_d(ocb_ctx* c, double* dst, const double* src, size_t n) {
    OCB_CUDA(c, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
}
static int upload_i(ocb_ctx* c, int32_t* dst, const int32_t* src, size_t n) {
    OCB_CUDA(c, cudaMemcpyAsync(dst, src, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

// (re)size the system vectors after nVtot changed, keeping the mesh part of x
static int resize_system(ocb_ctx* c)
{
    const size_t n = (size_t)c->nSys();
    OCB_CUDA(c, c->x.reserve(n, c->stream, true));
    OCB_CUDA(c, c->x0.reserve(n, c->stream));
    OCB_CUDA(c, c->g.reserve(n, c->stream));
    OCB_CUDA(c, c->p.reserve(n, c->stream));
    return 0;
}

static int upload_fixed_mask(ocb_ctx* c)
{
    OCB_CUDA(c, c->fixedMask.reserve((size_t)c->nVtot, c->stream));
    c->hFixed.resize((size_t)c->nVtot, 0);
    OCB_CUDA(c, cudaMemcpyAsync(c->fixedMask.p, c->hFixed.data(), (size_t)c->nVtot, cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// Locality order of the mesh vertices (breadth-first over the vertex adjacency, Cuthill-McKee style): rows that
// are coupled end up close in memory, so (i) a CTA's contiguous row range of the solver is a compact patch whose
// halo is small, and (ii) the UV gathers of the element kernels hit the same sectors.  The caller never sees it:
// every vertex-indexed array is converted at the C-ABI boundary.  OCB_NO_REORDER=1 keeps the caller's order.
static void compute_order(ocb_ctx* c, int nV, int nF, const int32_t* F)
{
    static const bool off = []() { const char* e = getenv("OCB_NO_REORDER"); return e && atoi(e); }();
    c->hPerm.assign((size_t)nV, 0);
    if (off) { for (int v = 0; v < nV; ++v) c->hPerm[v] = v; return; }
    std::vector<int32_t> ptr((size_t)nV + 1, 0);
    for (int t = 0; t < nF; ++t) for (int k = 0; k < 3; ++k) ptr[(size_t)F[(size_t)k * nF + t] + 1] += 2;
    for (int v = 0; v < nV; ++v) ptr[v + 1] += ptr[v];
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1), adj((size_t)ptr[nV]);
    for (int t = 0; t < nF; ++t) {
        const int a = F[t], b = F[(size_t)nF + t], d = F[2 * (size_t)nF + t];
        adj[fill[a]++] = b; adj[fill[a]++] = d; adj[fill[b]++] = a; adj[fill[b]++] = d; adj[fill[d]++] = a; adj[fill[d]++] = b;
    }
    std::vector<int32_t> order; order.reserve((size_t)nV);
    std::vector<uint8_t> seen((size_t)nV, 0);
    for (int s0 = 0; s0 < nV; ++s0) {
        if (seen[s0]) continue;
        seen[s0] = 1; order.push_back(s0);
        for (size_t head = order.size() - 1; head < order.size(); ++head) {
            const int v = order[head];
            for (int k = ptr[v]; k < ptr[v + 1]; ++k) { const int u = adj[k]; if (!seen[u]) { seen[u] = 1; order.push_back(u); } }
        }
    }
    for (int i = 0; i < nV; ++i) c->hPerm[order[i]] = i;
}
static int upload_perm(ocb_ctx* c)
{
    c->hPerm.resize((size_t)c->nVtot);
    for (int v = c->nV; v < c->nVtot; ++v) c->hPerm[v] = v;          // air interior vertices keep nV + k
    c->hInv.assign((size_t)c->nVtot, 0);
    for (int v = 0; v < c->nVtot; ++v) c->hInv[c->hPerm[v]] = v;
    OCB_CUDA(c, c->perm.reserve((size_t)c->nVtot + 1, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(c->perm.p, c->hPerm.data(), sizeof(int32_t) * c->nVtot, cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
// system vector: caller order on the host <-> internal order on the device
static int vec_to_device(ocb_ctx* c, double* dst_internal, const double* host_user)
{
    const size_t n = (size_t)c->nSys();
    OCB_CUDA(c, c->scratchV.reserve(n, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(c->scratchV.p, host_user, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return launch_permute_vec(c, c->scratchV.p, dst_internal, true);
}
static int vec_to_host(ocb_ctx* c, double* host_user, const double* src_internal)
{
    const size_t n = (size_t)c->nSys();
    OCB_CUDA(c, c->scratchV.reserve(n, c->stream));
    OCB_TRY(launch_permute_vec(c, src_internal, c->scratchV.p, false));
    OCB_CUDA(c, cudaMemcpyAsync(host_user, c->scratchV.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return 0;
}

// vertex -> corner incidence (entries element << 2 | corner, ascending element order per vertex) of an element list given as
// 3 x n SoA over nRows vertex ids; uploaded to ptrBuf / idxBuf
static int upload_corner_csr(ocb_ctx* c, int nRows, int n, const int32_t* F_soa, ocb::DevBuf<int32_t>& ptrBuf, ocb::DevBuf<int32_t>& idxBuf)
{
    std::vector<int32_t> ptr((size_t)nRows + 1, 0), idx((size_t)3 * n);
    for (int k = 0; k < 3; ++k) for (int t = 0; t < n; ++t) ptr[(size_t)F_soa[(size_t)k * n + t] + 1]++;
    for (int v = 0; v < nRows; ++v) ptr[v + 1] += ptr[v];
    std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
    for (int t = 0; t < n; ++t) for (int k = 0; k < 3; ++k) idx[fill[F_soa[(size_t)k * n + t]]++] = (t << 2) | k;
    OCB_CUDA(c, ptrBuf.reserve((size_t)nRows + 2, c->stream));
    OCB_CUDA(c, idxBuf.reserve((size_t)3 * n + 2, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(ptrBuf.p, ptr.data(), sizeof(int32_t) * ((size_t)nRows + 1), cudaMemcpyHostToDevice, c->stream));
    if (n > 0) OCB_CUDA(c, cudaMemcpyAsync(idxBuf.p, idx.data(), sizeof(int32_t) * (size_t)3 * n, cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));       // the host vectors go out of scope
    return 0;
}

static int need(ocb_ctx* c, bool cond, const char* what) { return cond ? 0 : set_err(c, OCB_ERR_STATE, what); }

}  // namespace ocb

using namespace ocb;

extern "C" {

const char* ocb_version(void) { return "optcuts_b200 0.1 (sm_100a, fp64)"; }

int ocb_create(ocb_ctx** out, int device)
{
    if (!out) return OCB_ERR_ARG;
    ocb_ctx* c = new (std::nothrow) ocb_ctx();
    if (!c) return OCB_ERR_ARG;
    c->device = device;
    { const char* e = getenv("OCB_PCG_SCALED_NORM"); c->pcgPlainNorm = !(e && atoi(e)); }
    { const char* e = getenv("OCB_SCALE_SYSTEM"); if (e) c->scaleSystem = atoi(e) != 0; }
    { const char* e = getenv("OCB_MAS_EQUILIBRATE"); if (e) c->masEquilibrate = atoi(e) != 0; }
    *out = c;
    return OCB_OK;
}

void ocb_destroy(ocb_ctx* c)
{
    host_timing_report();
    if (!c) return;
    if (c->inited) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        c->mesh.v.release(); c->mesh.rest.release(); c->mesh.slot.release(); c->mesh.rec.release(); c->mesh.vcSlot.release(); c->air.rec.release(); c->air.vcSlot.release();
        c->air.v.release(); c->air.rest.release(); c->air.slot.release();
        c->l2g.release(); c->vcPtrM.release(); c->vcIdxM.release(); c->vcPtrA.release(); c->vcIdxA.release(); c->g2l.release(); c->hel.release(); c->fixedMask.release(); c->perm.release(); c->scratchV.release(); c->stD.release(); c->stI.release();
        c->x.release(); c->x0.release(); c->g.release(); c->p.release();
        c->pr.release(); c->pz.release(); c->pd.release(); c->pd2.release(); c->pAp.release(); c->pb.release(); c->minv.release();
        c->rowPtr.release(); c->colIdx.release(); c->val.release();
        c->partials.release(); c->sync.release(); c->scratchD.release(); c->scratchI.release();
        direct_release(c);
        c->xSaved.release(); c->rowScale.release(); c->pcgHalo.release(); c->px.release(); c->rowOf.release(); c->vertOf.release(); c->userRow.release();
        c->masD.ints.release(); c->masD.geom.release(); c->masD.val.release(); c->masD.rcCta.release(); c->masD.inv.release(); c->masD.vinfo.release();
        prof_collect(c);
        for (auto e : c->profPool) cudaEventDestroy(e);
        if (c->dScal) cudaFree(c->dScal);
        if (c->hScal) cudaFreeHost(c->hScal);
        if (c->stPinned) cudaFreeHost(c->stPinned);
        if (c->ev0) cudaEventDestroy(c->ev0);
        if (c->ev1) cudaEventDestroy(c->ev1);
        if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    }
    delete c;
}

const char* ocb_last_error(const ocb_ctx* c) { return c ? c->err.c_str() : "null context"; }

int ocb_set_stream(ocb_ctx* c, void* s)
{
    // the handle is used as given: NULL is the legacy default stream (what torch.cuda.current_stream().cuda_stream is for
    // torch's default stream), so the library's work is ordered with the caller's own work on that stream
    if (!c) return OCB_ERR_ARG;
    if (c->inited) cudaSetDevice(c->device);
    if (c->inited && c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)s; c->own_stream = false; c->streamGiven = true;
    return OCB_OK;
}
int ocb_use_own_stream(ocb_ctx* c)
{
    if (!c) return OCB_ERR_ARG;
    if (c->own_stream) return OCB_OK;
    if (c->inited) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    c->stream = nullptr; c->streamGiven = false;
    if (c->inited) { OCB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    return OCB_OK;
}
int ocb_set_option(ocb_ctx* c, const char* key, double value)
{
    if (!c || !key) return OCB_ERR_ARG;
    if (!std::strcmp(key, "pcg_scaled_norm")) { c->pcgPlainNorm = value == 0.0; return OCB_OK; }
    if (!std::strcmp(key, "scale_system")) { c->scaleSystem = value != 0.0; return OCB_OK; }
    if (!std::strcmp(key, "force_direct")) { c->forceDirect = value != 0.0; return OCB_OK; }
    if (!std::strcmp(key, "mas_equilibrate")) { c->masEquilibrate = value != 0.0; c->precondValid = false; return OCB_OK; }
    return set_err(c, OCB_ERR_ARG, "ocb_set_option: unknown key");
}
int ocb_synchronize(ocb_ctx* c) { OCB_TRY(ensure_init(c)); OCB_CUDA(c, cudaStreamSynchronize(c->stream)); return OCB_OK; }
int ocb_timer_start(ocb_ctx* c) { OCB_TRY(ensure_init(c)); OCB_CUDA(c, cudaEventRecord(c->ev0, c->stream)); return OCB_OK; }
int ocb_timer_stop_ms(ocb_ctx* c, double* ms)
{
    OCB_TRY(ensure_init(c));
    OCB_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    OCB_CUDA(c, cudaEventSynchronize(c->ev1));
    float f = 0.f;
    OCB_CUDA(c, cudaEventElapsedTime(&f, c->ev0, c->ev1));
    if (ms) *ms = f;
    return OCB_OK;
}
int64_t ocb_launch_count(const ocb_ctx* c) { return c ? c->launches : 0; }

int ocb_profile_enable(ocb_ctx* c, int on)
{
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));
    prof_collect(c);
    c->prof = on != 0;
    for (int k = 0; k < K_COUNT; ++k) { c->profMs[k] = 0.0; c->profCnt[k] = 0; }
    return OCB_OK;
}
int ocb_profile_get(ocb_ctx* c, double* ms, int64_t* counts)
{
    if (!c || !ms || !counts) return OCB_ERR_ARG;
    prof_collect(c);
    for (int k = 0; k < K_COUNT; ++k) { ms[k] = c->profMs[k]; counts[k] = c->profCnt[k]; }
    return OCB_OK;
}
const char* ocb_profile_name(int k)
{
    static const char* names[K_COUNT] = {"energy", "gradient", "hessian_psd_scatter", "pcg", "step_bound", "step_forward",
                                         "jacobi_setup", "spmv", "rest_features", "pattern_slots", "misc", "stencil_newton", "mas_setup", "hessian_rows"};
    return (k >= 0 && k < K_COUNT) ? names[k] : "";
}
int ocb_profile_classes(void) { return K_COUNT; }

// ------------------------------------------------------------------------------------------- a1
int ocb_rest_features(ocb_ctx* c, int nV, int nF, const double* Vrest, const int32_t* F, double thres,
                      double* rest8, double* scalars3)
{
    if (!c || nV <= 0 || nF <= 0 || !Vrest || !F || !rest8) return set_err(c, OCB_ERR_ARG, "ocb_rest_features: bad argument");
    OCB_TRY(ensure_init(c));
    OCB_CUDA(c, c->scratchD.reserve((size_t)3 * nV + (size_t)8 * nF, c->stream));
    OCB_CUDA(c, c->scratchI.reserve((size_t)3 * nF, c->stream));
    double* dP = c->scratchD.p; double* dR = dP + (size_t)3 * nV;
    OCB_TRY(upload_d(c, dP, Vrest, (size_t)3 * nV));
    OCB_TRY(upload_i(c, c->scratchI.p, F, (size_t)3 * nF));
    OCB_TRY(launch_rest_features(c, nV, nF, dP, c->scratchI.p, thres, dR));
    OCB_CUDA(c, cudaMemcpyAsync(rest8, dR, sizeof(double) * 8 * (size_t)nF, cudaMemcpyDeviceToHost, c->stream));
    OCB_TRY(fetch_scalars(c));
    const double surf = c->hScal[S_MISC0];
    if (scalars3) { scalars3[0] = surf; scalars3[1] = c->hScal[S_MISC1] / (3.0 * nF); scalars3[2] = std::sqrt(surf / M_PI); }
    if (c->hScal[S_MISC2] > 0.0) return set_err(c, OCB_ERR_ARG, "mesh has a zero-area triangle (TriMesh.cpp:368-402)");
    return OCB_OK;
}

// ------------------------------------------------------------------------------------- problem data
static int upload_elems(ocb_ctx* c, ElemSet& S, std::vector<int32_t>& hostCopy, int n, const int32_t* F_global_soa, const double* rest8)
{
    S.n = n;
    OCB_CUDA(c, S.v.reserve((size_t)3 * n + 1, c->stream));
    OCB_CUDA(c, S.rest.reserve((size_t)8 * n + 1, c->stream));
    if (n > 0) {
        hostCopy.assign(F_global_soa, F_global_soa + (size_t)3 * n);
        OCB_TRY(upload_i(c, S.v.p, hostCopy.data(), (size_t)3 * n));
        OCB_TRY(upload_d(c, S.rest.p, rest8, (size_t)8 * n));
        // the same data once more as one 64-byte record per triangle (vertex-gather kernels)
        OCB_CUDA(c, S.rec.reserve((size_t)8 * n + 8, c->stream));
        std::vector<double> rec((size_t)8 * n, 0.0);
        for (int t = 0; t < n; ++t) {
            double* r = rec.data() + 8 * (size_t)t;
            int32_t iv[4] = {F_global_soa[t], F_global_soa[(size_t)n + t], F_global_soa[2 * (size_t)n + t], 0};
            std::memcpy(r, iv, 16);
            for (int q = 0; q < 5; ++q) r[2 + q] = rest8[(size_t)q * n + t];
        }
        OCB_TRY(upload_d(c, S.rec.p, rec.data(), (size_t)8 * n));
    } else hostCopy.clear();
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int ocb_set_mesh(ocb_ctx* c, int nV, int nF, const int32_t* F, const double* rest8, double surfaceArea,
                 const int32_t* fixed, int nFixed)
{
    HostTimer _ht("set_mesh");
    // every argument is checked before the context is touched: a failed call leaves the previous problem intact
    if (!c || nV <= 0 || nF <= 0 || !F || !rest8 || !(surfaceArea > 0.0) || nFixed < 0 || (nFixed > 0 && !fixed))
        return set_err(c, OCB_ERR_ARG, "ocb_set_mesh: bad argument");
    if ((int64_t)nF >= (INT32_MAX >> 2)) return set_err(c, OCB_ERR_ARG, "ocb_set_mesh: more than 2^29 triangles");
    for (size_t i = 0; i < (size_t)3 * nF; ++i) if (F[i] < 0 || F[i] >= nV) return set_err(c, OCB_ERR_ARG, "ocb_set_mesh: vertex index out of range");
    for (int i = 0; i < nFixed; ++i) if (fixed[i] < 0 || fixed[i] >= nV) return set_err(c, OCB_ERR_ARG, "fixed vertex out of range");
    OCB_TRY(ensure_init(c));
    c->haveUV = false; c->patternValid = c->matrixValid = c->precondValid = c->slotsValid = false; c->gradValid = false;
    c->nV = nV; c->nF = nF; c->surfaceArea = surfaceArea;
    // a new mesh drops the scaffold (the caller re-sends it, as the reference rebuilds it: Optimizer.cpp:483-488)
    c->nVa = c->nFa = c->nBnd = 0; c->air.n = 0; c->hFa.clear(); c->hL2G.clear(); c->wScafOverFa = 0.0;
    c->nVtot = nV;
    c->hFuser.assign(F, F + (size_t)3 * nF);
    c->meshAdjValid = false;
    compute_order(c, nV, nF, F);
    OCB_TRY(upload_perm(c));
    {
        std::vector<int32_t> Fi((size_t)3 * nF);
        for (size_t i = 0; i < Fi.size(); ++i) Fi[i] = c->hPerm[F[i]];
        OCB_TRY(upload_elems(c, c->mesh, c->hF, nF, Fi.data(), rest8));
        OCB_TRY(upload_corner_csr(c, nV, nF, Fi.data(), c->vcPtrM, c->vcIdxM));
    }
    OCB_TRY(launch_g2l(c));
    c->hFixed.assign((size_t)nV, 0);
    for (int i = 0; i < nFixed; ++i) c->hFixed[c->hPerm[fixed[i]]] = 1;
    OCB_TRY(resize_system(c));
    OCB_TRY(upload_fixed_mask(c));
    c->hXY.clear(); c->hXYMesh = c->hXYAir = false;
    return OCB_OK;
}

int ocb_set_air(ocb_ctx* c, int nVa, int nFa, const int32_t* Fa, const double* rest8, const int32_t* l2g, int nBnd,
                const int32_t* fixedAir, int nFixedAir, double wScafOverFa)
{
    HostTimer _ht("set_air");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(need(c, c->nV > 0, "ocb_set_air before ocb_set_mesh"));
    if (nFa > 0) {      // validate everything before the context is touched
        if (!Fa || !rest8 || !l2g || nVa <= 0 || nBnd < 0 || nBnd > nVa || nFixedAir < 0 || (nFixedAir > 0 && !fixedAir))
            return set_err(c, OCB_ERR_ARG, "ocb_set_air: bad argument");
        for (int i = 0; i < nVa; ++i) {
            const bool ok = (i < nBnd) ? (l2g[i] >= 0 && l2g[i] < c->nV) : (l2g[i] == c->nV + i - nBnd);
            if (!ok) return set_err(c, OCB_ERR_ARG, "ocb_set_air: localVI2Global does not follow Scaffold.cpp:179-184");
        }
        for (size_t i = 0; i < (size_t)3 * nFa; ++i) if (Fa[i] < 0 || Fa[i] >= nVa) return set_err(c, OCB_ERR_ARG, "ocb_set_air: vertex index out of range");
        for (int i = 0; i < nFixedAir; ++i) if (fixedAir[i] < 0 || fixedAir[i] >= nVa) return set_err(c, OCB_ERR_ARG, "fixed air vertex out of range");
    }
    OCB_TRY(ensure_init(c));
    c->patternValid = c->matrixValid = c->precondValid = c->slotsValid = false; c->gradValid = false;
    // fixed flags: the mesh's own (bit 0) survive, what a previous air mesh set (bit 1) does not
    c->hFixed.resize((size_t)c->nV);
    for (auto& f : c->hFixed) f &= 1;
    if (nFa <= 0) {
        c->nVa = c->nFa = c->nBnd = 0; c->air.n = 0; c->hFa.clear(); c->hL2G.clear(); c->wScafOverFa = 0.0;
        c->nVtot = c->nV;
        OCB_TRY(upload_perm(c));
    } else {
        const int nVtot = c->nV + nVa - nBnd;
        c->hL2G.resize((size_t)nVa);                      // air-local id -> INTERNAL global id
        for (int i = 0; i < nVa; ++i) c->hL2G[i] = i < nBnd ? c->hPerm[l2g[i]] : l2g[i];
        std::vector<int32_t> Fg((size_t)3 * nFa);
        for (size_t i = 0; i < (size_t)3 * nFa; ++i) Fg[i] = c->hL2G[Fa[i]];
        c->nVa = nVa; c->nFa = nFa; c->nBnd = nBnd; c->nVtot = nVtot; c->wScafOverFa = wScafOverFa;
        OCB_TRY(upload_perm(c));
        OCB_TRY(upload_elems(c, c->air, c->hFa, nFa, Fg.data(), rest8));
        OCB_TRY(upload_corner_csr(c, nVa, nFa, Fa, c->vcPtrA, c->vcIdxA));
        OCB_CUDA(c, c->l2g.reserve((size_t)nVa, c->stream));
        OCB_TRY(upload_i(c, c->l2g.p, c->hL2G.data(), (size_t)nVa));
        c->hFixed.resize((size_t)nVtot, 0);
        for (int i = 0; i < nFixedAir; ++i) c->hFixed[c->hL2G[fixedAir[i]]] |= 2;
    }
    OCB_TRY(launch_g2l(c));
    c->hXYAir = false;                                   // the air vertices are new: their mirror entries are stale
    OCB_TRY(resize_system(c));
    OCB_TRY(upload_fixed_mask(c));
    return OCB_OK;
}

int ocb_set_uv(ocb_ctx* c, const double* V, const double* Va)
{
    HostTimer _ht("set_uv");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(need(c, c->nV > 0, "ocb_set_uv before ocb_set_mesh"));
    OCB_TRY(ensure_init(c));
    const size_t nv = V ? 2 * (size_t)c->nV : 0, na = (Va && c->nVa > 0) ? 2 * (size_t)c->nVa : 0;
    OCB_CUDA(c, c->scratchD.reserve(nv + na + 2, c->stream));
    double* dV = V ? c->scratchD.p : nullptr;
    double* dVa = na ? c->scratchD.p + nv : nullptr;
    if (V) OCB_TRY(upload_d(c, dV, V, nv));
    if (na) OCB_TRY(upload_d(c, dVa, Va, na));
    OCB_TRY(launch_set_uv(c, dV, dVa));
    {   // host mirror in the device's internal numbering (set_uv_kernel's mapping)
        const size_t nTot = (size_t)c->nVtot;
        if (c->hXY.size() != 2 * nTot) { c->hXY.resize(2 * nTot, 0.0); c->hXYAir = false; }      // the mesh part (ids < nV) survives a new air mesh
        if (V) {
            const int nV = c->nV;
            for (int k = 0; k < nV; ++k) { const size_t q = (size_t)c->hPerm[k]; c->hXY[2 * q] = V[k]; c->hXY[2 * q + 1] = V[(size_t)nV + k]; }
            c->hXYMesh = true;
        }
        if (na) {
            const int nVa = c->nVa, nB = c->nBnd, nV = c->nV;
            for (int k = 0; k < nVa - nB; ++k) { c->hXY[2 * ((size_t)nV + k)] = Va[nB + k]; c->hXY[2 * ((size_t)nV + k) + 1] = Va[(size_t)nVa + nB + k]; }
            c->hXYAir = true;
        }
    }
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (V) c->haveUV = true;
    c->matrixValid = c->precondValid = false; c->gradValid = false;
    return OCB_OK;
}

int ocb_get_uv(ocb_ctx* c, double* V, double* Va)
{
    HostTimer _ht("get_uv");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "no UV on the device"));
    const size_t nv = V ? 2 * (size_t)c->nV : 0, na = (Va && c->nVa > 0) ? 2 * (size_t)c->nVa : 0;
    OCB_CUDA(c, c->scratchD.reserve(nv + na + 2, c->stream));
    double* dV = V ? c->scratchD.p : nullptr;
    double* dVa = na ? c->scratchD.p + nv : nullptr;
    OCB_TRY(launch_get_uv(c, dV, dVa));
    if (V) OCB_CUDA(c, cudaMemcpyAsync(V, dV, nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (na) OCB_CUDA(c, cudaMemcpyAsync(Va, dVa, na * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

int ocb_save_uv(ocb_ctx* c)
{
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_save_uv: no UV on the device"));
    OCB_CUDA(c, c->xSaved.reserve((size_t)c->nSys(), c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(c->xSaved.p, c->x.p, sizeof(double) * c->nSys(), cudaMemcpyDeviceToDevice, c->stream));
    c->xSavedN = c->nSys();
    return OCB_OK;
}
int ocb_restore_uv(ocb_ctx* c)
{
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->xSavedN == c->nSys() && c->xSavedN > 0, "ocb_restore_uv: no snapshot of this system size"));
    OCB_CUDA(c, cudaMemcpyAsync(c->x.p, c->xSaved.p, sizeof(double) * c->nSys(), cudaMemcpyDeviceToDevice, c->stream));
    c->hXYMesh = c->hXYAir = false;
    c->matrixValid = c->precondValid = false; c->gradValid = false;
    return OCB_OK;
}

static int64_t nnz_upper(const ocb_ctx* c)
{
    // LinSysSolver::set_pattern: free vertex rows: 2*k and 2*k-1 entries with k = #block cols >= row; the total does
    // not depend on the (symmetric) ordering, so it is counted on the solver-order pattern
    int64_t nnz = 0;
    for (int r = 0; r < c->nVtot; ++r) {
        if (c->hFixed[c->hVertOf[r]]) { nnz += 2; continue; }
        int k = 0;
        for (int b = c->hSRowPtr[r]; b < c->hSRowPtr[r + 1]; ++b) if (c->hSColIdx[b] >= r) ++k;
        nnz += 4 * (int64_t)k - 1;
    }
    return nnz;
}

int ocb_get_sizes(const ocb_ctx* c, int64_t* s)
{
    if (!c || !s) return OCB_ERR_ARG;
    s[0] = c->nV; s[1] = c->nF; s[2] = c->nVa; s[3] = c->nFa; s[4] = c->nBnd; s[5] = c->nSys();
    s[6] = c->patternValid ? nnz_upper(c) : 0; s[7] = c->patternValid ? c->nnzb : 0;
    return OCB_OK;
}

// ------------------------------------------------------------------------------------------- energy
int ocb_energy(ocb_ctx* c, double p0, double* E_total, double* E_sd, double* E_scaf)
{
    HostTimer _ht("energy");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_energy: no UV"));
    OCB_TRY(launch_energy(c, p0, false, 0.0));
    OCB_TRY(fetch_scalars(c));
    const double esd = c->hScal[S_E_MESH], escaf = c->nFa > 0 ? c->wScafOverFa * c->hScal[S_E_AIR] : 0.0;
    if (E_sd) *E_sd = esd;
    if (E_scaf) *E_scaf = escaf;
    if (E_total) *E_total = p0 * esd + escaf;
    if (c->hScal[S_N_INVERTED] > 0.0) return set_err(c, OCB_ERR_INVERTED, "element with negative signed UV area");
    return OCB_OK;
}

int ocb_energy_per_elem(ocb_ctx* c, int uniform, double* out)
{
    if (!c || !out) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_energy_per_elem: no UV"));
    OCB_CUDA(c, c->scratchD.reserve((size_t)c->nF, c->stream));
    OCB_TRY(launch_energy_per_elem(c, uniform, c->scratchD.p));
    OCB_CUDA(c, cudaMemcpyAsync(out, c->scratchD.p, sizeof(double) * c->nF, cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

// a3: SymDirichletEnergy::getEnergyValByElemID (:48-68): the value of ONE triangle (used by the local-stencil code for its
// initial energy, TriMesh.cpp:2306, 2474, 2563)
int ocb_energy_by_elem(ocb_ctx* c, int triI, int uniform, double* E)
{
    if (!c || !E) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));
    OCB_TRY(need(c, c->haveUV, "ocb_energy_by_elem: no UV"));
    if (triI < 0 || triI >= c->nF) return set_err(c, OCB_ERR_ARG, "ocb_energy_by_elem: triangle index out of range");
    OCB_CUDA(c, c->scratchD.reserve(2, c->stream));
    OCB_TRY(launch_energy_one_elem(c, triI, uniform, c->scratchD.p));
    OCB_CUDA(c, cudaMemcpyAsync(E, c->scratchD.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

int ocb_gradient(ocb_ctx* c, double p0, double* g_out, double* sqnorm)
{
    HostTimer _ht("gradient");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_gradient: no UV"));
    OCB_TRY(launch_gradient(c, p0));
    if (g_out) OCB_TRY(vec_to_host(c, g_out, c->g.p));
    OCB_TRY(fetch_scalars(c));
    if (sqnorm) *sqnorm = c->hScal[S_SQN_G];
    // the fused pass also left the energy at x: ocb_newton_step_ex(OCB_STEP_REUSE_GRADIENT) continues from here
    c->gradValid = true; c->gradP0 = p0;
    c->gradSqn = c->hScal[S_SQN_G]; c->gradEMesh = c->hScal[S_E_MESH]; c->gradEAir = c->hScal[S_E_AIR]; c->gradSqnMesh = c->hScal[S_SQN_G_MESH];
    return OCB_OK;
}

// ------------------------------------------------------------------------------------------ pattern
// The solver's own row order: the Hilbert-curve order of the current UVs when they are known (plus the
// preconditioner hierarchy on top of it, see ocb_mas.cu), identity otherwise.
static int choose_order(ocb_ctx* c)
{
    static const bool masOff = []() { const char* e = getenv("OCB_NO_MAS"); return e && atoi(e); }();
    const int n = c->nVtot;
    c->planGrid = pcg_plan_grid(c, n);
    c->masH = MasHost();
    if (!masOff && c->haveUV && c->nV > 0 && c->x.p && n >= 2 * kMasLeaf) {
        const char* noMirror = getenv("OCB_NO_UV_MIRROR");      // test switch: always download
        const bool mirrored = c->hXY.size() == 2 * (size_t)n && c->hXYMesh && (c->hXYAir || c->nVa - c->nBnd <= 0) && !(noMirror && atoi(noMirror));
        std::vector<double> xy;
        if (!mirrored) {                                   // x moved on the device since the last ocb_set_uv
            xy.resize(2 * (size_t)n);
            OCB_CUDA(c, cudaMemcpyAsync(xy.data(), c->x.p, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
            OCB_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        { HostTimer _h2("  mas_build_hierarchy"); OCB_TRY(mas_build_hierarchy(c, mirrored ? c->hXY.data() : xy.data(), c->planGrid)); }
    } else if (!masOff && c->nV == 0 && (int)c->hCoords.size() == 2 * n && n >= 2 * kMasLeaf) {     // bare solver with a coordinate hint
        HostTimer _h2("  mas_build_hierarchy");
        OCB_TRY(mas_build_hierarchy(c, c->hCoords.data(), c->planGrid));
    } else {
        c->hRowOf.resize((size_t)n); c->hVertOf.resize((size_t)n);
        for (int v = 0; v < n; ++v) c->hRowOf[v] = c->hVertOf[v] = v;
    }
    return 0;
}

// hSRowPtr/hSColIdx (the pattern in solver order) are complete: upload, preconditioner tables, element slots
static int finish_install(ocb_ctx* c)
{
    const int n = c->nVtot;
    c->nnzb = (int)c->hSColIdx.size();
    OCB_CUDA(c, c->rowPtr.reserve((size_t)n + 1, c->stream));
    OCB_CUDA(c, c->colIdx.reserve((size_t)c->nnzb + 1, c->stream));
    OCB_CUDA(c, c->val.reserve(4 * (size_t)c->nnzb + 4, c->stream));
    OCB_CUDA(c, c->rowOf.reserve((size_t)n + 1, c->stream));
    OCB_CUDA(c, c->vertOf.reserve((size_t)n + 1, c->stream));
    OCB_CUDA(c, c->userRow.reserve((size_t)n + 1, c->stream));
    OCB_TRY(upload_i(c, c->rowPtr.p, c->hSRowPtr.data(), (size_t)n + 1));
    OCB_TRY(upload_i(c, c->colIdx.p, c->hSColIdx.data(), (size_t)c->nnzb));
    OCB_TRY(upload_i(c, c->rowOf.p, c->hRowOf.data(), (size_t)n));
    OCB_TRY(upload_i(c, c->vertOf.p, c->hVertOf.data(), (size_t)n));
    std::vector<int32_t> userRow((size_t)n);
    for (int u = 0; u < n; ++u) userRow[u] = c->hRowOf[c->hPerm[u]];
    OCB_TRY(upload_i(c, c->userRow.p, userRow.data(), (size_t)n));
    OCB_CUDA(c, c->fixedMask.reserve((size_t)c->nVtot, c->stream));
    c->hFixed.resize((size_t)c->nVtot, 0);
    OCB_CUDA(c, cudaMemcpyAsync(c->fixedMask.p, c->hFixed.data(), (size_t)c->nVtot, cudaMemcpyHostToDevice, c->stream));
    // (pageable sources: cudaMemcpyAsync returns once the data is staged, so the local vectors may go out of scope)
    { HostTimer _h3("  mas_install"); OCB_TRY(mas_install(c)); }
    c->patternValid = true; c->matrixValid = c->precondValid = false; ++c->matrixVersion;
    HostTimer _h4("  build_slots");
    return launch_build_slots(c);
}

// hRowPtr/hColIdx hold the pattern in INTERNAL vertex order (the adjacency route): permute it into the solver order
static int install_pattern(ocb_ctx* c)
{
    HostTimer _ht("install_pattern");
    const int n = c->nVtot;
    OCB_TRY(choose_order(c));
    const int nnzb = (int)c->hColIdx.size();
    c->hSRowPtr.assign((size_t)n + 1, 0);
    c->hSColIdx.resize((size_t)nnzb);
    int w = 0;
    for (int r = 0; r < n; ++r) {
        const int v = c->hVertOf[r];
        for (int b = c->hRowPtr[v]; b < c->hRowPtr[v + 1]; ++b) c->hSColIdx[w++] = c->hRowOf[c->hColIdx[b]];     // unsorted rows
        c->hSRowPtr[r + 1] = w;
    }
    c->hRowPtr.clear(); c->hColIdx.clear();
    return finish_install(c);
}

int ocb_set_pattern(ocb_ctx* c, int nVtot, const int32_t* adjPtr, const int32_t* adjIdx, const int32_t* fixed, int nFixed)
{
    HostTimer _ht("set_pattern(adjacency)");
    if (!c || !adjPtr || !adjIdx) return set_err(c, OCB_ERR_ARG, "ocb_set_pattern: bad argument");
    OCB_TRY(ensure_init(c));
    if (c->nV == 0) {     // solver-only use (a bare LinSysSolver): the system size comes from the adjacency
        if (nVtot <= 0) return set_err(c, OCB_ERR_ARG, "ocb_set_pattern: empty adjacency");
        c->nVtot = nVtot;
        OCB_CUDA(c, c->g.reserve(2 * (size_t)nVtot, c->stream));
        OCB_CUDA(c, c->p.reserve(2 * (size_t)nVtot, c->stream));
    }
    if (nVtot != c->nVtot) return set_err(c, OCB_ERR_ARG, "ocb_set_pattern: nVtot does not match mesh + air sizes");
    if (c->nV == 0) {                                     // solver-only use: locality order straight from the adjacency
        static const bool off = []() { const char* e = getenv("OCB_NO_REORDER"); return e && atoi(e); }();
        c->hPerm.assign((size_t)nVtot, 0);
        if (off) { for (int v = 0; v < nVtot; ++v) c->hPerm[v] = v; }
        else {
            std::vector<int32_t> order; order.reserve((size_t)nVtot);
            std::vector<uint8_t> seen((size_t)nVtot, 0);
            for (int s0 = 0; s0 < nVtot; ++s0) {
                if (seen[s0]) continue;
                seen[s0] = 1; order.push_back(s0);
                for (size_t head = order.size() - 1; head < order.size(); ++head) {
                    const int v = order[head];
                    for (int k = adjPtr[v]; k < adjPtr[v + 1]; ++k) {
                        const int u = adjIdx[k];
                        if (u >= 0 && u < nVtot && !seen[u]) { seen[u] = 1; order.push_back(u); }
                    }
                }
            }
            for (int i = 0; i < nVtot; ++i) c->hPerm[order[i]] = i;
        }
        c->nV = nVtot;                                    // upload_perm keeps ids >= nV fixed: none here
        const int rc = upload_perm(c);
        c->nV = 0;
        if (rc < 0) return rc;
    }
    c->hFixed.assign((size_t)nVtot, 0);
    for (int i = 0; i < nFixed; ++i) {
        if (fixed[i] < 0 || fixed[i] >= nVtot) return set_err(c, OCB_ERR_ARG, "fixed vertex out of range");
        c->hFixed[c->hPerm[fixed[i]]] = (c->nV > 0 && fixed[i] >= c->nV) ? 2 : 1;     // merged set (Scaffold::mergeFixedV): air-only ids follow the mesh's
    }
    c->hCoords.clear();
    if (c->nV == 0 && !c->hHint.empty()) {                 // bare solver: positions from the hint, the rest by neighbour means
        const int nKnown = std::min<int>((int)(c->hHint.size() / 2), nVtot);
        std::vector<double> xy(2 * (size_t)nVtot, 0.0);
        std::vector<uint8_t> known((size_t)nVtot, 0);
        for (int v = 0; v < nKnown; ++v) { xy[2 * (size_t)v] = c->hHint[2 * (size_t)v]; xy[2 * (size_t)v + 1] = c->hHint[2 * (size_t)v + 1]; known[v] = 1; }
        for (int sweep = 0, left = nVtot - nKnown; sweep < 32 && left > 0; ++sweep) {
            std::vector<int> newly;
            for (int v = nKnown; v < nVtot; ++v) {
                if (known[v]) continue;
                double sx = 0.0, sy = 0.0; int k = 0;
                for (int q = adjPtr[v]; q < adjPtr[v + 1]; ++q) {
                    const int u = adjIdx[q];
                    if (u >= 0 && u < nVtot && known[u]) { sx += xy[2 * (size_t)u]; sy += xy[2 * (size_t)u + 1]; ++k; }
                }
                if (k > 0) { xy[2 * (size_t)v] = sx / k; xy[2 * (size_t)v + 1] = sy / k; newly.push_back(v); }
            }
            for (int v : newly) known[v] = 1;
            left -= (int)newly.size();
            if (newly.empty()) break;
        }
        c->hCoords.resize(2 * (size_t)nVtot);
        for (int v = 0; v < nVtot; ++v) { c->hCoords[2 * (size_t)c->hPerm[v]] = xy[2 * (size_t)v]; c->hCoords[2 * (size_t)c->hPerm[v] + 1] = xy[2 * (size_t)v + 1]; }
    }
    c->hRowPtr.assign((size_t)nVtot + 1, 0);
    c->hColIdx.clear();
    c->hColIdx.reserve((size_t)adjPtr[nVtot] + nVtot);
    std::vector<int32_t> rowBuf;
    for (int vi = 0; vi < nVtot; ++vi) {                  // vi: internal row, v: the caller's vertex
        const int v = c->hInv[vi];
        if (c->hFixed[vi]) { c->hColIdx.push_back(vi); }
        else {
            rowBuf.clear();
            rowBuf.push_back(vi);
            for (int k = adjPtr[v]; k < adjPtr[v + 1]; ++k) {
                const int nb = adjIdx[k];
                if (nb < 0 || nb >= nVtot) return set_err(c, OCB_ERR_ARG, "adjacency index out of range");
                if (k > adjPtr[v] && adjIdx[k - 1] >= nb) return set_err(c, OCB_ERR_ARG, "adjacency rows must be strictly ascending");
                if (nb == v) continue;
                const int ni = c->hPerm[nb];
                if (!c->hFixed[ni]) rowBuf.push_back(ni);
            }
            std::sort(rowBuf.begin(), rowBuf.end());
            c->hColIdx.insert(c->hColIdx.end(), rowBuf.begin(), rowBuf.end());
        }
        c->hRowPtr[vi + 1] = (int32_t)c->hColIdx.size();
    }
    return install_pattern(c);
}

int ocb_set_pattern_from_elements(ocb_ctx* c)
{
    HostTimer _ht("set_pattern_from_elements");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(need(c, c->nV > 0, "ocb_set_pattern_from_elements before ocb_set_mesh"));
    OCB_TRY(ensure_init(c));
    const int nVtot = c->nVtot, nV = c->nV;
    c->hFixed.resize((size_t)nVtot, 0);
    HostTimer* _ta = new HostTimer("  adjacency");
    // The MESH part of the adjacency changes only with ocb_set_mesh: its de-duplicated neighbour lists are built once and kept
    // (f2, first half: with bijectivity on the pattern is rebuilt after every Newton iteration because the AIR mesh changed).
    if (!c->meshAdjValid) {
        std::vector<int32_t> cnt((size_t)nV + 1, 0);
        const std::vector<int32_t>& F = c->hF; const int n = c->nF;
        for (int t = 0; t < n; ++t) for (int k = 0; k < 3; ++k) cnt[(size_t)F[(size_t)k * n + t] + 1] += 2;
        for (int v = 0; v < nV; ++v) cnt[v + 1] += cnt[v];
        std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1), buf((size_t)cnt[nV]);
        for (int t = 0; t < n; ++t) {
            const int a = F[t], b = F[(size_t)n + t], d = F[2 * (size_t)n + t];
            buf[fill[a]++] = b; buf[fill[a]++] = d; buf[fill[b]++] = a; buf[fill[b]++] = d; buf[fill[d]++] = a; buf[fill[d]++] = b;
        }
        c->hMeshAdjPtr.assign((size_t)nV + 1, 0);
        c->hMeshAdj.clear(); c->hMeshAdj.reserve(buf.size() / 2 + 16);
        std::vector<int32_t> stamp((size_t)nV, -1);
        for (int v = 0; v < nV; ++v) {
            for (int q = cnt[v]; q < fill[v]; ++q) { const int u = buf[q]; if (stamp[u] != v) { stamp[u] = v; c->hMeshAdj.push_back(u); } }
            c->hMeshAdjPtr[v + 1] = (int32_t)c->hMeshAdj.size();
        }
        c->meshAdjValid = true;
    }
    // air mesh: buckets with duplicates (small), de-duplicated while the rows are written
    std::vector<int32_t> acnt((size_t)nVtot + 1, 0);
    {
        const std::vector<int32_t>& F = c->hFa; const int n = c->nFa;
        for (int t = 0; t < n; ++t) for (int k = 0; k < 3; ++k) acnt[(size_t)F[(size_t)k * n + t] + 1] += 2;
    }
    for (int v = 0; v < nVtot; ++v) acnt[v + 1] += acnt[v];
    std::vector<int32_t> afill(acnt.begin(), acnt.end() - 1), abuf((size_t)acnt[nVtot]);
    {
        const std::vector<int32_t>& F = c->hFa; const int n = c->nFa;
        for (int t = 0; t < n; ++t) {
            const int a = F[t], b = F[(size_t)n + t], d = F[2 * (size_t)n + t];
            abuf[afill[a]++] = b; abuf[afill[a]++] = d; abuf[afill[b]++] = a; abuf[afill[b]++] = d; abuf[afill[d]++] = a; abuf[afill[d]++] = b;
        }
    }
    delete _ta;
    // the solver's row order first (it needs only the UVs), then the pattern directly in that order
    { HostTimer _tb("  choose_order (incl. UV download)"); OCB_TRY(choose_order(c)); }
    HostTimer* _tc = new HostTimer("  pattern rows");
    c->hRowPtr.clear(); c->hColIdx.clear();
    c->hSRowPtr.resize((size_t)nVtot + 1);
    c->hSColIdx.resize(c->hMeshAdj.size() + abuf.size() + (size_t)nVtot);      // upper bound, trimmed below
    c->hStamp.assign((size_t)nVtot, -1);
    {
        int32_t* out = c->hSColIdx.data();
        int32_t* rowPtr = c->hSRowPtr.data();
        int32_t* stamp = c->hStamp.data();
        const int32_t* vertOf = c->hVertOf.data(); const int32_t* rowOf = c->hRowOf.data();
        const uint8_t* fixed = c->hFixed.data();
        const int32_t* mp = c->hMeshAdjPtr.data(); const int32_t* ma = c->hMeshAdj.data();
        int w = 0;
        rowPtr[0] = 0;
        for (int r = 0; r < nVtot; ++r) {
            const int v = vertOf[r];
            if (fixed[v]) out[w++] = r;
            else {
                out[w++] = r;                                  // the diagonal block; rows stay unsorted (every look-up is a linear scan)
                const bool hasAir = acnt[v] < afill[v];
                if (v < nV) for (int q = mp[v]; q < mp[v + 1]; ++q) { const int u = ma[q]; if (hasAir) stamp[u] = v; if (!fixed[u]) out[w++] = rowOf[u]; }
                if (hasAir) {
                    stamp[v] = v;
                    for (int q = acnt[v]; q < afill[v]; ++q) { const int u = abuf[q]; if (stamp[u] == v) continue; stamp[u] = v; if (!fixed[u]) out[w++] = rowOf[u]; }
                }
            }
            rowPtr[r + 1] = w;
        }
        c->hSColIdx.resize((size_t)w);
    }
    delete _tc;
    HostTimer _td("  finish_install");
    return finish_install(c);
}

// ------------------------------------------------------------------------------------------ Hessian
int ocb_hessian_assemble(ocb_ctx* c, double p0)
{
    HostTimer _ht("hessian_assemble");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_hessian_assemble: no UV"));
    if (!c->patternValid && c->nV > 0) OCB_TRY(ocb_set_pattern_from_elements(c));     // no pattern handed over: the element lists define it
    OCB_TRY(need(c, c->patternValid && c->slotsValid, "ocb_hessian_assemble: no sparsity pattern (ocb_set_pattern)"));
    OCB_TRY(launch_hessian(c, p0));
    ++c->matrixVersion;
    c->matrixValid = true; c->precondValid = false; c->systemScaled = false;
    return OCB_OK;
}

int ocb_hessian_blocks(ocb_ctx* c, int uniform, double* out)
{
    HostTimer _ht("hessian_blocks");
    if (!c || !out) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_hessian_blocks: no UV"));
    OCB_CUDA(c, c->scratchD.reserve(36 * (size_t)c->nF, c->stream));
    OCB_TRY(launch_hessian_blocks(c, uniform, c->scratchD.p));
    OCB_CUDA(c, cudaMemcpyAsync(out, c->scratchD.p, sizeof(double) * 36 * (size_t)c->nF, cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

// a6: Energy::computeHessian(data, MatrixXd&, uniformWeight) -- the dense flavour the nested optimizers use (SymDirichletEnergy.cpp:306-427)
int ocb_hessian_dense(ocb_ctx* c, int uniform, double* H_out)
{
    HostTimer _ht("hessian_dense");
    if (!c || !H_out) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));
    OCB_TRY(need(c, c->haveUV, "ocb_hessian_dense: no UV"));
    if (c->nV > 4096) return set_err(c, OCB_ERR_ARG, "ocb_hessian_dense: more than 4096 vertices (a dense 2n x 2n matrix is for local stencils)");
    const size_t n = 2 * (size_t)c->nV;
    OCB_CUDA(c, c->scratchD.reserve(36 * (size_t)c->nF + n * n + 4, c->stream));
    OCB_CUDA(c, c->scratchI.reserve((size_t)c->nV + 2, c->stream));
    double* dB = c->scratchD.p; double* dH = dB + 36 * (size_t)c->nF;
    OCB_TRY(upload_i(c, c->scratchI.p, c->hInv.data(), (size_t)c->nV));
    OCB_TRY(launch_hessian_blocks(c, uniform, dB));
    OCB_CUDA(c, cudaMemsetAsync(dH, 0, sizeof(double) * n * n, c->stream));
    OCB_TRY(launch_dense_hessian(c, dB, c->scratchI.p, dH));
    OCB_CUDA(c, cudaMemcpyAsync(H_out, dH, sizeof(double) * n * n, cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

int ocb_hessian_triplets(ocb_ctx* c, int uniform, double* V, int32_t* I, int32_t* J, int64_t* n)
{
    HostTimer _ht("hessian_triplets");
    if (!c || !n) return OCB_ERR_ARG;
    // count first: dim^2 * free^2 per triangle with >=1 free vertex, + 2 per fixed vertex
    int64_t cntT = 0;
    const int nF = c->nF;
    for (int t = 0; t < nF; ++t) {
        int nf = 0;
        for (int k = 0; k < 3; ++k) nf += c->hFixed[c->hF[(size_t)k * nF + t]] ? 0 : 1;     // hF: internal ids, like hFixed
        cntT += 4 * (int64_t)nf * nf;
    }
    int nFixedMesh = 0;
    for (int v = 0; v < c->nV; ++v) nFixedMesh += c->hFixed[c->hPerm[v]] ? 1 : 0;
    cntT += 2 * (int64_t)nFixedMesh;
    *n = cntT;
    if (!V) return OCB_OK;
    if (!I || !J) return set_err(c, OCB_ERR_ARG, "ocb_hessian_triplets: I/J missing");
    std::vector<double> blocks(36 * (size_t)nF);
    OCB_TRY(ocb_hessian_blocks(c, uniform, blocks.data()));
    int64_t w = 0;
    for (int t = 0; t < nF; ++t) {
        int idx[3];
        for (int k = 0; k < 3; ++k) { idx[k] = c->hFuser[(size_t)k * nF + t]; if (c->hFixed[c->hPerm[idx[k]]]) idx[k] = -1; }   // triplets carry the caller's ids
        const double* B = blocks.data() + 36 * (size_t)t;
        for (int a = 0; a < 3; ++a) {
            if (idx[a] < 0) continue;
            for (int b = 0; b < 3; ++b) {
                if (idx[b] < 0) continue;
                for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
                    V[w] = B[(2 * a + i) * 6 + 2 * b + j]; I[w] = 2 * idx[a] + i; J[w] = 2 * idx[b] + j; ++w;
                }
            }
        }
    }
    for (int v = 0; v < c->nV; ++v) if (c->hFixed[c->hPerm[v]]) for (int i = 0; i < 2; ++i) { V[w] = 1.0; I[w] = J[w] = 2 * v + i; ++w; }
    return OCB_OK;
}

int ocb_update_values_triplets(ocb_ctx* c, int64_t nT, const int32_t* I, const int32_t* J, const double* S)
{
    HostTimer _ht("update_values_triplets");
    if (!c || nT < 0 || (nT > 0 && (!I || !J || !S))) return set_err(c, OCB_ERR_ARG, "ocb_update_values_triplets: bad argument");
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->patternValid, "ocb_update_values_triplets: no sparsity pattern"));
    OCB_CUDA(c, c->scratchI.reserve(2 * (size_t)nT + 2, c->stream));
    OCB_CUDA(c, c->scratchD.reserve((size_t)nT + 1, c->stream));
    if (nT > 0) {
        OCB_TRY(upload_i(c, c->scratchI.p, I, (size_t)nT));
        OCB_TRY(upload_i(c, c->scratchI.p + nT, J, (size_t)nT));
        OCB_TRY(upload_d(c, c->scratchD.p, S, (size_t)nT));
    }
    OCB_TRY(launch_triplet_scatter(c, nT, c->scratchI.p, c->scratchI.p + nT, c->scratchD.p));
    ++c->matrixVersion;
    c->matrixValid = true; c->precondValid = false; c->systemScaled = false;
    return OCB_OK;
}

long long ocb_matrix_version(const ocb_ctx* c) { return c ? c->matrixVersion : -1; }

int ocb_download_csr(ocb_ctx* c, int32_t* ia, int32_t* ja, double* a)
{
    if (!c || !ia || !ja || !a) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->patternValid && c->matrixValid && !c->systemScaled, "ocb_download_csr: no matrix"));
    std::vector<double> val(4 * (size_t)c->nnzb);
    OCB_CUDA(c, cudaMemcpyAsync(val.data(), c->val.p, sizeof(double) * val.size(), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    int64_t w = 0;
    ia[0] = 1;
    std::vector<std::pair<int32_t, int32_t>> row;         // (caller column id, block slot)
    for (int v = 0; v < c->nVtot; ++v) {                  // v: the caller's vertex, vi its internal row
        const int vi = c->hPerm[v], r = c->hRowOf[vi];                      // r: the solver row of the device BSR
        if (c->hFixed[vi]) {
            int bdiag = -1;
            for (int b = c->hSRowPtr[r]; b < c->hSRowPtr[r + 1]; ++b) if (c->hSColIdx[b] == r) bdiag = b;
            ja[w] = 2 * v + 1; a[w] = val[4 * (size_t)bdiag]; ++w; ia[2 * v + 1] = (int32_t)(w + 1);
            ja[w] = 2 * v + 2; a[w] = val[4 * (size_t)bdiag + 3]; ++w; ia[2 * v + 2] = (int32_t)(w + 1);
            continue;
        }
        row.clear();
        for (int b = c->hSRowPtr[r]; b < c->hSRowPtr[r + 1]; ++b) {
            const int col = c->hInv[c->hVertOf[c->hSColIdx[b]]];
            if (col >= v) row.push_back(std::make_pair((int32_t)col, (int32_t)b));
        }
        std::sort(row.begin(), row.end());
        for (int r = 0; r < 2; ++r) {
            for (const auto& cb : row) {
                for (int q = 0; q < 2; ++q) {
                    if (cb.first == v && q < r) continue;     // strict lower entry of the diagonal block
                    ja[w] = 2 * cb.first + q + 1; a[w] = val[4 * (size_t)cb.second + 2 * r + q]; ++w;
                }
            }
            ia[2 * v + r + 1] = (int32_t)(w + 1);
        }
    }
    return OCB_OK;
}

int ocb_multiply(ocb_ctx* c, const double* x, double* y)
{
    if (!c || !x || !y) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->patternValid && c->matrixValid && !c->systemScaled, "ocb_multiply: no matrix"));
    const size_t n = c->nSys();
    OCB_CUDA(c, c->scratchD.reserve(3 * n, c->stream));
    OCB_TRY(vec_to_device(c, c->scratchD.p, x));
    OCB_TRY(launch_gather_rows(c, c->scratchD.p, c->scratchD.p + n, true));
    OCB_TRY(launch_spmv(c, c->scratchD.p + n, c->scratchD.p + 2 * n));
    OCB_TRY(launch_gather_rows(c, c->scratchD.p + 2 * n, c->scratchD.p, false));
    OCB_TRY(vec_to_host(c, y, c->scratchD.p));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

// -------------------------------------------------------------------------------------------- solve
int ocb_factorize(ocb_ctx* c)
{
    HostTimer _ht("factorize");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->patternValid && c->matrixValid, "ocb_factorize: no matrix"));
    OCB_TRY(launch_jacobi_setup(c, !c->deferFactorCheck));
    OCB_TRY(launch_mas_setup(c));
    c->precondValid = true;
    return OCB_OK;
}

// Experimental (OCB_DIRECT_FIRST=1, off by default): after a Newton system CG could not handle, the next 4, 8, ... 32 systems go to the
// direct safety net without trying CG first.  Measured (150 iterations, profiles/r2_direct_first_experiment.txt): male_2 9.2 -> 6.2 s,
// but torusOnPlane 4.2 -> 8.0 s and cat_noUV 12.0 -> 48 s: the hard and the healthy systems of such a run are interleaved, and a
// direct solve costs 20x a healthy CG solve.  CG first, every time, is the default.
static bool direct_first(ocb_ctx* c)
{
    static const bool on = []() { const char* e = getenv("OCB_DIRECT_FIRST"); return e && atoi(e); }();
    return on && c->tolerateIndefinite && c->directSkip > 0 && direct_solver_available(c);
}

int ocb_solve(ocb_ctx* c, const double* rhs, double* x_out, double rel_tol, int max_it, int* iters, double* rel_res)
{
    HostTimer _ht("solve");
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->patternValid && c->matrixValid, "ocb_solve: no matrix"));
    const size_t n = c->nSys();
    if (rel_tol <= 0.0) rel_tol = 1e-12;
    if (max_it <= 0) max_it = (int)std::min<size_t>(20 * n, 2000000);
    const double* dRhs = c->g.p;
    bool negate = true;
    if (rhs) {
        OCB_CUDA(c, c->pb.reserve(n, c->stream));
        OCB_TRY(vec_to_device(c, c->pb.p, rhs));
        dRhs = c->pb.p; negate = false;
    }
    // The safety net FIRST: option force_direct (tests), or a Newton iteration right after one whose system CG could not handle
    // (direct_first(): a hard phase lasts for dozens of iterations; the failed CG attempt and the preconditioner set-up of each
    // are skipped, and CG gets another try after 4, 8, ... 32 iterations)
    const bool directFirst = direct_first(c);
    if ((c->forceDirect || directFirst) && direct_solver_available(c)) {
        int lifts = 0;
        const int rd = launch_direct_solve(c, dRhs, negate, &lifts);
        if (rd < 0) return rd;
        if (rd == 0) {
            if (directFirst) c->directSkip--;
            c->hScal[S_PCG_ITERS] = 0.0; c->hScal[S_PCG_STATUS] = 0.0; c->hScal[S_PCG_RELRES] = 0.0;
            if (x_out) OCB_TRY(vec_to_host(c, x_out, c->p.p));
            if (x_out) OCB_CUDA(c, cudaStreamSynchronize(c->stream));
            if (iters) *iters = 0;
            if (rel_res) *rel_res = 0.0;
            return OCB_OK;
        }
    }
    if (!c->precondValid) OCB_TRY(ocb_factorize(c));
    OCB_TRY(launch_pcg(c, dRhs, negate, rel_tol, max_it));
    OCB_TRY(fetch_scalars(c));
    if (c->hScal[S_JACOBI_BAD] != 0.0 && !c->tolerateIndefinite) {       // verdict of a set-up whose host check was deferred
        c->precondValid = false;
        return set_err(c, OCB_ERR_BREAKDOWN, "a diagonal 2x2 block of the matrix is not positive definite");
    }
    int itersTotal = (int)c->hScal[S_PCG_ITERS];
    static const bool dbg = []() { const char* e = getenv("OCB_PCG_DEBUG"); return e && atoi(e); }();
    {
        // Safety net (ocb_direct.cu): a system CG cannot handle -- the two-level preconditioner came out indefinite (status 3), the
        // assembled matrix is indefinite by rounding (2), or, inside a Newton iteration, the iteration cap was reached (1) -- is
        // solved by a dense Cholesky when it is small enough.  Healthy systems never come here.
        const int st0 = (int)c->hScal[S_PCG_STATUS];
        if (st0 == 0 && c->tolerateIndefinite) c->directBackoff = 0;          // CG handled this Newton system: the hard phase is over
        if ((st0 == 3 || ((st0 == 2 || st0 == 1) && c->tolerateIndefinite)) && direct_solver_available(c)) {
            if (dbg) fprintf(stderr, "[ocb pcg] CG gave up (status %d after %d iterations, r.M^-1 r = %.3e): direct solve of %d unknowns\n",
                             st0, itersTotal, st0 == 3 ? c->hScal[S_MISC0] : 0.0, (int)n);
            int lifts = 0;
            const int rd = launch_direct_solve(c, dRhs, negate, &lifts);
            if (rd < 0) return rd;
            if (rd == 0) {
                if (c->tolerateIndefinite) { c->directBackoff = std::min(std::max(2 * c->directBackoff, 4), 32); c->directSkip = c->directBackoff; }
                c->hScal[S_PCG_STATUS] = 0.0; c->hScal[S_PCG_RELRES] = 0.0;
                c->lastDirectLifts = lifts;
                if (st0 == 3) c->precondFallbacks++;
            }
        }
    }
    if ((int)c->hScal[S_PCG_STATUS] == 3) {
        // the two-level preconditioner came out indefinite (r.M^-1 r <= 0, detected on the device) and the dense safety net is
        // not available: repeat the solve with block-Jacobi only, which cannot fail on an SPD matrix
        c->precondFallbacks++;
        if (dbg) fprintf(stderr, "[ocb pcg] two-level preconditioner rejected at CG iteration %d: r.M^-1 r = %.6e (previous %.6e); repeating with block-Jacobi\n",
                         (int)c->hScal[S_PCG_ITERS], c->hScal[S_MISC0], c->hScal[S_MISC1]);
        OCB_TRY(launch_pcg(c, dRhs, negate, rel_tol, max_it, false));
        OCB_TRY(fetch_scalars(c));
        itersTotal += (int)c->hScal[S_PCG_ITERS];
    }
    if (x_out) OCB_TRY(vec_to_host(c, x_out, c->p.p));
    if (x_out) OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (iters) *iters = itersTotal;
    if (rel_res) *rel_res = c->hScal[S_PCG_RELRES];
    const int st = (int)c->hScal[S_PCG_STATUS];
    if (st == 2) return set_err(c, OCB_ERR_BREAKDOWN, "PCG breakdown: d^T A d <= 0 (matrix not SPD)");
    if (st == 1) return set_err(c, OCB_ERR_NOT_CONVERGED, "PCG reached max_it");
    if (st == 3) return set_err(c, OCB_ERR_BREAKDOWN, "PCG breakdown: preconditioner not positive definite");
    return OCB_OK;
}

int ocb_direct_level_blocks(int n, const int32_t* rowPtr, const int32_t* colIdx, int target, int32_t* pos, int32_t* blkOf, int32_t* blkBeg)
{
    if (n <= 0 || !rowPtr || !colIdx || !pos || !blkOf || !blkBeg || target < 1) return OCB_ERR_ARG;
    for (int v = 0; v < n; ++v) if (rowPtr[v + 1] < rowPtr[v]) return OCB_ERR_ARG;
    for (int q = 0; q < rowPtr[n]; ++q) if (colIdx[q] < 0 || colIdx[q] >= n) return OCB_ERR_ARG;
    return direct_level_blocks_host(n, rowPtr, colIdx, target, pos, blkOf, blkBeg);
}

int ocb_get_search_dir(ocb_ctx* c, double* p)
{
    if (!c || !p) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(vec_to_host(c, p, c->p.p));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}
int ocb_set_search_dir(ocb_ctx* c, const double* p)
{
    if (!c || !p) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->nVtot > 0, "ocb_set_search_dir before the system size is known"));
    OCB_TRY(vec_to_device(c, c->p.p, p));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

static void fill_precond_info(const MasHost& H, int32_t* info)
{
    for (int i = 0; i < 16; ++i) info[i] = 0;
    info[0] = H.enabled ? 1 : 0; info[1] = H.L; info[2] = H.Lloc; info[3] = H.grid;
    for (int l = 1; l <= H.L && l <= 12; ++l) info[3 + l] = (int32_t)H.lv[l - 1].childBeg.size() - 1;
}
int ocb_precond_info(const ocb_ctx* c, int32_t* info)
{
    if (!c || !info) return OCB_ERR_ARG;
    fill_precond_info(c->masH, info);
    info[15] = (int32_t)c->precondFallbacks;
    info[14] = (int32_t)c->directSolves;
    return OCB_OK;
}
int ocb_set_coordinate_hint(ocb_ctx* c, int n, const double* xy)
{
    if (!c || n < 0 || (n > 0 && !xy)) return set_err(c, OCB_ERR_ARG, "ocb_set_coordinate_hint: bad argument");
    c->hHint.resize(2 * (size_t)n);
    for (int v = 0; v < n; ++v) { c->hHint[2 * (size_t)v] = xy[v]; c->hHint[2 * (size_t)v + 1] = xy[(size_t)n + v]; }
    return OCB_OK;
}
int ocb_precond_hierarchy(ocb_ctx* c, int n, const double* xy, int grid, int32_t* vert_of, int32_t* info, int32_t* child_beg, int cap)
{
    if (!c || n <= 0 || !xy || !vert_of || !info || !child_beg) return set_err(c, OCB_ERR_ARG, "ocb_precond_hierarchy: bad argument");
    if (c->inited) return set_err(c, OCB_ERR_STATE, "ocb_precond_hierarchy needs a fresh context (it overwrites the system size)");
    c->nVtot = n;
    c->hFixed.assign((size_t)n, 0);
    OCB_TRY(mas_build_hierarchy(c, xy, grid));
    if (getenv("OCB_HIERARCHY_TWICE")) OCB_TRY(mas_build_hierarchy(c, xy, grid));     // timing aid: warm start from the previous order
    fill_precond_info(c->masH, info);
    std::memcpy(vert_of, c->hVertOf.data(), sizeof(int32_t) * (size_t)n);
    int w = 0;
    for (int l = 1; l <= c->masH.L; ++l) {
        const std::vector<int32_t>& cb = c->masH.lv[l - 1].childBeg;
        if (w + (int)cb.size() > cap) return set_err(c, OCB_ERR_ARG, "ocb_precond_hierarchy: child_beg too small");
        std::memcpy(child_beg + w, cb.data(), sizeof(int32_t) * cb.size());
        w += (int)cb.size();
    }
    c->nVtot = 0; c->masH = MasHost(); c->hFixed.clear();
    return w;
}

// ------------------------------------------------------------------------------- step bound / search
int ocb_step_bound(ocb_ctx* c, const double* dir, double* alpha)
{
    HostTimer _ht("step_bound");
    if (!c || !alpha) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_step_bound: no UV"));
    const double* dDir = c->p.p;
    if (dir) {
        OCB_CUDA(c, c->scratchD.reserve((size_t)c->nSys(), c->stream));
        OCB_TRY(vec_to_device(c, c->scratchD.p, dir));
        dDir = c->scratchD.p;
    }
    OCB_TRY(launch_step_bound(c, dDir, *alpha));
    OCB_TRY(fetch_scalars(c));
    *alpha = c->hScal[S_STEP_BOUND];
    return OCB_OK;
}

int ocb_step_forward(ocb_ctx* c, double alpha)
{
    if (!c) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_step_forward: no UV"));
    OCB_CUDA(c, cudaMemcpyAsync(c->x0.p, c->x.p, sizeof(double) * c->nSys(), cudaMemcpyDeviceToDevice, c->stream));
    OCB_TRY(launch_step_forward(c, alpha));
    c->hXYMesh = c->hXYAir = false;
    c->matrixValid = c->precondValid = false; c->gradValid = false;
    return OCB_OK;
}

// Line search proper (Optimizer.cpp:575-673).  x0 already holds the start point.  firstTrialFetched: the energy at
// x0 + alpha p has been launched AND fetched by the caller (ocb_newton_step chains it behind the step bound on the device).
static int line_search_core(ocb_ctx* c, double p0, double E_last, double lastScaf, double alpha, bool firstTrialFetched,
                            int allowEDecRelTol, ocb_linesearch_result* out)
{
    const bool scaf = c->nFa > 0;
    double E = 0.0, Escaf = 0.0, Esd = 0.0;
    int halvings = 0, stopped = 0;
    for (;;) {
        if (!firstTrialFetched) {
            OCB_TRY(launch_energy(c, p0, true, alpha));
            OCB_TRY(fetch_scalars(c));
        }
        firstTrialFetched = false;
        Esd = c->hScal[S_E_MESH];
        Escaf = scaf ? c->wScafOverFa * c->hScal[S_E_AIR] : 0.0;
        E = p0 * Esd + Escaf;
        // plain decrease test (:597); the inversion guard (:615-629) follows the loop
        if (!(E <= E_last)) {                  // also catches a non-finite trial energy
            alpha /= 2.0; ++halvings;
            if (alpha == 0.0) { stopped = 1; break; }
            continue;
        }
        break;
    }
    // inversion guard: the reference re-halves while any signed area is negative
    while (!stopped) {
        // S_N_INVERTED counts signed areas < 0, the test of TriMesh::checkInversion (TriMesh.cpp:1710-1734)
        if (!(c->hScal[S_N_INVERTED] > 0.0)) break;
        alpha /= 2.0; ++halvings;
        if (alpha == 0.0) { stopped = 1; break; }
        OCB_TRY(launch_energy(c, p0, true, alpha));
        OCB_TRY(fetch_scalars(c));
        Esd = c->hScal[S_E_MESH];
        Escaf = scaf ? c->wScafOverFa * c->hScal[S_E_AIR] : 0.0;
        E = p0 * Esd + Escaf;
    }
    OCB_TRY(launch_step_forward(c, alpha));
    c->hXYMesh = c->hXYAir = false;
    double eDec = E_last - E;
    if (scaf) eDec += (-lastScaf + Escaf);
    if (allowEDecRelTol && (eDec / E_last < 1.0e-6 * alpha) && (alpha > 1.0e-3)) stopped = 1;
    out->alpha = alpha; out->E_new = E; out->E_scaf_new = Escaf; out->E_sd_new = Esd; out->E_last = E_last;
    out->lastEDec = eDec; out->n_halvings = halvings; out->stopped = stopped;
    c->matrixValid = c->precondValid = false; c->gradValid = false;
    return OCB_OK;
}

int ocb_line_search(ocb_ctx* c, double p0, double E_last, double alpha0, int allowEDecRelTol, ocb_linesearch_result* out)
{
    HostTimer _ht("line_search");
    if (!c || !out) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_line_search: no UV"));
    const bool scaf = c->nFa > 0;
    double lastScaf = 0.0;
    OCB_CUDA(c, cudaMemcpyAsync(c->x0.p, c->x.p, sizeof(double) * c->nSys(), cudaMemcpyDeviceToDevice, c->stream));
    // the scaffold changed since E_last was computed (Optimizer.cpp:588-592); E_last <= 0 asks for a fresh
    // evaluation (the Symmetric Dirichlet energy is >= 4 * energyParam0 > 0)
    if (scaf || !(E_last > 0.0)) {
        OCB_TRY(launch_energy(c, p0, false, 0.0));
        OCB_TRY(fetch_scalars(c));
        lastScaf = scaf ? c->wScafOverFa * c->hScal[S_E_AIR] : 0.0;
        E_last = p0 * c->hScal[S_E_MESH] + lastScaf;
    }
    return line_search_core(c, p0, E_last, lastScaf, alpha0, false, allowEDecRelTol, out);
}

// One Newton iteration (Optimizer::solve(1), Optimizer.cpp:203-261, 505-673) with three host round trips instead of the
// six of the call-by-call sequence: {gradient norm, energy at x} | {block-Jacobi verdict, PCG status} | {step bound,
// first line-search trial}.  Same kernels, same arithmetic, same results as ocb_gradient + ocb_hessian_assemble +
// ocb_factorize + ocb_solve + ocb_step_bound + ocb_line_search.
int ocb_newton_step_ex(ocb_ctx* c, double p0, double targetGRes, double pcg_rel_tol, int pcg_max_it,
                       int allowEDecRelTol, int flags, ocb_newton_result* out)
{
    HostTimer _ht("newton_step");
    if (!c || !out) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));
    std::memset(out, 0, sizeof(*out));
    out->targetGRes = targetGRes;
    OCB_TRY(need(c, c->haveUV, "ocb_newton_step: no UV"));
    const bool scaf = c->nFa > 0;
    const double tA = wall_now();
    if (!((flags & OCB_STEP_REUSE_GRADIENT) && c->gradValid && c->gradP0 == p0)) {
        OCB_TRY(launch_gradient(c, p0));                     // also E_last with the current scaffold (Optimizer.cpp:588-592): one fused pass
        OCB_TRY(fetch_scalars(c));
        c->gradValid = true; c->gradP0 = p0;
        c->gradSqn = c->hScal[S_SQN_G]; c->gradEMesh = c->hScal[S_E_MESH]; c->gradEAir = c->hScal[S_E_AIR];
    }
    const double sqn = c->gradSqn;
    const double lastScaf = scaf ? c->wScafOverFa * c->gradEAir : 0.0;
    const double E_last = p0 * c->gradEMesh + lastScaf;
    out->sqn_g = sqn;
    if (!(flags & OCB_STEP_SKIP_CONVERGENCE_TEST) && sqn < targetGRes) { out->converged = 1; return OCB_OK; }
    if (!c->patternValid) OCB_TRY(ocb_set_pattern_from_elements(c));
    if (!((flags & OCB_STEP_REUSE_MATRIX) && c->matrixValid)) OCB_TRY(ocb_hessian_assemble(c, p0));
    if (c->scaleSystem && !c->systemScaled) { OCB_TRY(launch_scale_system(c)); c->precondValid = false; ++c->matrixVersion; }
    c->deferFactorCheck = true;
    c->tolerateIndefinite = true;                              // (see below; set here because direct_first() looks at it)
    const int rf = (c->precondValid || direct_first(c)) ? 0 : ocb_factorize(c);     // (the set-up is skipped when this solve goes to the direct safety net first)
    c->deferFactorCheck = false;
    if (rf < 0) { c->tolerateIndefinite = false; return rf; }
    int its = 0; double rr = 0.0;
    // Inside a Newton iteration a matrix that is SPD only up to rounding is not an error: non-PD diagonal blocks fall back to
    // the identity in the preconditioner, a CG breakdown returns the truncated iterate (see pcg_kernel), and the line search
    // decides -- what the reference gets from an LDL^T that never checks definiteness (EigenLibSolver.cpp:80-107).
    c->tolerateIndefinite = true;
    // iteration cap of the inexact-Newton safety net.  A healthy system needs ~0.6 sqrt(n) CG iterations (126 / 286 / 600 at 10k /
    // 160k / 1M faces), but ONE nearly degenerate triangle can raise that 30-fold on an otherwise ordinary state (bimba configs[1],
    // iteration 10: diagonal 3e-3 .. 1e10, 4 813 iterations to 1e-12 -- and a direction truncated at 2 000 moved the step bound by
    // 22 %, profiles/r2_direction_accuracy.txt), so the cap sits well above that; a system that has not converged by then (kappa
    // beyond 1e16 at an extremely distorted start) hands its current iterate -- a descent direction -- to the line search.
    if (pcg_max_it <= 0) pcg_max_it = std::max(10000, (int)(40.0 * std::sqrt((double)c->nSys())));
    int rs = ocb_solve(c, nullptr, nullptr, pcg_rel_tol, pcg_max_it, &its, &rr);
    // A breakdown means the ASSEMBLED matrix is indefinite: at an extremely distorted start (benchmark meshes male_2, cat_noUV,
    // horse ...: ||g||^2 up to 1e51) the rounding of entries of size 1e50 exceeds the soft eigenvalues by 30 orders of magnitude.
    // The reference's LDL^T never notices; CG does.  Only then the diagonal is lifted RELATIVELY, A_ii *= (1 + delta) with delta
    // escalating 1e-8 -> 1e-5 -> 1e-2, which covers rounding noise proportional to each row's own size and leaves soft rows alone,
    // and the solve is repeated.  Healthy systems never come here, so parity with the reference is untouched.
    for (int esc = 0; esc < 3 && rs == OCB_ERR_BREAKDOWN; ++esc) {
        static const double deltas[3] = {1.0e-8, 1.0e-5, 1.0e-2};
        OCB_TRY(launch_diag_shift(c, deltas[esc]));
        ++c->matrixVersion;
        c->precondValid = false;
        c->deferFactorCheck = true;
        const int rf2 = ocb_factorize(c);
        c->deferFactorCheck = false;
        if (rf2 < 0) return rf2;
        int its2 = 0;
        rs = ocb_solve(c, nullptr, nullptr, pcg_rel_tol, pcg_max_it, &its2, &rr);
        its += its2;
        out->reserved = esc + 1;                                 // how many times the diagonal was lifted
    }
    c->tolerateIndefinite = false;
    out->pcg_iters = its; out->pcg_rel_res = rr;
    out->pcg_status = rs;
    if (rs < 0 && rs != OCB_ERR_NOT_CONVERGED && rs != OCB_ERR_BREAKDOWN) return rs;
    const double tB = wall_now();
    // step bound, then the first trial at 0.99 * bound (Optimizer.cpp:580) chained on the device
    OCB_CUDA(c, cudaMemcpyAsync(c->x0.p, c->x.p, sizeof(double) * c->nSys(), cudaMemcpyDeviceToDevice, c->stream));
    OCB_TRY(launch_step_bound(c, c->p.p, 1.0));
    OCB_TRY(launch_energy(c, p0, true, 0.99, true));
    OCB_TRY(fetch_scalars(c));
    double alpha = c->hScal[S_STEP_BOUND];
    alpha *= 0.99;
    out->alpha_init = alpha;
    ocb_linesearch_result ls;
    OCB_TRY(line_search_core(c, p0, E_last, lastScaf, alpha, true, allowEDecRelTol, &ls));
    out->alpha = ls.alpha; out->E_new = ls.E_new; out->E_scaf_new = ls.E_scaf_new; out->E_sd_new = ls.E_sd_new;
    out->lastEDec = ls.lastEDec; out->n_halvings = ls.n_halvings; out->stopped = ls.stopped;
    out->E_last = E_last;
    out->ms_solve = 1e3 * (tB - tA); out->ms_line_search = 1e3 * (wall_now() - tB);
    return (rs == OCB_ERR_NOT_CONVERGED || rs == OCB_ERR_BREAKDOWN) ? OCB_ERR_NOT_CONVERGED : OCB_OK;
}

int ocb_newton_step(ocb_ctx* c, double p0, double targetGRes, double pcg_rel_tol, int pcg_max_it,
                    int allowEDecRelTol, ocb_newton_result* out)
{
    return ocb_newton_step_ex(c, p0, targetGRes, pcg_rel_tol, pcg_max_it, allowEDecRelTol, 0, out);
}

int ocb_gradient_info(ocb_ctx* c, double* info4)
{
    if (!c || !info4) return OCB_ERR_ARG;
    OCB_TRY(need(c, c->gradValid, "ocb_gradient_info: no gradient at the current x"));
    info4[0] = c->gradSqn; info4[1] = c->gradSqnMesh; info4[2] = c->gradEMesh; info4[3] = c->nFa > 0 ? c->wScafOverFa * c->gradEAir : 0.0;
    return OCB_OK;
}

// ------------------------------------------------------------------------------------------- seam
int ocb_seam_energy(ocb_ctx* c, int nCoh, const int32_t* cohE, const double* edgeLen, const int32_t* boundaryEdge,
                    double initSeamLen, double virtualRadius, double avgEdgeLen, int triSoup, double* E_se)
{
    if (!c || !E_se || nCoh < 0) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_seam_energy: no UV"));
    double sum = 0.0;
    if (nCoh > 0) {
        if (!cohE || !edgeLen || !boundaryEdge) return set_err(c, OCB_ERR_ARG, "ocb_seam_energy: bad argument");
        // boundary rows hold -1 indices: clamp for the gather, they are skipped by the flag
        std::vector<int32_t> coh(cohE, cohE + 4 * (size_t)nCoh);
        for (auto& v : coh) v = v < 0 ? 0 : c->hPerm[v];
        OCB_CUDA(c, c->scratchI.reserve(5 * (size_t)nCoh, c->stream));
        OCB_CUDA(c, c->scratchD.reserve((size_t)nCoh, c->stream));
        OCB_TRY(upload_i(c, c->scratchI.p, coh.data(), 4 * (size_t)nCoh));
        OCB_TRY(upload_i(c, c->scratchI.p + 4 * (size_t)nCoh, boundaryEdge, (size_t)nCoh));
        OCB_TRY(upload_d(c, c->scratchD.p, edgeLen, (size_t)nCoh));
        OCB_TRY(launch_seam(c, nCoh, c->scratchI.p, c->scratchD.p, c->scratchI.p + 4 * (size_t)nCoh, avgEdgeLen, triSoup));
        OCB_TRY(fetch_scalars(c));
        sum = c->hScal[S_MISC0];
    }
    *E_se = (sum + initSeamLen) / virtualRadius;
    return OCB_OK;
}

int ocb_divgrad_scores(ocb_ctx* c, double* out)
{
    if (!c || !out) return OCB_ERR_ARG;
    OCB_TRY(ensure_init(c));                             // selects the context's device (two contexts on two GPUs in one thread)
    OCB_TRY(need(c, c->haveUV, "ocb_divgrad_scores: no UV"));
    OCB_CUDA(c, c->pb.reserve(2 * (size_t)c->nV, c->stream));
    OCB_TRY(launch_divgrad(c, c->pb.p));
    OCB_TRY(launch_permute_scalar(c, c->nV, c->pb.p, c->pb.p + c->nV));
    OCB_CUDA(c, cudaMemcpyAsync(out, c->pb.p + c->nV, sizeof(double) * c->nV, cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    return OCB_OK;
}

int ocb_eval_stencils(ocb_ctx* c, const ocb_stencil_batch* B, int maxIter, double relGL2Tol, double* E_init, double* E_final,
                      double* UV_out, int32_t* iters, double* score, int32_t* status, int* argmax)
{
    if (!c || !B || B->nStencil < 0) return set_err(c, OCB_ERR_ARG, "ocb_eval_stencils: bad argument");
    if (argmax) *argmax = -1;
    if (B->nStencil == 0) return OCB_OK;
    if (!B->vert_ptr || !B->tri_ptr || !B->V_rest || !B->UV || !B->F || !B->is_free) return set_err(c, OCB_ERR_ARG, "ocb_eval_stencils: bad argument");
    OCB_TRY(ensure_init(c));
    const int nS = B->nStencil;
    const size_t nV = (size_t)B->vert_ptr[nS], nT = (size_t)B->tri_ptr[nS];
    for (int s = 0; s < nS; ++s) if (B->vert_ptr[s + 1] < B->vert_ptr[s] || B->tri_ptr[s + 1] < B->tri_ptr[s]) return set_err(c, OCB_ERR_ARG, "ocb_eval_stencils: ranges must be ascending");
    for (int s = 0; s < nS; ++s) {
        const int nv = B->vert_ptr[s + 1] - B->vert_ptr[s];
        for (int t = B->tri_ptr[s]; t < B->tri_ptr[s + 1]; ++t) for (int k = 0; k < 3; ++k)
            if (B->F[3 * (size_t)t + k] < 0 || B->F[3 * (size_t)t + k] >= nv) return set_err(c, OCB_ERR_ARG, "ocb_eval_stencils: local vertex index out of range");
    }
    // one double arena + one int arena on the device
    const size_t dIn = 3 * nV + 2 * nV + 2 * (size_t)nS, dOut = 2 * nV + 3 * (size_t)nS;
    OCB_CUDA(c, c->stD.reserve(dIn + dOut + 8, c->stream));
    const size_t iN = 2 * ((size_t)nS + 1) + 3 * nT + 2 * (size_t)nS + 2 + (nV + 3) / 4 + 4;
    OCB_CUDA(c, c->stI.reserve(iN, c->stream));
    double* dVr = c->stD.p; double* dUV = dVr + 3 * nV; double* dSc = dUV + 2 * nV; double* dOf = dSc + nS;
    double* dUVo = dOf + nS; double* dE0 = dUVo + 2 * nV; double* dE1 = dE0 + nS; double* dScore = dE1 + nS;
    int32_t* dVp = c->stI.p; int32_t* dTp = dVp + nS + 1; int32_t* dF = dTp + nS + 1; int32_t* dIt = dF + 3 * nT; int32_t* dSt = dIt + nS;
    int32_t* dArg = dSt + nS; uint8_t* dFree = reinterpret_cast<uint8_t*>(dArg + 2);
    OCB_TRY(upload_d(c, dVr, B->V_rest, 3 * nV)); OCB_TRY(upload_d(c, dUV, B->UV, 2 * nV));
    if (B->score_scale) OCB_TRY(upload_d(c, dSc, B->score_scale, (size_t)nS));
    if (B->score_offset) OCB_TRY(upload_d(c, dOf, B->score_offset, (size_t)nS));
    OCB_TRY(upload_i(c, dVp, B->vert_ptr, (size_t)nS + 1)); OCB_TRY(upload_i(c, dTp, B->tri_ptr, (size_t)nS + 1));
    OCB_TRY(upload_i(c, dF, B->F, 3 * nT));
    OCB_CUDA(c, cudaMemcpyAsync(dFree, B->is_free, nV, cudaMemcpyHostToDevice, c->stream));
    StencilHost h;
    h.nStencil = nS; h.vertPtr = dVp; h.triPtr = dTp; h.Vrest = dVr; h.UV = dUV; h.F = dF; h.isFree = dFree;
    h.scoreScale = B->score_scale ? dSc : nullptr; h.scoreOffset = B->score_offset ? dOf : nullptr;
    h.maxIter = maxIter > 0 ? maxIter : 100; h.relGL2Tol = relGL2Tol > 0.0 ? relGL2Tol : 1.0e-6;
    h.Einit = dE0; h.Efinal = dE1; h.UVout = dUVo; h.iters = dIt; h.score = dScore; h.status = dSt; h.argmax = dArg;
    OCB_TRY(launch_stencils(c, h));
    if (E_init) OCB_CUDA(c, cudaMemcpyAsync(E_init, dE0, sizeof(double) * nS, cudaMemcpyDeviceToHost, c->stream));
    if (E_final) OCB_CUDA(c, cudaMemcpyAsync(E_final, dE1, sizeof(double) * nS, cudaMemcpyDeviceToHost, c->stream));
    if (UV_out) OCB_CUDA(c, cudaMemcpyAsync(UV_out, dUVo, sizeof(double) * 2 * nV, cudaMemcpyDeviceToHost, c->stream));
    if (iters) OCB_CUDA(c, cudaMemcpyAsync(iters, dIt, sizeof(int32_t) * nS, cudaMemcpyDeviceToHost, c->stream));
    if (score) OCB_CUDA(c, cudaMemcpyAsync(score, dScore, sizeof(double) * nS, cudaMemcpyDeviceToHost, c->stream));
    if (status) OCB_CUDA(c, cudaMemcpyAsync(status, dSt, sizeof(int32_t) * nS, cudaMemcpyDeviceToHost, c->stream));
    int hArg = -1;
    OCB_CUDA(c, cudaMemcpyAsync(&hArg, dArg, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (argmax) *argmax = hArg;
    return OCB_OK;
}

int ocb_stencil_newton_step(ocb_ctx* c, const ocb_stencil_step_batch* B, double* UV_out, double* out6, int32_t* result)
{
    if (!c || !B || B->nStencil < 0) return set_err(c, OCB_ERR_ARG, "ocb_stencil_newton_step: bad argument");
    if (B->nStencil == 0) return OCB_OK;
    if (!B->vert_ptr || !B->tri_ptr || !B->n_mesh_vert || !B->n_mesh_tri || !B->V_rest || !B->UV || !B->F || !B->is_free || !B->area_thres || !B->target_gres ||
        !UV_out || !out6 || !result) return set_err(c, OCB_ERR_ARG, "ocb_stencil_newton_step: bad argument");
    OCB_TRY(ensure_init(c));
    const int nS = B->nStencil;
    for (int s = 0; s < nS; ++s)
        if (B->vert_ptr[s + 1] < B->vert_ptr[s] || B->tri_ptr[s + 1] < B->tri_ptr[s] || B->n_mesh_tri[s] > B->tri_ptr[s + 1] - B->tri_ptr[s] ||
            B->n_mesh_vert[s] > B->vert_ptr[s + 1] - B->vert_ptr[s]) return set_err(c, OCB_ERR_ARG, "ocb_stencil_newton_step: inconsistent ranges");
    const size_t nV = (size_t)B->vert_ptr[nS], nT = (size_t)B->tri_ptr[nS];
    // ONE pinned staging arena and ONE device arena: the whole batch goes up in a single copy and comes back in a single copy
    // (a lock-step round is ~40 us of kernel; ten pageable copies of a few KB each cost more than that)
    auto al = [](size_t bytes) { return (bytes + 15) / 16 * 16; };
    const size_t oVr = 0, oUV = oVr + al(24 * nV), oTh = oUV + al(16 * nV), oTg = oTh + al(8 * (size_t)nS);
    const size_t oVp = oTg + al(8 * (size_t)nS), oTp = oVp + al(4 * ((size_t)nS + 1)), oNv = oTp + al(4 * ((size_t)nS + 1)), oNt = oNv + al(4 * (size_t)nS);
    const size_t oF = oNt + al(4 * (size_t)nS), oFr = oF + al(12 * nT), inBytes = oFr + al(nV);
    const size_t oUVo = inBytes, oOut = oUVo + al(16 * nV), oRes = oOut + al(48 * (size_t)nS), total = oRes + al(4 * (size_t)nS);
    if (c->stPinnedCap < total) {
        if (c->stPinned) cudaFreeHost(c->stPinned);
        c->stPinned = nullptr; c->stPinnedCap = 0;
        OCB_CUDA(c, cudaMallocHost((void**)&c->stPinned, total + total / 2 + 4096));
        c->stPinnedCap = total + total / 2 + 4096;
    }
    OCB_CUDA(c, c->stD.reserve(total / 8 + 8, c->stream));
    unsigned char* hp = c->stPinned; unsigned char* dp = reinterpret_cast<unsigned char*>(c->stD.p);
    std::memcpy(hp + oVr, B->V_rest, 24 * nV); std::memcpy(hp + oUV, B->UV, 16 * nV);
    std::memcpy(hp + oTh, B->area_thres, 8 * (size_t)nS); std::memcpy(hp + oTg, B->target_gres, 8 * (size_t)nS);
    std::memcpy(hp + oVp, B->vert_ptr, 4 * ((size_t)nS + 1)); std::memcpy(hp + oTp, B->tri_ptr, 4 * ((size_t)nS + 1));
    std::memcpy(hp + oNv, B->n_mesh_vert, 4 * (size_t)nS); std::memcpy(hp + oNt, B->n_mesh_tri, 4 * (size_t)nS);
    std::memcpy(hp + oF, B->F, 12 * nT); std::memcpy(hp + oFr, B->is_free, nV);
    OCB_CUDA(c, cudaMemcpyAsync(dp, hp, inBytes, cudaMemcpyHostToDevice, c->stream));
    StencilStepHost h;
    h.nStencil = nS;
    h.vertPtr = reinterpret_cast<int32_t*>(dp + oVp); h.triPtr = reinterpret_cast<int32_t*>(dp + oTp);
    h.nVm = reinterpret_cast<int32_t*>(dp + oNv); h.nTm = reinterpret_cast<int32_t*>(dp + oNt);
    h.Vrest = reinterpret_cast<double*>(dp + oVr); h.UV = reinterpret_cast<double*>(dp + oUV); h.F = reinterpret_cast<int32_t*>(dp + oF);
    h.isFree = dp + oFr; h.areaThres = reinterpret_cast<double*>(dp + oTh); h.targetGRes = reinterpret_cast<double*>(dp + oTg); h.wScaf = B->w_scaf;
    h.UVout = reinterpret_cast<double*>(dp + oUVo); h.out6 = reinterpret_cast<double*>(dp + oOut); h.result = reinterpret_cast<int32_t*>(dp + oRes);
    OCB_TRY(launch_stencil_step(c, h));
    OCB_CUDA(c, cudaMemcpyAsync(hp + oUVo, dp + oUVo, total - oUVo, cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    std::memcpy(UV_out, hp + oUVo, 16 * nV); std::memcpy(out6, hp + oOut, 48 * (size_t)nS); std::memcpy(result, hp + oRes, 4 * (size_t)nS);
    return OCB_OK;
}

}  // extern "C"
