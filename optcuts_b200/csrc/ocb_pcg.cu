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
#include "ocb_mas.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>

namespace ocb {

static constexpr int kPcgBlock = 1024;
// the one-cluster mode runs 768 threads per CTA: its phases keep at most ~650 threads busy (two per block row of a
// <= 350-row slice), and every thread more lengthens the block barriers and the releasing cluster arrives
// (measured at 10k faces: 512 threads 1.94 ms, 768 1.73 ms, 1024 1.84 ms per solve; the grid modes lose with fewer)
static constexpr int kPcgBlockCluster = 768;

struct PcgParams {
    int nRows;                 // block rows (= global vertices)
    const int32_t* rowPtr; const int32_t* colIdx; const double* val; const double* minv;
    const double* rhs; int negate;
    const double* rowScale;    // 2 per row or nullptr: the system was scaled symmetrically, A~ = S A S: solve A~ y = S b, return x = S y
    double* x; double* r; double* z; double* d; double* d2; double* Ap;
    double* partials;          // 2 x gridDim SyncSlots (64 B each), double buffered by epoch parity
    double* scal;              // device scalar block
    double relTol; int maxIt;
    int scaledNorm;            // 1: converge in the block-Jacobi-scaled norm sqrt(r^T D^-1 r) (D = the 2x2 diagonal blocks); 0: plain 2-norm
    int maxBlkPerCta;          // SMEM mode: capacity of the per-CTA block arrays
    long long* dbg;            // optional: per-phase clock64 totals of CTA 0 (OCB_PCG_DEBUG=1)
    const int32_t* vertOf;     // solver row -> internal vertex (rhs gather / result scatter); nullptr = identity
    double* xOut;              // result in INTERNAL vertex order (x is the working copy in solver order)
    size_t masSmemOff;         // byte offset of the MAS scratch in dynamic shared memory
    size_t haloSmemOff;        // cluster mode: byte offset of the halo buffer (kHaloCap double2) in dynamic shared memory; 0 = none
    int32_t* haloIdx;          // cluster mode: gridDim x kHaloCap packed (owner << 20 | local row) of every halo slot
    int haloCap;               // slots available per CTA (what the shared-memory budget leaves, <= kHaloCap)
    MasView mas;               // mas.L == 0: block-Jacobi only
};

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

// ---------------------------------------------------------------------------------------------
// grid-wide all-reduce that doubles as the grid barrier.  No atomics and no second round trip: every
// CTA publishes each partial as ONE 16-byte packet {value, epoch} in its own slot (after a gpu-scope
// fence that makes the CTA's vector stores visible first); warp 0 of every CTA polls all packets with
// relaxed 16-byte loads -- the load that sees the epoch also carries the value -- and sums them in
// the same fixed order, so every CTA gets bitwise identical results.  Slots are double-buffered by
// epoch parity: a CTA can publish epoch e+2 only after every CTA has published e+1, i.e. after every
// CTA has finished reading epoch e.  Cross-CTA vectors are read through L2 (__ldcg) afterwards, so no
// L1 invalidation (acquire) is needed anywhere.
struct __align__(16) SyncPacket { double v; unsigned long long epoch; };
static constexpr int kSyncVals = 2;                         // packets per CTA slot
static constexpr int kPacketMaxCtas = 4096;                 // above this: counter barrier + one read pass (measured: no gain at 148 CTAs,
                                                            // the wait is skew between CTAs, not the mechanism -- kept for larger grids)
struct __align__(64) SyncSlot { SyncPacket p[kSyncVals]; unsigned long long pad[4]; };

__device__ __forceinline__ SyncPacket ld_packet(const SyncPacket* p)
{
    SyncPacket r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(*reinterpret_cast<unsigned long long*>(&r.v)), "=l"(r.epoch) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_packet(SyncPacket* p, double v, unsigned long long epoch)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(__double_as_longlong(v)), "l"(epoch) : "memory");
}

// The all-reduce is split into ARRIVE (publish this CTA's partials) and WAIT (collect everybody's), so that
// CTA-local work that does not depend on the result -- the preconditioner's local group solves -- runs while the
// slower CTAs are still on their way to the barrier.
static constexpr int kPollMax = 256;          // CTAs whose packets are polled by one thread each
struct ReduceSmem { double sm[kSyncVals][kPcgBlock / 32]; double bc[kSyncVals]; double vals[kSyncVals][kPollMax]; };

template <int NV>
__device__ __forceinline__ void grid_allreduce_arrive(ReduceSmem& R, SyncSlot* slots, int nB, unsigned long long epoch, double (&loc)[NV])
{
    static_assert(NV <= kSyncVals, "slot too small");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = loc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) R.sm[k][warp] = t;
    }
    __syncthreads();                       // orders every thread's vector stores before the fence below
    if (warp == 0) {
        SyncSlot* base = slots + (size_t)(epoch & 1ull) * nB;
        double t[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            t[k] = lane < (int)(blockDim.x >> 5) ? R.sm[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
        }
        if (lane == 0) {
            if (nB == 1) {
#pragma unroll
                for (int k = 0; k < NV; ++k) R.bc[k] = t[k];
            } else if (nB > kPacketMaxCtas) {
#pragma unroll
                for (int k = 0; k < NV; ++k) base[blockIdx.x].p[k].v = t[k];
            } else {
                __threadfence();           // cumulative: the whole CTA's stores are visible before the packets
#pragma unroll
                for (int k = 0; k < NV; ++k) st_packet(&base[blockIdx.x].p[k], t[k], epoch);
            }
        }
    }
}

template <int NV>
__device__ __forceinline__ void grid_allreduce_wait(ReduceSmem& R, SyncSlot* slots, int nB, unsigned long long epoch, double (&out)[NV])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SyncSlot* base = slots + (size_t)(epoch & 1ull) * nB;
    if (nB > kPacketMaxCtas) {
        // many CTAs: all-to-all packet polling costs O(nB^2) L2 requests; cross ONE counter barrier instead and read
        // all partials once, in the same fixed order everywhere.
        cooperative_groups::this_grid().sync();
        if (warp == 0) {
            double acc[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[k] = 0.0;
            for (int b = lane; b < nB; b += 32)
#pragma unroll
                for (int k = 0; k < NV; ++k) acc[k] += __ldcg(&base[b].p[k].v);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
                if (lane == 0) R.bc[k] = acc[k];
            }
        }
    } else if (nB > 1 && nB <= kPollMax) {
        // one thread per peer: all packets are in flight at once (ONE L2 round trip instead of nB/32 serial ones),
        // then warp 0 adds them in a fixed order, identical in every CTA
        if ((int)threadIdx.x < nB) {
            const int b = threadIdx.x;
            SyncPacket q[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) q[k] = ld_packet(&base[b].p[k]);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                while (q[k].epoch < epoch) q[k] = ld_packet(&base[b].p[k]);
                R.vals[k][b] = q[k].v;
            }
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                double acc = 0.0;
                for (int b = lane; b < nB; b += 32) acc += R.vals[k][b];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) R.bc[k] = acc;
            }
        }
    } else if (nB > 1 && warp == 0) {
        double acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = 0.0;
        for (int b = lane; b < nB; b += 32) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                SyncPacket q = ld_packet(&base[b].p[k]);
                while (q.epoch < epoch) q = ld_packet(&base[b].p[k]);
                acc[k] += q.v;
            }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            if (lane == 0) R.bc[k] = acc[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = R.bc[k];
    __syncthreads();                       // sm / bc may be rewritten by the next call
}

// ---------------------------------------------------------------------------------------------
// CLUSTER mode (small systems, <= ~17k faces): the whole solve runs in ONE thread-block cluster of 16 CTAs
// (16 SMs, slices resident in their shared memory).  The all-reduce then needs no global memory at all:
// every CTA drops its partial into every peer's shared memory through DSMEM and the hardware cluster
// barrier (arrive.release / wait.acquire, which also orders the global z/d stores) replaces the polled
// packets: ~0.3 us instead of ~1-3 us.
static constexpr int kClusterSize = 16;
template <int NV>
__device__ __forceinline__ void cluster_allreduce_arrive(ReduceSmem& R, double (*part)[kSyncVals][kClusterSize], unsigned long long epoch, double (&loc)[NV])
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = loc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) R.sm[k][warp] = t;
    }
    __syncthreads();
    // every CTA drops its partial into every peer's buffer (remote stores), then the cluster barrier; reading the
    // partials remotely AFTER the barrier instead was measured 2x slower (remote shared-memory loads ~2 us here)
    if (warp == 0) {
        const unsigned rank = cluster.block_rank();
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = lane < (int)(blockDim.x >> 5) ? R.sm[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane < kClusterSize) {          // lane l writes this CTA's partial into peer l's buffer
                double* remote = cluster.map_shared_rank(&part[epoch & 1ull][k][rank], lane);
                *remote = t;
            }
        }
    }
    // barrier.cluster.arrive.release by EVERY thread: each thread's own shared-memory stores (z, d slices that the peers
    // read through DSMEM) and global stores must be released by that thread.  A single cluster-scope fence by thread 0
    // followed by relaxed arrives saved the ~10 % membar stalls this costs but produced wrong search directions in ~1 of
    // 3 runs (tools/gpu_freerun_check.py): cumulativity does not cover the other threads' st.shared here.
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
template <int NV>
__device__ __forceinline__ void cluster_allreduce_wait(ReduceSmem& R, double (*part)[kSyncVals][kClusterSize], unsigned long long epoch, double (&out)[NV])
{
    static_assert(kClusterSize == 16 && NV <= 2, "one half-warp per value");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {                 // lanes 0-15: value 0, lanes 16-31: value 1; same tree in every CTA
        const int lane = threadIdx.x, k = lane >> 4;
        double t = k < NV ? part[epoch & 1ull][k][lane & 15] : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((lane & 15) == 0 && k < NV) R.bc[k] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = R.bc[k];
}

// ---------------------------------------------------------------------------------------------
// CTA-resident slice (SMEM mode): when a CTA's rows fit in shared memory (<= ~200 KB: up to ~170k faces
// over 148 CTAs) the matrix slice and all own-row vectors live there for the whole solve, so that an
// iteration touches global memory only for the halo gathers (z, d of neighbour rows), the two vectors
// neighbours read (z, d) and the sync packets.  Otherwise (GLOBAL mode) the slice is streamed from L2/HBM.
struct Slice {
    // ELL layout, slot-major: entry (k, r) at k * nLoc + r, so that consecutive threads (rows) touch consecutive
    // words: conflict-free shared-memory reads with ONE thread per scalar row and no shuffles
    int32_t* len;      // nLoc: blocks of the row held in the slice (<= W; longer rows continue in the global BSR)
    int32_t* col;      // W * nLoc
    double2* val2;     // W * nLoc * 2  (entry (k, r): [top row, bottom row] of the 2x2 block)
    double2* x; double2* r; double2* Ap; double2* z; double2* d[2];   // nLoc each; d is double-buffered like the global copy
    double4* minv;     // nLoc
    int W, nLoc;
};
__host__ __device__ inline size_t slice_bytes(int nLoc, int W)
{
    size_t b = 0;
    b += (((size_t)nLoc * 4 + 15) / 16) * 16;
    b += (((size_t)W * nLoc * 4 + 15) / 16) * 16;
    b += (size_t)W * nLoc * 32;
    b += (size_t)nLoc * 16 * 6;
    b += (size_t)nLoc * 32;
    return b;
}
__device__ __forceinline__ Slice carve(unsigned char* base, int nLoc, int W)
{
    Slice S; size_t o = 0;
    S.W = W; S.nLoc = nLoc;
    S.len = reinterpret_cast<int32_t*>(base + o);    o += (((size_t)nLoc * 4 + 15) / 16) * 16;
    S.col = reinterpret_cast<int32_t*>(base + o);    o += (((size_t)W * nLoc * 4 + 15) / 16) * 16;
    S.val2 = reinterpret_cast<double2*>(base + o);   o += (size_t)W * nLoc * 32;
    S.x = reinterpret_cast<double2*>(base + o);      o += (size_t)nLoc * 16;
    S.r = reinterpret_cast<double2*>(base + o);      o += (size_t)nLoc * 16;
    S.Ap = reinterpret_cast<double2*>(base + o);     o += (size_t)nLoc * 16;
    S.z = reinterpret_cast<double2*>(base + o);      o += (size_t)nLoc * 16;
    S.d[0] = reinterpret_cast<double2*>(base + o);   o += (size_t)nLoc * 16;
    S.d[1] = reinterpret_cast<double2*>(base + o);   o += (size_t)nLoc * 16;
    S.minv = reinterpret_cast<double4*>(base + o);
    return S;
}

// SMEM-mode SpMV with the fused direction update: one thread per SCALAR row (thread pair = block row), the row's
// blocks walked serially out of the ELL slice.  ~14 instructions per block and no shuffles: the 16-lanes-per-row
// mapping of the streaming path below costs ~10x more issue slots, which is what bounds a solve that runs on few SMs.
// Cluster mode: the off-slice entries of a CTA's ELL slice (~8 % of them) used to cost one distributed-shared-memory round
// trip EACH, issued one after the other from inside the row loop (plus an integer division for the owner).  Now every
// such entry owns a slot of a per-CTA halo buffer: at the start of the SpMV all slots are filled at once (z + beta d of the
// owner's shared memory, every remote load in flight together: ONE round trip), and the row loop reads local shared memory
// only.  The value in a slot is exactly the z + beta d the loop used to form, so the sums are bit-identical.
static constexpr int kHaloCap = 512;
template <bool DSMEM>
__device__ __forceinline__ double spmv_fused_ell(const PcgParams& P, const Slice& S, int rowBeg, int rowEnd, int rowsPer, const double* z,
                                                 const double* dOld, double* dNew, double beta,
                                                 const double2* __restrict__ sdOld, double2* __restrict__ sdNew,
                                                 double2* __restrict__ haloV = nullptr, int nHalo = 0)
{
    if (DSMEM && haloV) {
        cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
        const int32_t* hi = P.haloIdx + (size_t)blockIdx.x * kHaloCap;
        for (int h = threadIdx.x; h < nHalo; h += blockDim.x) {
            const int packed = __ldg(hi + h), owner = packed >> 20, lc = packed & 0xfffff;
            const double2 zz = cluster.map_shared_rank(S.z, owner)[lc];
            const double2 dd = cluster.map_shared_rank(const_cast<double2*>(sdOld), owner)[lc];
            haloV[h] = make_double2(zz.x + beta * dd.x, zz.y + beta * dd.y);
        }
        __syncthreads();
    }
    const double2* z2 = reinterpret_cast<const double2*>(z);
    const double2* d2 = reinterpret_cast<const double2*>(dOld);
    const double2* __restrict__ gval2 = reinterpret_cast<const double2*>(P.val);
    const int nLoc = rowEnd - rowBeg, half = threadIdx.x & 1;
    double dotAcc = 0.0;
    for (int lr = threadIdx.x >> 1; lr < nLoc; lr += blockDim.x >> 1) {
        const int len = S.len[lr];
        double acc = 0.0;
#pragma unroll 4
        for (int k = 0; k < len; ++k) {
            const int c = S.col[k * S.nLoc + lr];
            const double2 a = S.val2[2 * (k * S.nLoc + lr) + half];
            double2 zz, dd;
            if (DSMEM && c < 0) {                  // halo slot: z + beta d already formed
                const double2 hv = haloV[-c - 1];
                acc += a.x * hv.x + a.y * hv.y;
                continue;
            }
            if (c >= rowBeg && c < rowEnd) { zz = S.z[c - rowBeg]; dd = sdOld[c - rowBeg]; }     // own range: no global traffic
            else if (DSMEM) {                      // halo straight out of the owner CTA's shared memory (slots exhausted)
                const int owner = c / rowsPer, lc = c - owner * rowsPer;
                cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
                zz = cluster.map_shared_rank(S.z, owner)[lc];
                dd = cluster.map_shared_rank(const_cast<double2*>(sdOld), owner)[lc];
            }
            else { zz = __ldcg(z2 + c); dd = __ldcg(d2 + c); }
            acc += a.x * (zz.x + beta * dd.x) + a.y * (zz.y + beta * dd.y);
        }
        const int row = rowBeg + lr;
        const int gEnd = __ldg(P.rowPtr + row + 1);
        for (int b = __ldg(P.rowPtr + row) + len; b < gEnd; ++b) {       // rows longer than the slice width (rare)
            const int c = __ldg(P.colIdx + b);
            const double2 a = __ldg(gval2 + 2 * (size_t)b + half);
            double2 zz, dd;
            if (DSMEM) {
                const int owner = c / rowsPer, lc = c - owner * rowsPer;
                cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
                zz = cluster.map_shared_rank(S.z, owner)[lc];
                dd = cluster.map_shared_rank(const_cast<double2*>(sdOld), owner)[lc];
            } else { zz = __ldcg(z2 + c); dd = __ldcg(d2 + c); }
            acc += a.x * (zz.x + beta * dd.x) + a.y * (zz.y + beta * dd.y);
        }
        const double dn = reinterpret_cast<const double*>(S.z)[2 * lr + half] + beta * reinterpret_cast<const double*>(sdOld)[2 * lr + half];
        reinterpret_cast<double*>(sdNew)[2 * lr + half] = dn;
        reinterpret_cast<double*>(S.Ap)[2 * lr + half] = acc;
        if (!DSMEM) dNew[2 * (size_t)row + half] = dn;          // peers of a cluster read it from shared memory instead
        dotAcc += acc * dn;
    }
    return dotAcc;
}

// Ap[own rows] = A * d_new with d_new = z + beta * d_old evaluated on the fly for the gathered columns
// (so no barrier is needed between the direction update and the SpMV); writes d_new for the own rows
// into the other direction buffer.  z and d_old of other CTAs were written before the last grid sync:
// they are read through L2 (__ldcg), never through the non-coherent L1.
// Mapping: 16 lanes per block row (lane pair = one 2x2 block, even lane its top row, odd lane its bottom
// row), kUnroll row pairs per warp in flight so that every lane has several independent 16-byte loads
// outstanding (HBM needs ~26 KB in flight per SM).
static constexpr int kUnroll = 2;
__device__ __forceinline__ double spmv_fused(const PcgParams& P, int rowBeg, int rowEnd, const double* z,
                                             const double* dOld, double* dNew, double beta)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 15, kblk = sub >> 1, half = sub & 1;
    const double2* __restrict__ gval2 = reinterpret_cast<const double2*>(P.val);
    const double2* z2 = reinterpret_cast<const double2*>(z);
    const double2* d2 = reinterpret_cast<const double2*>(dOld);
    double dotAcc = 0.0;
    for (int rowBase = rowBeg + warp * 2 * kUnroll; rowBase < rowEnd; rowBase += (kPcgBlock / 32) * 2 * kUnroll) {
        int row[kUnroll], b0[kUnroll], end[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            row[u] = rowBase + 2 * u + (lane >> 4);
            const bool active = row[u] < rowEnd;
            b0[u] = (active ? __ldg(P.rowPtr + row[u]) : 0) + kblk;
            end[u] = active ? __ldg(P.rowPtr + row[u] + 1) : 0;
        }
        int col[kUnroll]; double2 a[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const bool has = b0[u] < end[u];
            col[u] = has ? __ldg(P.colIdx + b0[u]) : -1;
            a[u] = has ? __ldg(gval2 + 2 * (size_t)b0[u] + half) : make_double2(0.0, 0.0);
        }
        double2 zz[kUnroll], dd[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            zz[u] = col[u] >= 0 ? __ldcg(z2 + col[u]) : make_double2(0.0, 0.0);
            dd[u] = col[u] >= 0 ? __ldcg(d2 + col[u]) : make_double2(0.0, 0.0);
        }
        double acc[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            acc[u] = a[u].x * (zz[u].x + beta * dd[u].x) + a[u].y * (zz[u].y + beta * dd[u].y);
            for (int b = b0[u] + 8; b < end[u]; b += 8) {          // rows with more than 8 blocks (rare)
                const int c = __ldg(P.colIdx + b);
                const double2 av = __ldg(gval2 + 2 * (size_t)b + half);
                const double2 zv = __ldcg(z2 + c), dv = __ldcg(d2 + c);
                acc[u] += av.x * (zv.x + beta * dv.x) + av.y * (zv.y + beta * dv.y);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            double t = acc[u];
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            if (row[u] < rowEnd && sub < 2) {
                const double dn = __ldcg(z + 2 * (size_t)row[u] + half) + beta * __ldcg(dOld + 2 * (size_t)row[u] + half);
                P.Ap[2 * (size_t)row[u] + half] = t;
                dNew[2 * (size_t)row[u] + half] = dn;
                dotAcc += t * dn;
            }
        }
    }
    return dotAcc;
}

// MODE 0: slice streamed from global, grid-wide packet all-reduce; 1: slice in shared memory, packets;
// 2: slice in shared memory, one 16-CTA cluster, DSMEM all-reduce
template <int MODE>
__global__ void __launch_bounds__(MODE == 2 ? kPcgBlockCluster : kPcgBlock, 1)
pcg_kernel(PcgParams P)
{
    constexpr bool SMEM = MODE >= 1;
    constexpr int kBlk = MODE == 2 ? kPcgBlockCluster : kPcgBlock;
    __shared__ double clusterPart[2][kSyncVals][kClusterSize];
    __shared__ ReduceSmem redSm;
#define ARRIVE(NV, loc) do { ++epoch; if (MODE == 2) cluster_allreduce_arrive<NV>(redSm, clusterPart, epoch, loc); else grid_allreduce_arrive<NV>(redSm, slots, nB, epoch, loc); } while (0)
#define WAIT(NV, out) do { if (MODE == 2) cluster_allreduce_wait<NV>(redSm, clusterPart, epoch, out); else grid_allreduce_wait<NV>(redSm, slots, nB, epoch, out); } while (0)
#define ALLREDUCE(NV, loc, out) do { ARRIVE(NV, loc); WAIT(NV, out); } while (0)
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int nB = gridDim.x;
    const int rowsPer = (P.nRows + nB - 1) / nB;
    const int rowBeg = min(P.nRows, (int)blockIdx.x * rowsPer), rowEnd = min(P.nRows, rowBeg + rowsPer);
    const int nLoc = rowEnd - rowBeg;
    SyncSlot* slots = reinterpret_cast<SyncSlot*>(P.partials);
    unsigned long long epoch = 0;
    // distributed shared memory may only be touched once every CTA of the cluster has started executing: without this
    // barrier the first all-reduce's remote stores could land in a CTA that is not resident yet (compute-sanitizer
    // racecheck: "block that might not have entered yet", profiles/r2_sanitizer_*.log)
    if (MODE == 2) cooperative_groups::this_cluster().sync();
    Slice S = {};
    if (SMEM) {
        S = carve(smemRaw, rowsPer, P.maxBlkPerCta);          // maxBlkPerCta carries the ELL width W here
        const double2* gv = reinterpret_cast<const double2*>(P.val);
        for (int lr = threadIdx.x >> 1; lr < nLoc; lr += kBlk / 2) {
            const int b0 = P.rowPtr[rowBeg + lr], len = min(S.W, P.rowPtr[rowBeg + lr + 1] - b0), half = threadIdx.x & 1;
            if (half == 0) S.len[lr] = len;
            for (int k = 0; k < len; ++k) {
                if (half == 0) S.col[k * S.nLoc + lr] = P.colIdx[b0 + k];
                S.val2[2 * (k * S.nLoc + lr) + half] = gv[2 * (size_t)(b0 + k) + half];
            }
        }
    }

    // cluster mode: give every off-slice entry of the slice a halo slot
    double2* haloV = nullptr;
    __shared__ int sHaloCount;
    int nHalo = 0;
    if (MODE == 2 && P.haloSmemOff && P.haloIdx) {
        haloV = reinterpret_cast<double2*>(smemRaw + P.haloSmemOff);
        if (threadIdx.x == 0) sHaloCount = 0;
        __syncthreads();
        for (int e = threadIdx.x; e < nLoc * S.W; e += kBlk) {
            const int lr = e % nLoc, k = e / nLoc;
            if (k >= S.len[lr]) continue;
            const int c = S.col[k * S.nLoc + lr];
            if (c >= rowBeg && c < rowEnd) continue;
            const int slot = atomicAdd(&sHaloCount, 1);
            if (slot < P.haloCap) {
                const int owner = c / rowsPer;
                P.haloIdx[(size_t)blockIdx.x * kHaloCap + slot] = (owner << 20) | (c - owner * rowsPer);
                S.col[k * S.nLoc + lr] = -(slot + 1);
            }
        }
        __syncthreads();
        nHalo = min(sHaloCount, P.haloCap);
    }

    const bool MAS = P.mas.L > 0;
    MasSmem MS = {};
    if (MAS) { MS = mas_carve(smemRaw + P.masSmemOff, P.mas); mas_init(P.mas, MS, blockIdx.x, rowBeg, rowEnd); }
    auto getR = [&](int lr) -> double2 { return SMEM ? S.r[lr] : reinterpret_cast<const double2*>(P.r)[rowBeg + lr]; };
    // z = z0 (block-Jacobi part, already stored) + coarse correction of the row's leaf; returns this thread's share of r.z
    auto mas_finish = [&]() -> double {
        double acc = 0.0;
        for (int row = rowBeg + threadIdx.x; row < rowEnd; row += kBlk) {
            const int lr = row - rowBeg;
            const float4 vi = MS.vinfo[lr];
            const double* e = MS.e + (size_t)(__float_as_int(vi.w) - MS.leaf0) * kMasDof;
            double2 zz = SMEM ? S.z[lr] : reinterpret_cast<const double2*>(P.z)[row];
            const double2 r2 = getR(lr);
            zz.x += mas_unpack(vi.x) * e[0] + mas_unpack(vi.y) * e[1] + mas_unpack(vi.z) * e[2];
            zz.y += mas_unpack(vi.x) * e[3] + mas_unpack(vi.y) * e[4] + mas_unpack(vi.z) * e[5];
            if (SMEM) S.z[lr] = zz;
            if (MODE != 2) reinterpret_cast<double2*>(P.z)[row] = zz;
            acc += r2.x * zz.x + r2.y * zz.y;
        }
        return acc;
    };

    // ---- init: x = 0, r = b, z = Minv r, d = 0 ; rz = r.z, bb = b.b
    double loc[2] = {0.0, 0.0}, red[2];
    for (int row = rowBeg + threadIdx.x; row < rowEnd; row += kBlk) {
        const size_t src = P.vertOf ? (size_t)P.vertOf[row] : (size_t)row;
        double b0 = P.negate ? -P.rhs[2 * src] : P.rhs[2 * src];
        double b1 = P.negate ? -P.rhs[2 * src + 1] : P.rhs[2 * src + 1];
        if (P.rowScale) { b0 *= P.rowScale[2 * (size_t)row]; b1 *= P.rowScale[2 * (size_t)row + 1]; }
        const double4 m = reinterpret_cast<const double4*>(P.minv)[row];
        double2 zz; zz.x = m.x * b0 + m.y * b1; zz.y = m.y * b0 + m.w * b1;
        if (SMEM) {
            const int lr = row - rowBeg;
            S.x[lr] = make_double2(0.0, 0.0); S.r[lr] = make_double2(b0, b1); S.z[lr] = zz;
            S.d[0][lr] = make_double2(0.0, 0.0); S.d[1][lr] = make_double2(0.0, 0.0);
            S.minv[lr] = m;
        } else {
            reinterpret_cast<double2*>(P.x)[row] = make_double2(0.0, 0.0);
            reinterpret_cast<double2*>(P.r)[row] = make_double2(b0, b1);
        }
        reinterpret_cast<double2*>(P.z)[row] = zz;
        reinterpret_cast<double2*>(P.d)[row] = make_double2(0.0, 0.0);
        reinterpret_cast<double2*>(P.d2)[row] = make_double2(0.0, 0.0);
        loc[0] += b0 * zz.x + b1 * zz.y;
        loc[1] += b0 * b0 + b1 * b1;
    }
    if (MAS) { __syncthreads(); mas_restrict(P.mas, MS, blockIdx.x, getR); }
    ARRIVE(2, loc);
    if (MAS) mas_local_solves(P.mas, MS);
    WAIT(2, red);
    double rz = red[0];
    const double bb = red[1];
    // Optional (ocb_set_option "pcg_scaled_norm"): convergence in the norm the block-Jacobi part of the preconditioner
    // induces, r^T D^-1 r against b^T D^-1 b (both ride on the all-reduce that exists anyway), so that every DOF converges
    // relative to its own stiffness (the scaffold's rows are 4-8 orders of magnitude softer than a distorted mesh's).
    const double bzb = red[0];
    if (MAS && bb > 0.0) {
        mas_down<MODE == 2 ? 9 : 5>(P.mas, MS, blockIdx.x);
        double lz[1] = {mas_finish()}, rzv[1];
        ALLREDUCE(1, lz, rzv);
        rz = rzv[0];
    }
    const double tol2 = P.relTol * P.relTol * (P.scaledNorm ? bzb : bb);
    double rr = bb, rDr = bzb, beta = 0.0;
    int it = 0, status = 0, cur = 0;
    double rzFail = 0.0;                          // diagnostics of a rejected preconditioner (status 3): the offending r.M^-1 r
    if (bb > 0.0 && !(rz > 0.0)) { status = 3; rzFail = rz; }
    if (bb > 0.0 && status == 0) {
        for (;;) {
            // ---- phase A: d_new = z + beta d_old (fused) ; Ap = A d_new ; pAp
            double* dNew = cur ? P.d : P.d2;
            long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
            if (P.dbg) t0 = clock64();
            double2* sdOld = cur ? S.d[1] : S.d[0];
            double2* sdNew = cur ? S.d[0] : S.d[1];
            double* dOldG = cur ? P.d2 : P.d;
            double la[1] = {SMEM ? spmv_fused_ell<MODE == 2>(P, S, rowBeg, rowEnd, rowsPer, P.z, dOldG, dNew, beta, sdOld, sdNew, haloV, nHalo)
                                 : spmv_fused(P, rowBeg, rowEnd, P.z, dOldG, dNew, beta)}, ra[1];
            if (P.dbg) { __syncthreads(); t1 = clock64(); }
            ALLREDUCE(1, la, ra);
            if (P.dbg) t2 = clock64();
            const double dAd = ra[0];
            if (!(dAd > 0.0)) { status = 2; break; }
            const double alpha = rz / dAd;
            // ---- phase B (own rows only): x += alpha d ; r -= alpha Ap ; z = Minv r ; rz', rr
            loc[0] = 0.0; loc[1] = 0.0;
            for (int lr = threadIdx.x; lr < ((nLoc + 31) & ~31); lr += kBlk) {        // whole warps: the leaf restriction shuffles
                const bool valid = lr < nLoc;
                const int row = rowBeg + lr;
                double2 r2 = make_double2(0.0, 0.0);
                if (valid) {
                    const double2 dd = SMEM ? sdNew[lr] : reinterpret_cast<const double2*>(dNew)[row];
                    const double2 ap = SMEM ? S.Ap[lr] : reinterpret_cast<const double2*>(P.Ap)[row];
                    double2 xx = SMEM ? S.x[lr] : reinterpret_cast<double2*>(P.x)[row];
                    r2 = SMEM ? S.r[lr] : reinterpret_cast<double2*>(P.r)[row];
                    xx.x += alpha * dd.x; xx.y += alpha * dd.y;
                    r2.x -= alpha * ap.x; r2.y -= alpha * ap.y;
                    const double4 m = SMEM ? S.minv[lr] : reinterpret_cast<const double4*>(P.minv)[row];
                    double2 zz; zz.x = m.x * r2.x + m.y * r2.y; zz.y = m.y * r2.x + m.w * r2.y;
                    if (SMEM) { S.x[lr] = xx; S.r[lr] = r2; S.z[lr] = zz; }
                    else { reinterpret_cast<double2*>(P.x)[row] = xx; reinterpret_cast<double2*>(P.r)[row] = r2; }
                    if (MAS ? !SMEM : MODE != 2) reinterpret_cast<double2*>(P.z)[row] = zz;
                    loc[0] += r2.x * zz.x + r2.y * zz.y;
                    loc[1] += r2.x * r2.x + r2.y * r2.y;
                }
                if (MAS) mas_restrict_leaf(MS, lr, valid, r2);          // one call site for all 32 lanes
            }
            long long u0 = 0, u1 = 0, u2 = 0, u3 = 0, u4 = 0, u5 = 0;
            if (P.dbg) { __syncthreads(); u0 = clock64(); }
            if (MAS) { __syncthreads(); mas_restrict(P.mas, MS, blockIdx.x, getR, 2); }
            if (P.dbg) { __syncthreads(); t3 = clock64(); }
            ARRIVE(2, loc);
            if (P.dbg) { __syncthreads(); u1 = clock64(); }
            if (MAS) mas_local_solves(P.mas, MS);
            if (P.dbg) { __syncthreads(); u2 = clock64(); }
            WAIT(2, red);
            if (P.dbg) u3 = clock64();
            double rzNew = red[0];
            rDr = red[0];                     // r^T D^-1 r (the MAS part of r.z is added further down)
            rr = red[1];
            ++it;
            cur ^= 1;
            if ((P.scaledNorm ? rDr : rr) <= tol2) { status = 0; break; }
            if (it >= P.maxIt) { status = 1; break; }
            if (MAS) {
                long long st[2] = {0, 0};
                mas_down<MODE == 2 ? 9 : 5>(P.mas, MS, blockIdx.x, P.dbg ? st : nullptr);
                if (P.dbg) { __syncthreads(); u4 = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) { P.dbg[13] += st[0] - u3; P.dbg[14] += st[1] - st[0]; P.dbg[15] += u4 - st[1]; } }
                double lz[1] = {mas_finish()}, rzv[1];
                if (P.dbg) { __syncthreads(); u5 = clock64(); }
                ALLREDUCE(1, lz, rzv);
                rzNew = rzv[0];
            }
            if (P.dbg && threadIdx.x == 0 && blockIdx.x == 0) {
                P.dbg[0] += t1 - t0; P.dbg[1] += t2 - t1; P.dbg[2] += t3 - t2; P.dbg[3] += clock64() - t3; P.dbg[4] += 1;
                // finer split: [5] phase B, [6] restriction, [7] arrive, [8] group solves, [9] wait, [10] coarse + down, [11] finish, [12] last all-reduce
                P.dbg[5] += u0 - t2; P.dbg[6] += t3 - u0; P.dbg[7] += u1 - t3; P.dbg[8] += u2 - u1; P.dbg[9] += u3 - u2;
                P.dbg[10] += u4 - u3; P.dbg[11] += u5 - u4; P.dbg[12] += clock64() - u5;
            }
            if (!(rzNew > 0.0)) { status = 3; rzFail = rzNew; break; }        // r.M^-1 r <= 0 (or NaN): the preconditioner is not SPD; the host retries with block-Jacobi
            beta = rzNew / rz;
            rz = rzNew;
        }
    }
    __syncthreads();
    // Truncated CG (Steihaug): when a direction of non-positive curvature shows up (d.Ad <= 0: the projected Hessian of an
    // extremely distorted start, entries spanning 60 orders of magnitude, is only PSD up to rounding) the iterate reached so
    // far is returned -- still a descent direction -- and, if that happens on the very first direction, the preconditioned
    // steepest-descent direction z = M^-1 b; the line search takes it from there, like it does behind the reference's LDL^T.
    const bool steepest = (status == 2 && it == 0);
    for (int row = rowBeg + threadIdx.x; row < rowEnd; row += kBlk) {
        const size_t dst = P.vertOf ? (size_t)P.vertOf[row] : (size_t)row;
        double2 xo = make_double2(0.0, 0.0);
        if (bb > 0.0) {
            if (steepest) xo = SMEM ? S.z[row - rowBeg] : reinterpret_cast<const double2*>(P.z)[row];
            else xo = SMEM ? S.x[row - rowBeg] : reinterpret_cast<const double2*>(P.x)[row];
            if (P.rowScale) { xo.x *= P.rowScale[2 * (size_t)row]; xo.y *= P.rowScale[2 * (size_t)row + 1]; }
        }
        reinterpret_cast<double2*>(P.xOut)[dst] = xo;
    }

    if (MODE == 2) cooperative_groups::this_cluster().sync();      // no CTA may exit while a peer still reads its shared memory
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.scal[S_PCG_ITERS] = (double)it;
        P.scal[S_PCG_RELRES] = bb > 0.0 ? (P.scaledNorm ? sqrt(rDr / bzb) : sqrt(rr / bb)) : 0.0;
        P.scal[S_PCG_STATUS] = (double)status;
        P.scal[S_PCG_BNORM] = sqrt(bb);
        if (status == 3) { P.scal[S_MISC0] = rzFail; P.scal[S_MISC1] = rz; }
    }
#undef ALLREDUCE
#undef ARRIVE
#undef WAIT
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
                    const double* __restrict__ val, double* __restrict__ minv, int* __restrict__ bad, double* __restrict__ scal)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        double a00 = 0.0, a01 = 0.0, a11 = 0.0;
        for (int b = rowPtr[row]; b < rowPtr[row + 1]; ++b)
            if (colIdx[b] == row) { a00 = val[4 * (size_t)b]; a01 = 0.5 * (val[4 * (size_t)b + 1] + val[4 * (size_t)b + 2]); a11 = val[4 * (size_t)b + 3]; }
        const double det = a00 * a11 - a01 * a01;
        if (!(a00 > 0.0) || !(det > 0.0)) { atomicAdd(bad, 1); scal[S_JACOBI_BAD] = 1.0; minv[4 * (size_t)row] = 1.0; minv[4 * (size_t)row + 1] = 0.0; minv[4 * (size_t)row + 2] = 0.0; minv[4 * (size_t)row + 3] = 1.0; continue; }
        minv[4 * (size_t)row] = a11 / det; minv[4 * (size_t)row + 1] = -a01 / det;
        minv[4 * (size_t)row + 2] = -a01 / det; minv[4 * (size_t)row + 3] = a00 / det;
    }
}

// Symmetric diagonal scaling A~ = S A S, S = diag(a_ii)^-1/2 (ocb_newton_step_ex; option "scale_system").  At a distorted start
// the entries of the projected Hessian span 30-60 orders of magnitude (benchmark meshes male_2, cat_noUV, horse: ||g||^2 up to
// 1e51); CG's recurrences on the unscaled matrix then lose positivity to cancellation after a few iterations (d.Ad <= 0 although
// A is PSD).  On the scaled system every row has unit diagonal, the 2-norm stopping test becomes the D-scaled one, and the
// preconditioner is built from A~.
__global__ void __launch_bounds__(256)
row_scale_kernel(int nRows, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, const double* __restrict__ val, double* __restrict__ scale)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256) {
        double a00 = 0.0, a11 = 0.0;
        for (int b = rowPtr[row]; b < rowPtr[row + 1]; ++b) if (colIdx[b] == row) { a00 = val[4 * (size_t)b]; a11 = val[4 * (size_t)b + 3]; }
        scale[2 * (size_t)row] = (a00 > 0.0 && a00 < 1e300) ? rsqrt(a00) : 1.0;
        scale[2 * (size_t)row + 1] = (a11 > 0.0 && a11 < 1e300) ? rsqrt(a11) : 1.0;
    }
}
__global__ void __launch_bounds__(256)
scale_matrix_kernel(int nRows, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, double* __restrict__ val, const double* __restrict__ scale)
{
    // 4 lanes per block row
    for (long w = (blockIdx.x * 256L + threadIdx.x) >> 2; w < nRows; w += (gridDim.x * 256L) >> 2) {
        const int row = (int)w;
        const double r0 = scale[2 * (size_t)row], r1 = scale[2 * (size_t)row + 1];
        for (int b = rowPtr[row] + (threadIdx.x & 3); b < rowPtr[row + 1]; b += 4) {
            const int col = colIdx[b];
            const double c0 = scale[2 * (size_t)col], c1 = scale[2 * (size_t)col + 1];
            double2* v = reinterpret_cast<double2*>(val + 4 * (size_t)b);
            double2 t0 = v[0], t1 = v[1];
            t0.x *= r0 * c0; t0.y *= r0 * c1; t1.x *= r1 * c0; t1.y *= r1 * c1;
            v[0] = t0; v[1] = t1;
        }
    }
}

#define KCHECK(c) do { (c)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return cuda_fail((c), _e, __func__); } while (0)

static PcgParams make_params(ocb_ctx* c)
{
    PcgParams P;
    P.nRows = c->nVtot; P.rowPtr = c->rowPtr.p; P.colIdx = c->colIdx.p; P.val = c->val.p; P.minv = c->minv.p;
    P.rhs = nullptr; P.negate = 0; P.rowScale = nullptr; P.x = c->px.p; P.xOut = c->p.p; P.vertOf = nullptr; P.masSmemOff = 0; P.haloSmemOff = 0; P.haloIdx = nullptr; P.haloCap = 0; P.mas = MasView(); P.mas.L = 0;
    P.r = c->pr.p; P.z = c->pz.p; P.d = c->pd.p; P.d2 = c->pd2.p; P.Ap = c->pAp.p;
    P.partials = c->partials.p; P.scal = c->dScal; P.relTol = 1e-12; P.maxIt = 1; P.scaledNorm = 1; P.maxBlkPerCta = 0; P.dbg = nullptr;
    return P;
}

// grid for the stand-alone SpMV
static int spmv_grid(ocb_ctx* c)
{
    const int rowsPerCtaPass = (kPcgBlock / 32) * 2;
    int need = (c->nVtot + rowsPerCtaPass - 1) / rowsPerCtaPass;
    int g = c->numSMs < need ? c->numSMs : need;
    return g < 1 ? 1 : g;
}

// Partition of the block rows over the persistent CTAs.  Small systems: few, fat CTAs (the grid-wide
// all-reduce costs ~0.9 us at 16 CTAs and ~3 us at 148) with the slice resident in shared memory;
// large systems: one CTA per SM, slice streamed from L2/HBM.
struct PcgPlan { int grid; bool smem; bool cluster; int maxBlk; size_t smemBytes; };
// upper bound of the MAS scratch of one CTA (the hierarchy is built after the plan)
static size_t mas_smem_estimate(int rowsPer, int grid)
{
    // mirrors mas_build_hierarchy: leaves of <= 8 rows, groups of <= 8 nodes inside a CTA, coarse level = the first
    // level with <= kMasCoarseMax DOFs over the whole grid (upper bounds: every CTA as full as the fullest)
    int local = 0, k = (rowsPer + kMasLeaf - 1) / kMasLeaf, nC = 0;
    for (;;) {
        local += k;
        if ((long)k * grid * kMasDof <= mas_coarse_cap(rowsPer * grid)) { nC = k * grid * kMasDof; break; }
        if (k <= 1) { nC = grid * kMasDof; break; }
        k = (k + kMasGroup - 1) / kMasGroup;
    }
    const int ldC = (nC + kMasCoarseBlk - 1) / kMasCoarseBlk * kMasCoarseBlk;
    return mas_smem_bytes(local + kMasMaxLevels, rowsPer, ldC) + 64;
}
static PcgPlan pcg_plan(ocb_ctx* c)
{
    static const int targetRows = []() { const char* e = getenv("OCB_PCG_ROWS_PER_CTA"); int v = e ? atoi(e) : 256; return v < 32 ? 32 : v; }();
    static const bool allowSmem = []() { const char* e = getenv("OCB_PCG_NO_SMEM"); return !(e && atoi(e)); }();
    const size_t limit = 214 * 1024;                        // slice + preconditioner scratch; the cluster mode adds its 8 KB halo buffer on top
    static const bool allowCluster = []() { const char* e = getenv("OCB_PCG_NO_CLUSTER"); return !(e && atoi(e)); }();
    PcgPlan pl; pl.smem = false; pl.cluster = false; pl.maxBlk = 0; pl.smemBytes = 0;
    const int n = c->nVtot;
    const int ellW = 12;                                    // constant: the plan must not depend on the pattern (the row order and the
                                                            // preconditioner hierarchy are built for this grid BEFORE the pattern
                                                            // exists); longer rows (valence > 11) continue in the global BSR
    auto slice_need = [&](int g, int& W) {
        W = ellW;
        return slice_bytes((n + g - 1) / g, ellW) + mas_smem_estimate((n + g - 1) / g, g);
    };
    if (allowSmem) {                                           // tiny system: ONE CTA, block-level reductions only
        int W = 0;
        const size_t bytes = slice_need(1, W);
        if (bytes <= limit) { pl.smem = true; pl.maxBlk = W; pl.smemBytes = bytes; pl.grid = 1; return pl; }
    }
    if (allowSmem && allowCluster && c->clusterOk != 0) {      // one 16-CTA cluster if every slice fits
        int maxBlk = 0;
        const size_t bytes = slice_need(kClusterSize, maxBlk);
        if (bytes <= limit) {
            if (c->clusterOk < 0) {                           // probe once: can such a cluster be scheduled?
                cudaFuncSetAttribute(pcg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(221 * 1024));
                cudaFuncSetAttribute(pcg_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(kClusterSize); cfg.blockDim = dim3(kPcgBlockCluster); cfg.dynamicSmemBytes = limit;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = kClusterSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int nClusters = 0;
                cudaError_t e = cudaOccupancyMaxActiveClusters(&nClusters, pcg_kernel<2>, &cfg);
                c->clusterOk = (e == cudaSuccess && nClusters >= 1) ? 1 : 0;
                cudaGetLastError();
            }
            if (c->clusterOk == 1) { pl.smem = true; pl.cluster = true; pl.maxBlk = maxBlk; pl.smemBytes = bytes; pl.grid = kClusterSize; return pl; }
        }
    }
    int g = (n + targetRows - 1) / targetRows;
    if (g > c->numSMs) g = c->numSMs;
    if (g < 1) g = 1;
    for (; allowSmem; ) {
        int maxBlk = 0;
        const size_t bytes = slice_need(g, maxBlk);
        if (bytes <= limit) { pl.smem = true; pl.maxBlk = maxBlk; pl.smemBytes = bytes; break; }
        if (g >= c->numSMs) break;
        g = std::min(c->numSMs, g + (g + 3) / 4);
    }
    if (!pl.smem) g = c->numSMs;
    pl.grid = g;
    return pl;
}

int pcg_plan_grid(ocb_ctx* c, int nRows)
{
    (void)nRows;
    return pcg_plan(c).grid;
}

// system vector between INTERNAL vertex order and solver row order
__global__ void __launch_bounds__(256)
gather_rows_kernel(int nRows, const int32_t* __restrict__ vertOf, const double2* __restrict__ in, double2* __restrict__ out, int toRows)
{
    for (int r = blockIdx.x * 256 + threadIdx.x; r < nRows; r += gridDim.x * 256) {
        const int v = vertOf ? vertOf[r] : r;
        if (toRows) out[r] = in[v]; else out[v] = in[r];
    }
}
int launch_gather_rows(ocb_ctx* c, const double* in, double* out, bool toRows)
{
    int grid = (c->nVtot + 255) / 256; if (grid > c->numSMs * 8) grid = c->numSMs * 8; if (grid < 1) grid = 1;
    gather_rows_kernel<<<grid, 256, 0, c->stream>>>(c->nVtot, c->vertOf.p, reinterpret_cast<const double2*>(in), reinterpret_cast<double2*>(out), toRows ? 1 : 0);
    KCHECK(c);
    return 0;
}

int launch_spmv(ocb_ctx* c, const double* dx, double* dy)
{
    ProfScope prof(c, K_SPMV);
    PcgParams P = make_params(c);
    spmv_kernel<<<spmv_grid(c), kPcgBlock, 0, c->stream>>>(P, dx, dy);
    KCHECK(c);
    return 0;
}

// relative diagonal regularisation A_ii *= (1 + delta): used ONLY after a CG breakdown (see ocb_newton_step_ex)
__global__ void __launch_bounds__(256)
diag_shift_kernel(int nRows, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, double* __restrict__ val, double factor)
{
    for (int row = blockIdx.x * 256 + threadIdx.x; row < nRows; row += gridDim.x * 256)
        for (int b = rowPtr[row]; b < rowPtr[row + 1]; ++b)
            if (colIdx[b] == row) { val[4 * (size_t)b] *= factor; val[4 * (size_t)b + 3] *= factor; }
}
int launch_diag_shift(ocb_ctx* c, double delta)
{
    ProfScope prof(c, K_JACOBI_SETUP);
    int grid = (c->nVtot + 255) / 256; if (grid > c->numSMs * 8) grid = c->numSMs * 8; if (grid < 1) grid = 1;
    diag_shift_kernel<<<grid, 256, 0, c->stream>>>(c->nVtot, c->rowPtr.p, c->colIdx.p, c->val.p, 1.0 + delta);
    KCHECK(c);
    return 0;
}

int launch_scale_system(ocb_ctx* c)
{
    ProfScope prof(c, K_JACOBI_SETUP);
    OCB_CUDA(c, c->rowScale.reserve(2 * (size_t)c->nVtot + 2, c->stream));
    int grid = (c->nVtot + 255) / 256; if (grid > c->numSMs * 8) grid = c->numSMs * 8; if (grid < 1) grid = 1;
    row_scale_kernel<<<grid, 256, 0, c->stream>>>(c->nVtot, c->rowPtr.p, c->colIdx.p, c->val.p, c->rowScale.p);
    KCHECK(c);
    int g2 = (4 * c->nVtot + 255) / 256; if (g2 > c->numSMs * 8) g2 = c->numSMs * 8; if (g2 < 1) g2 = 1;
    scale_matrix_kernel<<<g2, 256, 0, c->stream>>>(c->nVtot, c->rowPtr.p, c->colIdx.p, c->val.p, c->rowScale.p);
    KCHECK(c);
    c->systemScaled = true;
    return 0;
}

int launch_jacobi_setup(ocb_ctx* c, bool check)
{
    int* bad = reinterpret_cast<int*>(c->sync.p + 8);
    OCB_CUDA(c, cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    OCB_CUDA(c, cudaMemsetAsync(c->dScal + S_JACOBI_BAD, 0, sizeof(double), c->stream));
    OCB_CUDA(c, c->minv.reserve(4 * (size_t)c->nVtot, c->stream));
    int grid = (c->nVtot + 255) / 256; if (grid > c->numSMs * 8) grid = c->numSMs * 8; if (grid < 1) grid = 1;
    {
        ProfScope prof(c, K_JACOBI_SETUP);
        jacobi_setup_kernel<<<grid, 256, 0, c->stream>>>(c->nVtot, c->rowPtr.p, c->colIdx.p, c->val.p, c->minv.p, bad, c->dScal);
        KCHECK(c);
    }
    if (!check) return 0;
    int hBad = 0;
    OCB_CUDA(c, cudaMemcpyAsync(&hBad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (hBad) return set_err(c, OCB_ERR_BREAKDOWN, "a diagonal 2x2 block of the matrix is not positive definite");
    return 0;
}

int launch_pcg(ocb_ctx* c, const double* d_rhs, bool negate_rhs, double rel_tol, int max_it, bool allowMas)
{
    ProfScope prof(c, K_PCG);
    const size_t n = c->nSys();
    OCB_CUDA(c, c->pr.reserve(n, c->stream)); OCB_CUDA(c, c->pz.reserve(n, c->stream));
    OCB_CUDA(c, c->pd.reserve(n, c->stream)); OCB_CUDA(c, c->pAp.reserve(n, c->stream));
    OCB_CUDA(c, c->pd2.reserve(n, c->stream));
    OCB_CUDA(c, c->p.reserve(n, c->stream));
    OCB_CUDA(c, c->px.reserve(n, c->stream));
    const PcgPlan pl = pcg_plan(c);
    const int grid = pl.grid;
    const size_t slotDoubles = 2 * (size_t)grid * (sizeof(SyncSlot) / sizeof(double));
    OCB_CUDA(c, c->partials.reserve(slotDoubles + 64, c->stream));
    PcgParams P = make_params(c);
    P.rhs = d_rhs; P.negate = negate_rhs ? 1 : 0; P.relTol = rel_tol; P.maxIt = max_it; P.maxBlkPerCta = pl.maxBlk;
    P.scaledNorm = c->pcgPlainNorm ? 0 : 1;
    P.rowScale = c->systemScaled ? c->rowScale.p : nullptr;
    P.vertOf = c->vertOf.p;
    size_t smemBytes = pl.smem ? pl.smemBytes - mas_smem_estimate((c->nVtot + grid - 1) / grid, grid) : 0;     // the slice alone
    smemBytes = (smemBytes + 15) / 16 * 16;
    if (allowMas && c->masH.enabled && c->masH.grid == grid && c->masD.view.L > 0) {
        P.mas = c->masD.view;
        P.masSmemOff = smemBytes;
        // streaming mode has the shared memory to spare: keep the CTA's rows of the coarse inverse resident
        const size_t cinvBytes = (size_t)P.mas.maxOwnC * kMasDof * P.mas.ldC * 4;
        P.mas.cinvInSmem = (!pl.smem && cinvBytes <= 120 * 1024) ? 1 : 0;
        smemBytes += mas_smem_bytes(P.mas.maxLocalNodes, P.mas.rowsPer, P.mas.ldC, P.mas.cinvInSmem ? P.mas.maxOwnC * kMasDof : 0);
    }
    const size_t smemCap = 221 * 1024;      // + 5.2 KB static (reduction scratch, cluster partials) <= 227 KB per CTA
    static const bool haloOn = []() { const char* e = getenv("OCB_PCG_NO_HALO"); return !(e && atoi(e)); }();
    if (pl.cluster && haloOn) {             // the halo buffer takes what the budget leaves (7 KB = 448 slots at 10k faces; ~200 are used)
        smemBytes = (smemBytes + 15) / 16 * 16;
        const size_t avail = smemCap > smemBytes ? smemCap - smemBytes : 0;
        const int slots = (int)std::min<size_t>(kHaloCap, avail / sizeof(double2));
        if (slots >= 64) {
            P.haloSmemOff = smemBytes; P.haloCap = slots;
            smemBytes += (size_t)slots * sizeof(double2);
            OCB_CUDA(c, c->pcgHalo.reserve((size_t)grid * kHaloCap + 4, c->stream));
            P.haloIdx = c->pcgHalo.p;
        }
    }
    if (smemBytes > smemCap) return set_err(c, OCB_ERR_STATE, "PCG: shared-memory plan exceeds the SM capacity");
    OCB_CUDA(c, cudaMemsetAsync(c->partials.p, 0, slotDoubles * sizeof(double), c->stream));
    static const bool dbgOn = []() { const char* e = getenv("OCB_PCG_DEBUG"); return e && atoi(e); }();
    long long* dDbg = nullptr;
    if (dbgOn) { cudaMalloc((void**)&dDbg, 16 * sizeof(long long)); cudaMemset(dDbg, 0, 16 * sizeof(long long)); P.dbg = dDbg; }
    void* args[] = {&P};
    if (!c->pcgSmemAttr) {
        OCB_CUDA(c, cudaFuncSetAttribute(pcg_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemCap));
        OCB_CUDA(c, cudaFuncSetAttribute(pcg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemCap));
        OCB_CUDA(c, cudaFuncSetAttribute(pcg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemCap));
        c->pcgSmemAttr = smemCap;
    }
    if (pl.cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kPcgBlockCluster); cfg.dynamicSmemBytes = smemBytes; cfg.stream = c->stream;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kClusterSize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        OCB_CUDA(c, cudaLaunchKernelEx(&cfg, pcg_kernel<2>, P));
    } else if (pl.smem) {
        OCB_CUDA(c, cudaLaunchCooperativeKernel((void*)pcg_kernel<1>, dim3(grid), dim3(kPcgBlock), args, smemBytes, c->stream));
    } else {
        OCB_CUDA(c, cudaLaunchCooperativeKernel((void*)pcg_kernel<0>, dim3(grid), dim3(kPcgBlock), args, smemBytes, c->stream));
    }
    c->launches++;
    if (dbgOn) {
        long long h[16];
        cudaStreamSynchronize(c->stream);
        cudaMemcpy(h, dDbg, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(dDbg);
        const double n = h[4] > 0 ? (double)h[4] : 1.0;
        fprintf(stderr, "[ocb pcg] mode %s grid %d iters %lld  cycles/iter: spmv %.0f  sync1 %.0f  update %.0f  sync2 %.0f\n",
                pl.cluster ? "cluster" : (pl.smem ? "smem" : "global"), grid, h[4], h[0] / n, h[1] / n, h[2] / n, h[3] / n);
        fprintf(stderr, "[ocb pcg]   coarse/down = coarse residual load %.0f + coarse rows %.0f + down sweep %.0f\n", h[13] / n, h[14] / n, h[15] / n);
        fprintf(stderr, "[ocb pcg]   update = phaseB %.0f + restrict %.0f ; sync2 = arrive %.0f + group solves %.0f + wait %.0f + coarse/down %.0f + finish %.0f + allreduce %.0f\n",
                h[5] / n, h[6] / n, h[7] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n, h[12] / n);
    }
    return 0;
}

}  // namespace ocb
