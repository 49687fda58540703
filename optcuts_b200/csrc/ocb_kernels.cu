// optcuts_b200 — element kernels (energy, gradient, Hessian + PSD projection + BSR scatter, step
// bound, line-search stepping), pattern helpers and small reductions.  sm_100a, fp64.
// Compiled with -fmad=false so per-element values are bit-identical to the CPU reference's
// (SSE2, no FMA) for identical inputs; sums differ only in association order.
//
// All of these are HBM-streaming kernels: per element they read 3 int32 vertex ids + 5 rest doubles
// (coalesced, SoA) and gather 3 double2 UVs (L2-resident: every vertex is shared by ~6 triangles).
// Grids are sized as a multiple of the SM count and grid-stride over the elements; reductions are
// two-level (warp shuffle -> shared -> one partial per block -> last block sums in fixed order) and
// therefore deterministic.
#include "ocb_internal.cuh"
#include "ocb_element.cuh"
#include <algorithm>

namespace ocb {

static constexpr int kBlock = 256;

// ---------------------------------------------------------------------------------------------
// Per-thread asynchronous prefetch queue (LDGSTS, cp.async): the streaming per-element data (vertex ids + rest
// doubles) of the NEXT kDepth-1 rounds of the grid-stride loop is copied global -> shared memory while the current
// round gathers its UVs and computes.  Every thread reads back only the slots it filled itself, so the queue needs
// no block barrier, costs no registers, and keeps (kDepth-1) x 52 B per thread in flight towards HBM: with the
// plain register loop the dependent UV gathers left the memory system idle half of the time (energy: 2.6 TB/s).
static constexpr int kDepth = 3;
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(void* s, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void* s, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// queue of {v0,v1,v2} (+ the 5 rest doubles of the value/gradient formulas when ND == 5) over mesh + air elements
template <int ND>
struct ElemQueue {
    int32_t qi[kDepth][3][kBlock];
    double qd[kDepth][ND > 0 ? ND : 1][kBlock];
    __device__ __forceinline__ void issue(const ElemView& M, const ElemView& A, int e, int total, int st) {
        if (e < total) {
            const bool isAir = e >= M.n;
            const ElemView& S = isAir ? A : M;
            const int t = isAir ? e - M.n : e;
            const int tid = threadIdx.x;
            cp_async4(&qi[st][0][tid], S.v0 + t); cp_async4(&qi[st][1][tid], S.v1 + t); cp_async4(&qi[st][2][tid], S.v2 + t);
            if (ND == 5) {
                cp_async8(&qd[st][0][tid], S.area + t); cp_async8(&qd[st][1][tid], S.areaSq + t);
                cp_async8(&qd[st][2][tid], S.e0 + t); cp_async8(&qd[st][3][tid], S.e1 + t); cp_async8(&qd[st][4][tid], S.d + t);
            }
        }
        cp_commit();      // an empty group keeps the group count in step with the loop
    }
};

static inline int grid_for(const ocb_ctx* c, long n, int perSM = 8) {
    long g = (n + kBlock - 1) / kBlock;
    long cap = (long)c->numSMs * perSM;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---------------------------------------------------------------------------------------------
// deterministic block reduction of NV values + grid finalisation by the last block
template <int NV, bool IS_MIN>
__device__ __forceinline__ void reduce_finalize(double (&v)[NV], double* __restrict__ partials,
                                                unsigned* __restrict__ ticket, double* __restrict__ out,
                                                const int* outSlots)
{
    __shared__ double sm[NV][kBlock / 32];
    __shared__ bool isLast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, t, o);
            t = IS_MIN ? fmin(t, other) : t + other;
        }
        if (lane == 0) sm[k][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = sm[k][0];
            for (int w = 1; w < kBlock / 32; ++w) t = IS_MIN ? fmin(t, sm[k][w]) : t + sm[k][w];
            partials[(size_t)blockIdx.x * NV + k] = t;
        }
        __threadfence();
        const unsigned tk = atomicAdd(ticket, 1u);
        isLast = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    // last block: every thread sums a strided subset in fixed order, then the same tree
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = IS_MIN ? INFINITY : 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kBlock)
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const double t = __ldcg(&partials[(size_t)b * NV + k]);
            acc[k] = IS_MIN ? fmin(acc[k], t) : acc[k] + t;
        }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, t, o);
            t = IS_MIN ? fmin(t, other) : t + other;
        }
        if (lane == 0) sm[k][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = sm[k][0];
            for (int w = 1; w < kBlock / 32; ++w) t = IS_MIN ? fmin(t, sm[k][w]) : t + sm[k][w];
            out[outSlots[k]] = t;
        }
        *ticket = 0u;
    }
}

struct Slots3 { int s[3]; };
struct Slots1 { int s[1]; };

// ---------------------------------------------------------------------------------------------
// a2/a13/a14: energy of mesh + air elements, optionally at x0 + alpha * p (line search)
template <bool STEPPED>
__global__ void __launch_bounds__(kBlock)
energy_kernel(ElemView M, ElemView A, const double* __restrict__ x, const double* __restrict__ p,
              double alpha, double* __restrict__ partials, unsigned* __restrict__ ticket,
              double* __restrict__ scal, Slots3 slots, int alphaFromStepBound)
{
    // first line-search trial without a host round trip: alpha = 0.99 * (step bound left by the previous launch),
    // the same product the host forms after its fetch (Optimizer.cpp:580)
    if (STEPPED && alphaFromStepBound) alpha = __dmul_rn(alpha, scal[S_STEP_BOUND]);
    __shared__ ElemQueue<5> Q;
    double acc[3] = {0.0, 0.0, 0.0};   // mesh sum, air sum, #elements with signed area < 0
    const int total = M.n + A.n, stride = gridDim.x * kBlock, tid = threadIdx.x;
    int e = blockIdx.x * kBlock + tid, st = 0;
#pragma unroll
    for (int k = 0; k < kDepth - 1; ++k) Q.issue(M, A, e + k * stride, total, k);
    for (; e < total; e += stride, st = (st + 1 == kDepth) ? 0 : st + 1) {
        Q.issue(M, A, e + (kDepth - 1) * stride, total, (st + kDepth - 1) % kDepth);
        cp_wait<kDepth - 1>();
        const bool isAir = e >= M.n;
        const ElemView& S = isAir ? A : M;
        const int i0 = Q.qi[st][0][tid], i1 = Q.qi[st][1][tid], i2 = Q.qi[st][2][tid];
        const double area = Q.qd[st][0][tid], A2 = Q.qd[st][1][tid], e0 = Q.qd[st][2][tid], e1 = Q.qd[st][3][tid], d = Q.qd[st][4][tid];
        Vec2 U1, U2, U3;
        if (STEPPED) { U1 = ld2_step(x, p, alpha, i0); U2 = ld2_step(x, p, alpha, i1); U3 = ld2_step(x, p, alpha, i2); }
        else { U1 = ld2(x, i0); U2 = ld2(x, i1); U3 = ld2(x, i2); }
        const double w = S.uniform ? 1.0 : area / S.surfaceArea;
        double dbArea;
        const double E = sd_energy(U2 - U1, U3 - U1, A2, e0, e1, d, w, dbArea);
        if (isAir) acc[1] += E; else acc[0] += E;
        if (dbArea < 0.0) acc[2] += 1.0;
    }
    reduce_finalize<3, false>(acc, partials, ticket, scal, slots.s);
}

__global__ void __launch_bounds__(kBlock)
energy_per_elem_kernel(ElemView M, const double* __restrict__ x, double* __restrict__ out)
{
    for (int t = blockIdx.x * kBlock + threadIdx.x; t < M.n; t += gridDim.x * kBlock) {
        const Vec2 U1 = ld2(x, M.v0[t]), U2 = ld2(x, M.v1[t]), U3 = ld2(x, M.v2[t]);
        const double w = M.uniform ? 1.0 : M.area[t] / M.surfaceArea;
        double dbArea;
        out[t] = sd_energy(U2 - U1, U3 - U1, M.areaSq[t], M.e0[t], M.e1[t], M.d[t], w, dbArea);
    }
}

// ---------------------------------------------------------------------------------------------
// a2/a4/a13: gradient + energy + ||g||^2 in ONE pass, assembled by a vertex gather (no atomics, no memset): the thread
// of vertex v walks its incident corners in ascending element order -- the order in which the reference's serial loop
// adds them (SymDirichletEnergy.cpp:264-298) -- first the mesh's, then (through g2l / the air-local numbering) the air
// mesh's, and forms  g_v = energyParam0 * g_mesh_v + (w_scaf/|Fa|) * g_air_v  exactly like Optimizer::computeGradient
// + Scaffold::augmentGradient (Optimizer.cpp:783-797, Scaffold.cpp:210-229).  With the element unit compiled
// -fmad=false the gradient is therefore bit-identical to the reference's, run after run.  A triangle's value is added
// by the thread of its corner 0, so the energy and the inversion count come out of the same pass.
struct GatherView {
    const int32_t* vcPtrM; const int32_t* vcIdxM;     // mesh corners by internal vertex
    const int32_t* vcPtrA; const int32_t* vcIdxA;     // air corners by air-local vertex
    const int32_t* g2l;                               // mesh vertex -> air-local alias or -1
    int nV, nVtot, nBnd;
};
struct Slots5 { int s[5]; };

// one triangle's 64-byte record: four 16-byte loads, two sectors (the SoA arrays cost eight sectors per visit)
struct ElemRecord { int i0, i1, i2; double area, A2, e0, e1, d; };
__device__ __forceinline__ ElemRecord load_record(const double* __restrict__ rec, int t)
{
    const int4 iv = __ldg(reinterpret_cast<const int4*>(rec) + 4 * (size_t)t);
    const double2 a = __ldg(reinterpret_cast<const double2*>(rec) + 4 * (size_t)t + 1), b = __ldg(reinterpret_cast<const double2*>(rec) + 4 * (size_t)t + 2);
    const double2 cc = __ldg(reinterpret_cast<const double2*>(rec) + 4 * (size_t)t + 3);
    ElemRecord R; R.i0 = iv.x; R.i1 = iv.y; R.i2 = iv.z; R.area = a.x; R.A2 = a.y; R.e0 = b.x; R.e1 = b.y; R.d = cc.x;
    return R;
}

static constexpr int kGradLanes = 4;        // lanes per vertex: the corner evaluations (7 fp64 divisions each) of a vertex run side by side

// The kGradLanes lanes of a vertex evaluate its corners round-robin (corner j on lane j % kGradLanes); the sum is then
// formed in corner order from shuffled values -- every lane ends up with the same, order-exact sum -- so the gradient
// is bit-identical to a serial walk.  Element values / inversion counts are accumulated by the lane that evaluated
// corner 0 of the triangle.
template <bool AIR_SET>
__device__ __forceinline__ void gather_corners(const ElemView& S, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, int row,
                                               const double* __restrict__ x, int lane, unsigned groupMask, int groupBase,
                                               Vec2& gsum, double& Esum, double& nInv)
{
    const int q0 = row >= 0 ? ptr[row] : 0, q1 = row >= 0 ? ptr[row + 1] : 0;
    // the trip count is uniform over the group (same row) but not over the warp: shuffles are masked per group
    for (int base = q0; base < q1; base += kGradLanes) {
        const int q = base + lane;
        Vec2 gk = mk(0.0, 0.0);
        if (q < q1) {
            const int code = __ldg(idx + q), t = code >> 2, k = code & 3;
            const ElemRecord R = load_record(S.rec, t);
            const double w = AIR_SET ? 1.0 : R.area / S.surfaceArea;
            double E, dbArea;
            sd_corner(ld2(x, R.i0), ld2(x, R.i1), ld2(x, R.i2), R.A2, R.e0, R.e1, R.d, w, k, gk, E, dbArea);
            if (k == 0) { Esum += E; if (dbArea < 0.0) nInv += 1.0; }
        }
#pragma unroll
        for (int l = 0; l < kGradLanes; ++l) {
            const double gx = __shfl_sync(groupMask, gk.x, groupBase + l), gy = __shfl_sync(groupMask, gk.y, groupBase + l);
            if (base + l < q1) { gsum.x += gx; gsum.y += gy; }
        }
    }
}

__global__ void __launch_bounds__(kBlock, 3)
grad_gather_kernel(ElemView M, ElemView A, GatherView G, const double* __restrict__ x, const uint8_t* __restrict__ fixedMask,
                   double* __restrict__ g, double* __restrict__ partials, unsigned* __restrict__ ticket,
                   double* __restrict__ scal, Slots5 slots)
{
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // E mesh, E air, #inverted, ||g||^2, ||g_mesh||^2 (unscaled mesh term)
    double2* g2 = reinterpret_cast<double2*>(g);
    const int lane = threadIdx.x & (kGradLanes - 1);
    const int groupBase = (threadIdx.x & 31) & ~(kGradLanes - 1);
    const unsigned groupMask = ((1u << kGradLanes) - 1u) << groupBase;
    const long nGroups = (long)gridDim.x * kBlock / kGradLanes;
    for (long w = (blockIdx.x * (long)kBlock + threadIdx.x) / kGradLanes; w < G.nVtot; w += nGroups) {
        const int v = (int)w;
        Vec2 gm = mk(0.0, 0.0), ga = mk(0.0, 0.0);
        int la = -1;
        if (v < G.nV) { if (A.n > 0) la = __ldg(G.g2l + v); }
        else la = G.nBnd + (v - G.nV);
        gather_corners<false>(M, G.vcPtrM, G.vcIdxM, v < G.nV ? v : -1, x, lane, groupMask, groupBase, gm, acc[0], acc[2]);
        gather_corners<true>(A, G.vcPtrA, G.vcIdxA, la, x, lane, groupMask, groupBase, ga, acc[1], acc[2]);
        if (lane != 0) continue;
        const unsigned fx = fixedMask[v];
        // per-term masking like the reference: the mesh term zeroes the mesh's fixed vertices, the air term the air mesh's
        if (fx & 1u) gm = mk(0.0, 0.0);
        if (fx & 2u) ga = mk(0.0, 0.0);
        Vec2 gv = mk(M.scale * gm.x, M.scale * gm.y);
        if (la >= 0) { gv.x += A.scale * ga.x; gv.y += A.scale * ga.y; }
        if (fx) gv = mk(0.0, 0.0);                // the merged fixed set (Scaffold::mergeFixedV) has no free DOF here
        g2[v] = make_double2(gv.x, gv.y);
        acc[3] += gv.x * gv.x; acc[3] += gv.y * gv.y;
        acc[4] += gm.x * gm.x; acc[4] += gm.y * gm.y;
    }
    reduce_finalize<5, false>(acc, partials, ticket, scal, slots.s);
}

// plain ||v||^2 of a system vector (solver-side helpers)
__global__ void __launch_bounds__(kBlock)
sqnorm_kernel(const double* __restrict__ v, int n, double* __restrict__ partials,
              unsigned* __restrict__ ticket, double* __restrict__ scal, Slots1 slots)
{
    double acc[1] = {0.0};
    const double2* v2 = reinterpret_cast<const double2*>(v);
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < n / 2; i += gridDim.x * kBlock) {
        const double2 t = v2[i];
        acc[0] += t.x * t.x;
        acc[0] += t.y * t.y;
    }
    reduce_finalize<1, false>(acc, partials, ticket, scal, slots.s);
}

// vertex -> corner incidence of the air mesh is rebuilt with every new air mesh: alias table of the mesh boundary
__global__ void __launch_bounds__(kBlock)
g2l_set_kernel(int nBnd, const int32_t* __restrict__ l2g, int32_t* __restrict__ g2l)
{
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < nBnd; i += gridDim.x * kBlock) g2l[l2g[i]] = i;
}

// ---------------------------------------------------------------------------------------------
// a10: element -> BSR block slot map.  slot[(3k+l) * n + t] = index of block (v_k, v_l) or -1.
__global__ void __launch_bounds__(kBlock)
build_slots_kernel(int n, const int32_t* __restrict__ v0, const int32_t* __restrict__ v1, const int32_t* __restrict__ v2,
                   const uint8_t* __restrict__ fixedMask, const int32_t* __restrict__ rowPtr,
                   const int32_t* __restrict__ colIdx, int32_t* __restrict__ slot, int* __restrict__ missing,
                   const int32_t* __restrict__ rowOf)
{
    for (int t = blockIdx.x * kBlock + threadIdx.x; t < n; t += gridDim.x * kBlock) {
        const int idx[3] = {v0[t], v1[t], v2[t]};
        const int rw[3] = {rowOf[idx[0]], rowOf[idx[1]], rowOf[idx[2]]};     // the BSR lives in the solver's row order
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                int s = -1;
                if (!fixedMask[idx[k]] && !fixedMask[idx[l]]) {
                    // rows are ~7 blocks long and NOT sorted (the host skips the sort): linear scan
                    const int lo = rowPtr[rw[k]], hi = rowPtr[rw[k] + 1];
                    const int col = rw[l];
                    for (int b = lo; b < hi; ++b) if (colIdx[b] == col) { s = b; break; }
                    if (s < 0) atomicAdd(missing, 1);
                }
                slot[(size_t)(3 * k + l) * n + t] = s;
            }
    }
}

// ---------------------------------------------------------------------------------------------
// a5/a9/a11/a18: per-element Hessian + PSD projection, then a deterministic row gather into the BSR values.
//
// Pass 1 (hessian_elem_kernel, one thread per element): the 6 upper 2x2 blocks live in registers; after the projection
// they are scaled (energyParams[0] for the mesh, w_scaf/|Fa| for the air mesh: Optimizer.cpp:821-832, Scaffold.cpp:231-248)
// and stored block-major, hel[b][e] = one 32-byte sector per block, so a warp writes 1 KB runs.
// Pass 2 (hessian_rows_kernel, one thread per block row = vertex): walks the vertex's incident corners in ascending
// element order (mesh, then air) and adds row k of each element block matrix into the row's BSR blocks, accumulating in
// shared memory; every value of the matrix is therefore the sum, in the reference's triplet order
// (LinSysSolver::update_a, LinSysSolver.hpp:147-157), of the same addends: no atomics, no memset, bit-reproducible.
// Fixed vertices get their identity row here (addDiagonalToMatrix, SymDirichletEnergy.cpp:541-548).
template <bool TO_HEL>
struct HessQueue {
    int32_t qi[2][3][kBlock];
    double qd[2][5][kBlock];
    __device__ __forceinline__ void issue(const ElemView& M, const ElemView& A, int e, int total, int st) {
        if (e < total) {
            const bool isAir = e >= M.n;
            const ElemView& S = isAir ? A : M;
            const int t = isAir ? e - M.n : e;
            const int tid = threadIdx.x;
            cp_async4(&qi[st][0][tid], S.v0 + t); cp_async4(&qi[st][1][tid], S.v1 + t); cp_async4(&qi[st][2][tid], S.v2 + t);
            cp_async8(&qd[st][0][tid], S.area + t); cp_async8(&qd[st][1][tid], S.areaSq + t);
            cp_async8(&qd[st][2][tid], S.k0 + t); cp_async8(&qd[st][3][tid], S.k1 + t); cp_async8(&qd[st][4][tid], S.kd + t);
        }
        cp_commit();
    }
};

template <bool TO_HEL>
__global__ void __launch_bounds__(kBlock, 2)
hessian_elem_kernel(ElemView M, ElemView A, const double* __restrict__ x, double* __restrict__ hel,
                    double* __restrict__ out36)
{
    __shared__ HessQueue<TO_HEL> Q;
    const int total = M.n + (TO_HEL ? A.n : 0), stride = gridDim.x * kBlock, tid = threadIdx.x;
    int e = blockIdx.x * kBlock + tid, st = 0;
    Q.issue(M, A, e, total, 0);
    for (; e < total; e += stride, st ^= 1) {
        Q.issue(M, A, e + stride, total, st ^ 1);
        cp_wait<1>();
        const bool isAir = e >= M.n;
        const ElemView& S = isAir ? A : M;
        const int t = isAir ? e - M.n : e;
        const Vec2 U1 = ld2(x, Q.qi[st][0][tid]), U2 = ld2(x, Q.qi[st][1][tid]), U3 = ld2(x, Q.qi[st][2][tid]);
        const double w = S.uniform ? 1.0 : Q.qd[st][0][tid] / S.surfaceArea;
        double Hb[6][2][2];
        sd_hessian(U1, U2, U3, Q.qd[st][1][tid], Q.qd[st][2][tid], Q.qd[st][3][tid], Q.qd[st][4][tid], w, Hb);
        sd_project_psd(Hb);
        if (TO_HEL) {
            const double sc = S.scale;
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                double2* dst = reinterpret_cast<double2*>(hel + 4 * ((size_t)b * total + e));
                dst[0] = make_double2(sc * Hb[b][0][0], sc * Hb[b][0][1]);
                dst[1] = make_double2(sc * Hb[b][1][0], sc * Hb[b][1][1]);
            }
        } else {
            // dense 6x6, row-major, for parity tests against makePD
            double* o = out36 + 36 * (size_t)t;
            const int bOf[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int l = 0; l < 3; ++l)
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            o[(2 * k + i) * 6 + (2 * l + j)] = (k <= l) ? Hb[bOf[k][l]][i][j] : Hb[bOf[k][l]][j][i];
        }
    }
}

static constexpr int kRowLanes = 8;         // lanes per block row of the row gather (rows are ~7 blocks long)

// Pass 2, one group of 8 lanes per block row (= vertex): lane q owns the row's BSR blocks q, q + 8, ...  Every lane
// walks ALL incident corners of the vertex in ascending element order (the group reads the same addresses: broadcasts)
// and adds row k of an element's block matrix into its own block when the element's slot map says so -- a register
// accumulator per owned block, contributions in the reference's triplet order, one 32-byte store per block.
__device__ __forceinline__ void row_lane_add(const double* __restrict__ hel, size_t nE, size_t eBase, const int4 cs, int myBlock, double (&acc)[4])
{
    // cs = {element << 2 | corner, slot(k,0), slot(k,1), slot(k,2)}: one 16-byte load per corner, read front to back
    const int k = cs.x & 3;
    const int l = (cs.y == myBlock) ? 0 : ((cs.z == myBlock) ? 1 : ((cs.w == myBlock) ? 2 : -1));
    if (l < 0) return;
    // block (k,l) of the element: stored as is for k <= l, as the transpose of (l,k) otherwise
    const int lo = k < l ? k : l, hi = k < l ? l : k;
    const int b = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);                 // (0,0)(0,1)(0,2)(1,1)(1,2)(2,2) -> 0..5
    const double2* src = reinterpret_cast<const double2*>(hel + 4 * ((size_t)b * nE + eBase + (size_t)(cs.x >> 2)));
    const double2 r0 = __ldcg(src), r1 = __ldcg(src + 1);
    const bool tr = k > l;
    acc[0] += r0.x; acc[1] += tr ? r1.x : r0.y; acc[2] += tr ? r0.y : r1.x; acc[3] += r1.y;
}

// per corner of the vertex->corner list: the corner code and the three BSR slots of its row of the element block matrix
__global__ void __launch_bounds__(kBlock)
build_vcslot_kernel(int nCorners, int n, const int32_t* __restrict__ vcIdx, const int32_t* __restrict__ slot, int32_t* __restrict__ vcSlot)
{
    for (int q = blockIdx.x * kBlock + threadIdx.x; q < nCorners; q += gridDim.x * kBlock) {
        const int code = vcIdx[q], t = code >> 2, k = code & 3;
        reinterpret_cast<int4*>(vcSlot)[q] = make_int4(code, slot[(size_t)(3 * k) * n + t], slot[(size_t)(3 * k + 1) * n + t], slot[(size_t)(3 * k + 2) * n + t]);
    }
}

__global__ void __launch_bounds__(kBlock)
hessian_rows_kernel(ElemView M, ElemView A, GatherView G, const double* __restrict__ hel, const uint8_t* __restrict__ fixedMask,
                    const int32_t* __restrict__ rowOf, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx,
                    double* __restrict__ val)
{
    const size_t nE = (size_t)M.n + A.n;
    const int lane = threadIdx.x & (kRowLanes - 1);
    for (long w = (blockIdx.x * (long)kBlock + threadIdx.x) / kRowLanes; w < G.nVtot; w += (long)gridDim.x * kBlock / kRowLanes) {
        const int v = (int)w;
        const int r = rowOf[v], lo = rowPtr[r], nb = rowPtr[r + 1] - lo;
        const unsigned fx = fixedMask[v];
        if (fx) {
            // identity scaled like every other triplet of its term (mask bit 0: fixed by the mesh, bit 1: by the air mesh)
            const double dgn = ((fx & 1u) ? M.scale : 0.0) + ((fx & 2u) ? A.scale : 0.0);
            for (int b = lo + lane; b < lo + nb; b += kRowLanes)
                if (colIdx[b] == r) { double2* o = reinterpret_cast<double2*>(val + 4 * (size_t)b); o[0] = make_double2(dgn, 0.0); o[1] = make_double2(0.0, dgn); }
            continue;
        }
        const int la = (v < G.nV) ? (A.n > 0 ? __ldg(G.g2l + v) : -1) : G.nBnd + (v - G.nV);
        const int qM0 = v < G.nV ? G.vcPtrM[v] : 0, qM1 = v < G.nV ? G.vcPtrM[v + 1] : 0;
        const int qA0 = la >= 0 ? G.vcPtrA[la] : 0, qA1 = la >= 0 ? G.vcPtrA[la + 1] : 0;
        for (int b = lo + lane; b < lo + nb; b += kRowLanes) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int q = qM0; q < qM1; ++q) row_lane_add(hel, nE, 0, __ldg(reinterpret_cast<const int4*>(M.vcSlot) + q), b, acc);
            for (int q = qA0; q < qA1; ++q) row_lane_add(hel, nE, (size_t)M.n, __ldg(reinterpret_cast<const int4*>(A.vcSlot) + q), b, acc);
            double2* o = reinterpret_cast<double2*>(val + 4 * (size_t)b);
            o[0] = make_double2(acc[0], acc[1]); o[1] = make_double2(acc[2], acc[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a7: step bound (min over all mesh + air elements)
__global__ void __launch_bounds__(kBlock)
step_bound_kernel(ElemView M, ElemView A, const double* __restrict__ x, const double* __restrict__ dir,
                  double alpha0, double* __restrict__ partials, unsigned* __restrict__ ticket,
                  double* __restrict__ scal, Slots1 slots)
{
    __shared__ ElemQueue<0> Q;
    double acc[1] = {alpha0};
    const int total = M.n + A.n, stride = gridDim.x * kBlock, tid = threadIdx.x;
    int e = blockIdx.x * kBlock + tid, st = 0;
#pragma unroll
    for (int k = 0; k < kDepth - 1; ++k) Q.issue(M, A, e + k * stride, total, k);
    for (; e < total; e += stride, st = (st + 1 == kDepth) ? 0 : st + 1) {
        Q.issue(M, A, e + (kDepth - 1) * stride, total, (st + kDepth - 1) % kDepth);
        cp_wait<kDepth - 1>();
        const int i0 = Q.qi[st][0][tid], i1 = Q.qi[st][1][tid], i2 = Q.qi[st][2][tid];
        acc[0] = sd_step_bound(ld2(x, i0), ld2(x, i1), ld2(x, i2), ld2(dir, i0), ld2(dir, i1), ld2(dir, i2), acc[0]);
    }
    // the reference compares every bound against the running (alpha0-initialised) minimum
    reduce_finalize<1, true>(acc, partials, ticket, scal, slots.s);
}

// a14: x = x0 + alpha p  (Optimizer::stepForward + Scaffold::stepForward)
__global__ void __launch_bounds__(kBlock)
step_forward_kernel(int n, const double* __restrict__ x0, const double* __restrict__ p, double alpha, double* __restrict__ x)
{
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock)
        x[i] = __dadd_rn(x0[i], __dmul_rn(alpha, p[i]));
}

// ---------------------------------------------------------------------------------------------
// LinSysSolver::update_a mirror: triplets (i<=j kept, mirrored into the full BSR)
__global__ void __launch_bounds__(kBlock)
triplet_scatter_kernel(long nT, const int32_t* __restrict__ I, const int32_t* __restrict__ J, const double* __restrict__ S,
                       const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, double* __restrict__ val,
                       int* __restrict__ missing, const int32_t* __restrict__ perm)
{
    for (long k = blockIdx.x * (long)kBlock + threadIdx.x; k < nT; k += (long)gridDim.x * kBlock) {
        const int i = I[k], j = J[k];
        if (i > j) continue;                    // the reference keeps the upper triangle of the CALLER's numbering
        const int bi = perm[i >> 1], bj = perm[j >> 1], ri = i & 1, rj = j & 1;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int row = pass ? bj : bi, col = pass ? bi : bj, rr = pass ? rj : ri, cc = pass ? ri : rj;
            if (pass && i == j) break;     // diagonal scalar entry: once
            int s = -1;
            for (int b = rowPtr[row], hi = rowPtr[row + 1]; b < hi; ++b) if (colIdx[b] == col) { s = b; break; }
            if (s < 0) { atomicAdd(missing, 1); continue; }
            atomicAdd(&val[4 * (size_t)s + 2 * rr + cc], S[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// UV layout conversion: Eigen column-major (all u, then all v) <-> interleaved system vector
__global__ void __launch_bounds__(kBlock)
set_uv_kernel(int nV, const double* __restrict__ V, int nVa, int nBnd, const double* __restrict__ Va, double* __restrict__ x,
              const int32_t* __restrict__ perm)
{
    // perm: user vertex id -> internal (locality-ordered) id; air interior vertices keep nV + k
    const int total = (V ? nV : 0) + (Va ? (nVa - nBnd) : 0);
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < total; i += gridDim.x * kBlock) {
        int k = i;
        if (V) {
            if (k < nV) { const int q = perm[k]; x[2 * q] = V[k]; x[2 * q + 1] = V[nV + k]; continue; }
            k -= nV;
        }
        const int a = nBnd + k;   // interior air vertex
        x[2 * (nV + k)] = Va[a]; x[2 * (nV + k) + 1] = Va[nVa + a];
    }
}
__global__ void __launch_bounds__(kBlock)
get_uv_kernel(int nV, double* __restrict__ V, int nVa, const int32_t* __restrict__ l2g, double* __restrict__ Va, const double* __restrict__ x,
              const int32_t* __restrict__ perm)
{
    const int total = (V ? nV : 0) + (Va ? nVa : 0);
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < total; i += gridDim.x * kBlock) {
        int k = i;
        if (V) {
            if (k < nV) { const int q = perm[k]; V[k] = x[2 * q]; V[nV + k] = x[2 * q + 1]; continue; }
            k -= nV;
        }
        const int gidx = l2g[k];
        Va[k] = x[2 * gidx]; Va[nVa + k] = x[2 * gidx + 1];
    }
}

// system vectors cross the API in the caller's vertex order; on the device they live in the internal order
__global__ void __launch_bounds__(kBlock)
permute_vec_kernel(int nVtot, const int32_t* __restrict__ perm, const double* __restrict__ in, double* __restrict__ out, int toInternal)
{
    const double2* in2 = reinterpret_cast<const double2*>(in);
    double2* out2 = reinterpret_cast<double2*>(out);
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < nVtot; k += gridDim.x * kBlock) {
        const int q = perm[k];
        if (toInternal) out2[q] = in2[k]; else out2[k] = in2[q];
    }
}
__global__ void __launch_bounds__(kBlock)
permute_scalar_kernel(int n, const int32_t* __restrict__ perm, const double* __restrict__ in, double* __restrict__ out)
{
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < n; k += gridDim.x * kBlock) out[k] = in[perm[k]];
}

// ---------------------------------------------------------------------------------------------
// a1: rest-frame features (TriMesh::computeFeatures arithmetic, TriMesh.cpp:355-398)
__global__ void __launch_bounds__(kBlock)
rest_features_kernel(int nV, int nF, const double* __restrict__ P, const int32_t* __restrict__ F, double thres,
                     double* __restrict__ rest8, double* __restrict__ partials, unsigned* __restrict__ ticket,
                     double* __restrict__ scal, Slots3 slots)
{
    double acc[3] = {0.0, 0.0, 0.0};   // surface area, sum of edge lengths, #zero-area triangles
    const double sqrt3 = sqrt(3.0);
    for (int t = blockIdx.x * kBlock + threadIdx.x; t < nF; t += gridDim.x * kBlock) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const double ax = P[i1] - P[i0], ay = P[nV + i1] - P[nV + i0], az = P[2 * nV + i1] - P[2 * nV + i0];
        const double bx = P[i2] - P[i0], by = P[nV + i2] - P[nV + i0], bz = P[2 * nV + i2] - P[2 * nV + i0];
        const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
        double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
        if (area == 0.0) acc[2] += 1.0;
        double A2, e0, e1, d, k0, k1, kd;
        if (area < thres) {
            area = thres; A2 = thres * thres;
            e0 = e1 = 4.0 / sqrt3 * thres; d = e0 / 2.0;
            k0 = k1 = 2.0 / sqrt3 / thres; kd = k0 / 2.0;
        } else {
            A2 = area * area;
            e0 = ax * ax + ay * ay + az * az; e1 = bx * bx + by * by + bz * bz; d = ax * bx + ay * by + az * bz;
            k0 = e0 / 2. / A2; k1 = e1 / 2. / A2; kd = d / 2. / A2;
        }
        acc[0] += area;
        // igl::avg_edge_length: mean over the 3|F| triangle edges
        const double ex = bx - ax, ey = by - ay, ez = bz - az;
        acc[1] += sqrt(ax * ax + ay * ay + az * az) + sqrt(bx * bx + by * by + bz * bz) + sqrt(ex * ex + ey * ey + ez * ez);
        rest8[t] = area; rest8[(size_t)nF + t] = A2; rest8[2 * (size_t)nF + t] = e0; rest8[3 * (size_t)nF + t] = e1;
        rest8[4 * (size_t)nF + t] = d; rest8[5 * (size_t)nF + t] = k0; rest8[6 * (size_t)nF + t] = k1; rest8[7 * (size_t)nF + t] = kd;
    }
    reduce_finalize<3, false>(acc, partials, ticket, scal, slots.s);
}

// ---------------------------------------------------------------------------------------------
// a15: seam length (TriMesh::computeSeamSparsity)
__global__ void __launch_bounds__(kBlock)
seam_kernel(int nCoh, const int32_t* __restrict__ coh, const double* __restrict__ len, const int32_t* __restrict__ bnd,
            const double* __restrict__ x, double avgEdgeLen, int triSoup,
            double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ scal, Slots1 slots)
{
    double acc[1] = {0.0};
    const double thres = 1.0e-2;
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < nCoh; c += gridDim.x * kBlock) {
        if (bnd[c]) continue;
        bool take = !triSoup;
        if (!take) {
            const Vec2 a = ld2(x, coh[c]) - ld2(x, coh[2 * nCoh + c]);
            const Vec2 b = ld2(x, coh[nCoh + c]) - ld2(x, coh[3 * nCoh + c]);
            take = (sqrt(dot(a, a)) / avgEdgeLen > thres) || (sqrt(dot(b, b)) / avgEdgeLen > thres);
        }
        if (take) acc[0] += len[c];
    }
    reduce_finalize<1, false>(acc, partials, ticket, scal, slots.s);
}

// ---------------------------------------------------------------------------------------------
// a8: per-vertex sample standard deviation of the incident corner gradients (mesh term, area weights):
// SymDirichletEnergy::computeLocalGradient (:215-256) + computeDivGradPerVert (:108-149).  One thread per vertex, two
// walks over its corners in ascending triangle order (mean, then squared deviations): the reference's summation
// order, no atomics, so the candidate ordering it feeds (TriMesh.cpp:562-596) is reproducible to the bit.
__global__ void __launch_bounds__(kBlock)
divgrad_gather_kernel(ElemView M, GatherView G, const double* __restrict__ x, double* __restrict__ out)
{
    for (int v = blockIdx.x * kBlock + threadIdx.x; v < G.nV; v += gridDim.x * kBlock) {
        const int q0 = G.vcPtrM[v], q1 = G.vcPtrM[v + 1], n = q1 - q0;
        double mx = 0.0, my = 0.0, dev = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            for (int q = q0; q < q1; ++q) {
                const int code = __ldg(G.vcIdxM + q), t = code >> 2, k = code & 3;
                const ElemRecord R = load_record(M.rec, t);
                const double w = R.area / M.surfaceArea;
                Vec2 gk; double E, dbArea;
                sd_corner(ld2(x, R.i0), ld2(x, R.i1), ld2(x, R.i2), R.A2, R.e0, R.e1, R.d, w, k, gk, E, dbArea);
                if (pass == 0) { mx += gk.x; my += gk.y; }
                else { const double dx = gk.x - mx, dy = gk.y - my; dev += dx * dx + dy * dy; }
            }
            if (pass == 0) { mx /= n; my /= n; }
        }
        out[v] = (n <= 1) ? 0.0 : sqrt(dev / (n - 1.0));
    }
}

// ---------------------------------------------------------------------------------------------
// a6: dense Hessian of the mesh term (SymDirichletEnergy::computeHessian, dense flavour, :306-427): the thread of vertex v adds
// row k of every incident element's projected block matrix into dense rows 2v, 2v + 1 in ascending triangle order (the
// order of the reference's serial addBlockToMatrix loop); fixed vertices: zero row / column, unit diagonal.  `out` is
// pre-zeroed, n = 2 nV, indexed by the CALLER's vertex ids (inv: internal -> caller).
__global__ void __launch_bounds__(kBlock)
dense_hessian_kernel(ElemView M, GatherView G, const double* __restrict__ blocks36, const uint8_t* __restrict__ fixedMask,
                     const int32_t* __restrict__ inv, double* __restrict__ out)
{
    const size_t n = 2 * (size_t)G.nV;
    for (int v = blockIdx.x * kBlock + threadIdx.x; v < G.nV; v += gridDim.x * kBlock) {
        const size_t r = 2 * (size_t)inv[v];
        if (fixedMask[v]) { out[r * n + r] = 1.0; out[(r + 1) * n + r + 1] = 1.0; continue; }
        for (int q = G.vcPtrM[v]; q < G.vcPtrM[v + 1]; ++q) {
            const int code = G.vcIdxM[q], t = code >> 2, k = code & 3;
            const int idx[3] = {M.v0[t], M.v1[t], M.v2[t]};
            const double* B = blocks36 + 36 * (size_t)t;
            for (int l = 0; l < 3; ++l) {
                if (fixedMask[idx[l]]) continue;
                const size_t c = 2 * (size_t)inv[idx[l]];
                for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) out[(r + i) * n + c + j] += B[(2 * k + i) * 6 + 2 * l + j];
            }
        }
    }
}

// =============================================================================================
// launchers
#define KCHECK(c) do { (c)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return cuda_fail((c), _e, __func__); } while (0)

// grid of the queue kernels: one wave of resident CTAs (occupancy queried once per kernel), so that every CTA
// runs the prefetch queue over the same number of rounds and none starts behind another
template <auto Kern>
static int resident_grid(const ocb_ctx* c, long n) {
    static int perSM = 0;
    if (!perSM) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, Kern, kBlock, 0) != cudaSuccess || b < 1) b = 1;
        perSM = b;
    }
    return grid_for(c, n, perSM);
}

static int ensure_reduce_bufs(ocb_ctx* c, int grid, int nv) {
    OCB_CUDA(c, c->partials.reserve((size_t)grid * nv + 64, c->stream));
    return 0;
}

int launch_energy(ocb_ctx* c, double p0, bool stepped, double alpha, bool alphaFromStepBound)
{
    ProfScope prof(c, K_ENERGY);
    const ElemView M = view_of(c, c->mesh, false, p0, 0), A = view_of(c, c->air, true, c->wScafOverFa, 1);
    const int grid = stepped ? resident_grid<energy_kernel<true>>(c, (long)M.n + A.n) : resident_grid<energy_kernel<false>>(c, (long)M.n + A.n);
    OCB_TRY(ensure_reduce_bufs(c, grid, 3));
    Slots3 sl; sl.s[0] = S_E_MESH; sl.s[1] = S_E_AIR; sl.s[2] = S_N_INVERTED;
    if (stepped) energy_kernel<true><<<grid, kBlock, 0, c->stream>>>(M, A, c->x0.p, c->p.p, alpha, c->partials.p, c->sync.p, c->dScal, sl, alphaFromStepBound ? 1 : 0);
    else energy_kernel<false><<<grid, kBlock, 0, c->stream>>>(M, A, c->x.p, nullptr, 0.0, c->partials.p, c->sync.p, c->dScal, sl, 0);
    KCHECK(c);
    return 0;
}

int launch_energy_one_elem(ocb_ctx* c, int t, int uniform, double* d_out)
{
    ProfScope prof(c, K_ENERGY);
    ElemView M = view_of(c, c->mesh, false, 1.0, uniform);
    M.v0 += t; M.v1 += t; M.v2 += t; M.area += t; M.areaSq += t; M.e0 += t; M.e1 += t; M.d += t; M.n = 1;     // a one-triangle view
    energy_per_elem_kernel<<<1, kBlock, 0, c->stream>>>(M, c->x.p, d_out);
    KCHECK(c);
    return 0;
}

int launch_energy_per_elem(ocb_ctx* c, int uniform, double* d_out)
{
    ProfScope prof(c, K_ENERGY);
    const ElemView M = view_of(c, c->mesh, false, 1.0, uniform);
    energy_per_elem_kernel<<<grid_for(c, M.n), kBlock, 0, c->stream>>>(M, c->x.p, d_out);
    KCHECK(c);
    return 0;
}

int launch_sqnorm(ocb_ctx* c, const double* v, int n, int slot)
{
    if (n & 1) return set_err(c, OCB_ERR_ARG, "sqnorm: system vectors hold two entries per vertex");
    const int grid = grid_for(c, n / 2, 4);
    OCB_TRY(ensure_reduce_bufs(c, grid, 1));
    Slots1 sl; sl.s[0] = slot;
    sqnorm_kernel<<<grid, kBlock, 0, c->stream>>>(v, n, c->partials.p, c->sync.p, c->dScal, sl);
    KCHECK(c);
    return 0;
}

static GatherView gather_of(const ocb_ctx* c)
{
    GatherView G;
    G.vcPtrM = c->vcPtrM.p; G.vcIdxM = c->vcIdxM.p; G.vcPtrA = c->vcPtrA.p; G.vcIdxA = c->vcIdxA.p; G.g2l = c->g2l.p;
    G.nV = c->nV; G.nVtot = c->nVtot; G.nBnd = c->nBnd;
    return G;
}

// gradient + energy at x + ||g||^2 in one launch (scalars: S_E_MESH, S_E_AIR, S_N_INVERTED, S_SQN_G, S_SQN_G_MESH)
int launch_gradient(ocb_ctx* c, double p0)
{
    ProfScope prof(c, K_GRADIENT);
    const ElemView M = view_of(c, c->mesh, false, p0, 0), A = view_of(c, c->air, true, c->wScafOverFa, 1);
    const int grid = grid_for(c, (long)c->nVtot * kGradLanes, 8);
    OCB_TRY(ensure_reduce_bufs(c, grid, 5));
    Slots5 sl; sl.s[0] = S_E_MESH; sl.s[1] = S_E_AIR; sl.s[2] = S_N_INVERTED; sl.s[3] = S_SQN_G; sl.s[4] = S_SQN_G_MESH;
    grad_gather_kernel<<<grid, kBlock, 0, c->stream>>>(M, A, gather_of(c), c->x.p, c->fixedMask.p, c->g.p, c->partials.p, c->sync.p, c->dScal, sl);
    KCHECK(c);
    return 0;
}

int launch_g2l(ocb_ctx* c)
{
    OCB_CUDA(c, c->g2l.reserve((size_t)c->nV + 1, c->stream));
    OCB_CUDA(c, cudaMemsetAsync(c->g2l.p, 0xFF, sizeof(int32_t) * (size_t)c->nV, c->stream));
    if (c->nBnd > 0) {
        g2l_set_kernel<<<grid_for(c, c->nBnd, 1), kBlock, 0, c->stream>>>(c->nBnd, c->l2g.p, c->g2l.p);
        KCHECK(c);
    }
    return 0;
}

int launch_build_slots(ocb_ctx* c)
{
    int* missing = reinterpret_cast<int*>(c->sync.p + 8);
    OCB_CUDA(c, cudaMemsetAsync(missing, 0, sizeof(int), c->stream));
    for (int which = 0; which < 2; ++which) {
        ElemSet& S = which ? c->air : c->mesh;
        if (S.n == 0) continue;
        OCB_CUDA(c, S.slot.reserve((size_t)9 * S.n, c->stream));
        build_slots_kernel<<<grid_for(c, S.n), kBlock, 0, c->stream>>>(S.n, S.v.p, S.v.p + S.n, S.v.p + 2 * (size_t)S.n,
                                                                      c->fixedMask.p, c->rowPtr.p, c->colIdx.p, S.slot.p, missing, c->rowOf.p);
        KCHECK(c);
        OCB_CUDA(c, S.vcSlot.reserve((size_t)12 * S.n + 4, c->stream));
        build_vcslot_kernel<<<grid_for(c, 3L * S.n), kBlock, 0, c->stream>>>(3 * S.n, S.n, which ? c->vcIdxA.p : c->vcIdxM.p, S.slot.p, S.vcSlot.p);
        KCHECK(c);
    }
    int hMissing = 0;
    OCB_CUDA(c, cudaMemcpyAsync(&hMissing, missing, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (hMissing) return set_err(c, OCB_ERR_STATE, "pattern does not cover every element edge (adjacency inconsistent with the element lists)");
    c->slotsValid = true;
    return 0;
}

int launch_hessian(ocb_ctx* c, double p0)
{
    const ElemView M = view_of(c, c->mesh, false, p0, 0), A = view_of(c, c->air, true, c->wScafOverFa, 1);
    const size_t nE = (size_t)M.n + A.n;
    OCB_CUDA(c, c->hel.reserve(24 * nE + 4, c->stream));
    {
        ProfScope prof(c, K_HESSIAN);
        hessian_elem_kernel<true><<<resident_grid<hessian_elem_kernel<true>>(c, (long)nE), kBlock, 0, c->stream>>>(M, A, c->x.p, c->hel.p, nullptr);
        KCHECK(c);
    }
    ProfScope prof2(c, K_HESSIAN_ROWS);
    hessian_rows_kernel<<<grid_for(c, (long)c->nVtot * kRowLanes, 8), kBlock, 0, c->stream>>>(M, A, gather_of(c), c->hel.p, c->fixedMask.p, c->rowOf.p, c->rowPtr.p, c->colIdx.p, c->val.p);
    KCHECK(c);
    return 0;
}

int launch_hessian_blocks(ocb_ctx* c, int uniform, double* d_out36)
{
    ProfScope prof(c, K_HESSIAN);
    const ElemView M = view_of(c, c->mesh, false, 1.0, uniform);
    ElemView A = M; A.n = 0;
    hessian_elem_kernel<false><<<resident_grid<hessian_elem_kernel<false>>(c, M.n), kBlock, 0, c->stream>>>(M, A, c->x.p, nullptr, d_out36);
    KCHECK(c);
    return 0;
}

int launch_dense_hessian(ocb_ctx* c, const double* d_blocks36, const int32_t* d_inv, double* d_out)
{
    ProfScope prof(c, K_HESSIAN_ROWS);
    const ElemView M = view_of(c, c->mesh, false, 1.0, 0);
    dense_hessian_kernel<<<grid_for(c, c->nV, 8), kBlock, 0, c->stream>>>(M, gather_of(c), d_blocks36, c->fixedMask.p, d_inv, d_out);
    KCHECK(c);
    return 0;
}

int launch_step_bound(ocb_ctx* c, const double* d_dir, double alpha0)
{
    ProfScope prof(c, K_STEP_BOUND);
    const ElemView M = view_of(c, c->mesh, false, 1.0, 0), A = view_of(c, c->air, true, 1.0, 1);
    const int grid = resident_grid<step_bound_kernel>(c, (long)M.n + A.n);
    OCB_TRY(ensure_reduce_bufs(c, grid, 1));
    Slots1 sl; sl.s[0] = S_STEP_BOUND;
    step_bound_kernel<<<grid, kBlock, 0, c->stream>>>(M, A, c->x.p, d_dir, alpha0, c->partials.p, c->sync.p, c->dScal, sl);
    KCHECK(c);
    return 0;
}

int launch_step_forward(ocb_ctx* c, double alpha)
{
    ProfScope prof(c, K_STEP_FORWARD);
    step_forward_kernel<<<grid_for(c, c->nSys(), 4), kBlock, 0, c->stream>>>(c->nSys(), c->x0.p, c->p.p, alpha, c->x.p);
    KCHECK(c);
    return 0;
}

int launch_triplet_scatter(ocb_ctx* c, int64_t nT, const int32_t* dI, const int32_t* dJ, const double* dS)
{
    int* missing = reinterpret_cast<int*>(c->sync.p + 8);
    OCB_CUDA(c, cudaMemsetAsync(missing, 0, sizeof(int), c->stream));
    OCB_CUDA(c, cudaMemsetAsync(c->val.p, 0, sizeof(double) * 4 * (size_t)c->nnzb, c->stream));
    if (nT > 0) {
        triplet_scatter_kernel<<<grid_for(c, nT), kBlock, 0, c->stream>>>((long)nT, dI, dJ, dS, c->rowPtr.p, c->colIdx.p, c->val.p, missing, c->userRow.p);
        KCHECK(c);
    }
    int hMissing = 0;
    OCB_CUDA(c, cudaMemcpyAsync(&hMissing, missing, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (hMissing) return set_err(c, OCB_ERR_ARG, "triplet outside the sparsity pattern");
    return 0;
}

int launch_set_uv(ocb_ctx* c, const double* dV, const double* dVa)
{
    ProfScope prof(c, K_MISC);
    const int total = (dV ? c->nV : 0) + (dVa ? (c->nVa - c->nBnd) : 0);
    if (total <= 0) return 0;
    set_uv_kernel<<<grid_for(c, total, 4), kBlock, 0, c->stream>>>(c->nV, dV, c->nVa, c->nBnd, dVa, c->x.p, c->perm.p);
    KCHECK(c);
    return 0;
}
int launch_get_uv(ocb_ctx* c, double* dV, double* dVa)
{
    ProfScope prof(c, K_MISC);
    const int total = (dV ? c->nV : 0) + (dVa ? c->nVa : 0);
    if (total <= 0) return 0;
    get_uv_kernel<<<grid_for(c, total, 4), kBlock, 0, c->stream>>>(c->nV, dV, c->nVa, c->l2g.p, dVa, c->x.p, c->perm.p);
    KCHECK(c);
    return 0;
}

int launch_permute_vec(ocb_ctx* c, const double* in, double* out, bool toInternal)
{
    ProfScope prof(c, K_MISC);
    permute_vec_kernel<<<grid_for(c, c->nVtot, 4), kBlock, 0, c->stream>>>(c->nVtot, c->perm.p, in, out, toInternal ? 1 : 0);
    KCHECK(c);
    return 0;
}
int launch_permute_scalar(ocb_ctx* c, int n, const double* in, double* out)
{
    permute_scalar_kernel<<<grid_for(c, n, 4), kBlock, 0, c->stream>>>(n, c->perm.p, in, out);
    KCHECK(c);
    return 0;
}

int launch_rest_features(ocb_ctx* c, int nV, int nF, const double* dVrest, const int32_t* dF, double thres, double* dRest8)
{
    ProfScope prof(c, K_FEATURES);
    const int grid = grid_for(c, nF);
    OCB_TRY(ensure_reduce_bufs(c, grid, 3));
    Slots3 sl; sl.s[0] = S_MISC0; sl.s[1] = S_MISC1; sl.s[2] = S_MISC2;
    rest_features_kernel<<<grid, kBlock, 0, c->stream>>>(nV, nF, dVrest, dF, thres, dRest8, c->partials.p, c->sync.p, c->dScal, sl);
    KCHECK(c);
    return 0;
}

int launch_seam(ocb_ctx* c, int nCoh, const int32_t* dCoh, const double* dLen, const int32_t* dBnd, double avgEdgeLen, int triSoup)
{
    ProfScope prof(c, K_MISC);
    const int grid = grid_for(c, nCoh, 1);
    OCB_TRY(ensure_reduce_bufs(c, grid, 1));
    Slots1 sl; sl.s[0] = S_MISC0;
    seam_kernel<<<grid, kBlock, 0, c->stream>>>(nCoh, dCoh, dLen, dBnd, c->x.p, avgEdgeLen, triSoup, c->partials.p, c->sync.p, c->dScal, sl);
    KCHECK(c);
    return 0;
}

int launch_divgrad(ocb_ctx* c, double* d_out)
{
    ProfScope prof(c, K_MISC);
    const ElemView M = view_of(c, c->mesh, false, 1.0, 0);
    divgrad_gather_kernel<<<grid_for(c, c->nV, 8), kBlock, 0, c->stream>>>(M, gather_of(c), c->x.p, d_out);
    KCHECK(c);
    return 0;
}

}  // namespace ocb
