// optcuts_b200 — two-level additive Schwarz preconditioner (multilevel below the coarse level) of the PCG solve:
// device-side view and the CTA-level apply (included by ocb_pcg.cu; set up by ocb_mas.cu).
//
//   M^-1 = blockdiag2x2(A)^-1 + sum_{l=1..L-1} P_l D_l^-1 P_l^T + P_L (P_L^T A P_L)^-1 P_L^T
//
// The solver rows are ordered along a Hilbert curve through the UV positions so that a persistent CTA's
// contiguous row range is a compact patch, split into leaves of <= 8 vertices.  Level-l nodes carry 6 DOFs: the
// affine displacement fields (1, x, y) x (u, v) on the node, x/y in node-local coordinates.  8 consecutive nodes
// form a group (= a node of level l+1); below the coarse level L, D_l is the block diagonal of the Galerkin
// matrix P_l^T A P_l over the groups (48x48 blocks, inverted on the device, kept in fp32).  The coarse level L is
// the first level with at most kMasCoarseMax DOFs: its Galerkin matrix is inverted EXACTLY (dense blocked
// Gauss-Jordan in fp64 across the whole GPU, stored in fp32).  All levels live inside one CTA (groups never
// straddle a CTA); the only communication is the coarse residual (6 values per coarse node), which crosses CTAs
// through global memory at the grid barrier that also carries |r|^2.  Node tables and per-row basis values are
// staged in shared memory once per solve, and the CTA-local group solves run between the arrive and the wait
// of that barrier.  Measured on the reference's matrices (tools/mas_proto.py --exact-from): 2466 -> 174 CG
// iterations at 10k faces (Tutte state; 276 with block-diagonal coarse levels), 11 500 -> 286 at 160k (757).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ocb {

static constexpr int kMasLeaf = 8;          // vertices per leaf
static constexpr int kMasGroup = 8;         // nodes per group
static constexpr int kMasDof = 6;           // DOFs per node
static constexpr int kMasBlk = kMasGroup * kMasDof;   // 48
static constexpr int kMasMaxLevels = 12;

// Storage format of every inverse (group blocks and the coarse matrix): the HIGH 32 bits of the fp64 value, rounded to
// nearest -- 4 bytes per entry like fp32, a 20-bit mantissa (the CG iteration counts are those of fp32 storage:
// 158 / 93 / 286, tools/mas_proto.py --prec hi32) and NO conversion instruction on the way back: F2F.F64.F32 issues at
// ~4 per clock per SM and was the whole cost of the group and coarse solves (5.2k of 8.0k cycles at 10k faces, 29k
// cycles per CG iteration at 1M faces, profiles/r1e_pcg_phase_cycles.txt).  The buffers are typed float for the 16-byte
// vector loads; the bits are never interpreted as fp32.
#ifdef __CUDACC__
__device__ __forceinline__ float mas_pack(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v) + 0x80000000ull;
    return __uint_as_float((unsigned)(b >> 32));
}
__device__ __forceinline__ double mas_unpack(float w) { return __hiloint2double((int)__float_as_uint(w), 0); }
#endif

// one level of the hierarchy as the set-up kernels see it (global node indices)
struct MasLevel {
    int nNodes, nGroups;
    const int32_t* childBeg;   // nNodes + 1: rows (level 1) or nodes of the level below
    const int32_t* groupBeg;   // nGroups + 1: nodes of THIS level per group (= childBeg of the level above)
    const int32_t* parent;     // nNodes: group of the node
    const double4* geom;       // nNodes: {cx, cy, s, -}
    float* inv;                // nGroups x 48 x 48
};

static constexpr int kMasCoarseBlk = 48;    // tile of the dense coarse inversion
static constexpr int kMasCoarseMax = 3072;  // largest coarse system (DOFs)
// Largest coarse system for a solve with n block rows: the dense inverse costs ~0.37 us per DOF (sequential pivots) plus a
// cubic term (12 ms at 3072 DOFs), so small systems must not get a coarse level of thousands of DOFs (a 4096-vertex disk
// of the batch workload did: 15 ms per Newton step instead of 1.5).  n/8 keeps the set-up at ~10-20 % of the solve.
__host__ __device__ inline int mas_coarse_cap(int nRows)
{
    const int c = nRows / 8;
    return c < 600 ? 600 : (c > kMasCoarseMax ? kMasCoarseMax : c);
}

// Host-precomputed tables, CTA-local indices.  A CTA's local nodes are numbered level 1 first, then level 2, ...
// up to its coarse-level nodes.
struct MasView {
    int L;                     // coarse level; 0: preconditioner disabled (block-Jacobi only)
    int nC, ldC;               // coarse DOFs and the row stride of cinv (multiple of 48)
    int maxLocalNodes;         // max over CTAs of the nodes of levels 1..L
    int rowsPer;               // rows per CTA (vinfo staging)
    int maxOwnC;               // max over CTAs of the coarse nodes a CTA owns
    int cinvInSmem;            // 1: the CTA's rows of the coarse inverse are staged in shared memory for the whole solve
    const int32_t* ctaNodeOff; // grid + 1: first entry of every CTA in nodeA/nodeB/nodeX
    const int32_t* ctaSolve;   // grid: local nodes below level L (they take part in a CTA-local group solve)
    const int32_t* ctaLvOff;   // grid x (kMasMaxLevels + 1): first local node of level l at [l - 1]; [L] = all local nodes
    const int32_t* ctaLeafBeg; // grid + 1: first (global) leaf of every CTA
    const int32_t* ctaCBeg;    // grid + 1: first coarse node of every CTA
    const int4* nodeA;         // {local index of the group's first child, nk = 6 * children of the group, 6 * slot in the group, local index of the parent}
    const int4* nodeB;         // {float offset of the group's inverse, first child (local node; local row for leaves), children, level}
    const double4* nodeX;      // {tx, ty, rho, -}: transfer of the node's coefficients to its parent's basis
    const float* inv;          // group inverses of the levels below L (48 x 48 each, symmetric)
    const float* cinv;         // inverse of the coarse Galerkin matrix, nC rows of ldC floats
    const float4* vinfo;       // per row: {m, m*lx, m*ly, leaf id (int bits)}, m = 0 for fixed vertices
    double* rcC;               // nC: restricted residual on the coarse level (exchange buffer)
};

__host__ __device__ inline size_t mas_smem_bytes(int maxLocalNodes, int rowsPer, int ldC, int cinvRows = 0)
{
    size_t b = (size_t)cinvRows * ldC * 4;                         // the CTA's rows of the coarse inverse (optional)
    b += ((size_t)maxLocalNodes * kMasDof + kMasBlk) * 8 * 2;      // rc (+ zero pad), e
    b += (size_t)maxLocalNodes * (16 + 16 + 32);                   // nodeA, nodeB, nodeX
    b += (size_t)rowsPer * 16;                                     // vinfo
    b += (size_t)ldC * 8;                                          // coarse residual, all nodes
    b += (kMasMaxLevels + 2) * 4 + 64;                             // lvOff, slack
    return (b + 15) / 16 * 16;
}

#ifdef __CUDACC__
struct MasSmem {
    double* rc; double* e; double* rcAll; float* cinv;
    int4* nodeA; int4* nodeB; double4* nodeX; float4* vinfo;
    int* lvOff;
    int nLoc, nSolve, leaf0, cBeg, nOwnC;
};
__device__ __forceinline__ MasSmem mas_carve(unsigned char* base, const MasView& M)
{
    MasSmem S; size_t o = 0;
    S.nodeX = reinterpret_cast<double4*>(base + o);  o += (size_t)M.maxLocalNodes * 32;
    S.nodeA = reinterpret_cast<int4*>(base + o);     o += (size_t)M.maxLocalNodes * 16;
    S.nodeB = reinterpret_cast<int4*>(base + o);     o += (size_t)M.maxLocalNodes * 16;
    S.vinfo = reinterpret_cast<float4*>(base + o);   o += (size_t)M.rowsPer * 16;
    S.rc = reinterpret_cast<double*>(base + o);      o += ((size_t)M.maxLocalNodes * kMasDof + kMasBlk) * 8;
    S.e = reinterpret_cast<double*>(base + o);       o += ((size_t)M.maxLocalNodes * kMasDof + kMasBlk) * 8;
    S.rcAll = reinterpret_cast<double*>(base + o);   o += (size_t)M.ldC * 8;
    S.lvOff = reinterpret_cast<int*>(base + o);      o += (kMasMaxLevels + 2) * 4;
    o = (o + 15) / 16 * 16;
    S.cinv = reinterpret_cast<float*>(base + o);
    S.nLoc = 0; S.nSolve = 0; S.leaf0 = 0; S.cBeg = 0; S.nOwnC = 0;
    return S;
}

// stage the CTA's tables in shared memory; call once per solve
__device__ __forceinline__ void mas_init(const MasView& M, MasSmem& S, int cta, int rowBeg, int rowEnd)
{
    const int nT = blockDim.x, t = threadIdx.x;
    const int n0 = M.ctaNodeOff[cta];
    S.nLoc = M.ctaNodeOff[cta + 1] - n0;
    S.nSolve = M.ctaSolve[cta];
    S.leaf0 = M.ctaLeafBeg[cta];
    S.cBeg = M.ctaCBeg[cta];
    S.nOwnC = M.ctaCBeg[cta + 1] - S.cBeg;
    for (int i = t; i <= kMasMaxLevels; i += nT) S.lvOff[i] = M.ctaLvOff[(size_t)cta * (kMasMaxLevels + 1) + i];
    for (int i = t; i < S.nLoc; i += nT) { S.nodeA[i] = M.nodeA[n0 + i]; S.nodeB[i] = M.nodeB[n0 + i]; S.nodeX[i] = M.nodeX[n0 + i]; }
    for (int i = t; i < rowEnd - rowBeg; i += nT) S.vinfo[i] = M.vinfo[rowBeg + i];
    for (int i = t; i < M.maxLocalNodes * kMasDof + kMasBlk; i += nT) { S.rc[i] = 0.0; S.e[i] = 0.0; }
    for (int i = t; i < M.ldC; i += nT) S.rcAll[i] = 0.0;
    if (M.cinvInSmem) {                                // the CTA's rows of the coarse inverse: read from L2 once per solve instead of once per iteration
        const float4* src = reinterpret_cast<const float4*>(M.cinv + (size_t)S.cBeg * kMasDof * M.ldC);
        float4* dst = reinterpret_cast<float4*>(S.cinv);
        const int n4 = S.nOwnC * kMasDof * (M.ldC >> 2);
        for (int i = t; i < n4; i += nT) dst[i] = __ldg(src + i);
    }
    __syncthreads();
}

// ---- restriction: r (own rows, via getR(localRow) -> double2) -> rc of every local level; publishes the CTA
// node's 6 values.  The caller then ARRIVES at the grid barrier, runs mas_local_solves, and waits.
template <class GetR>
__device__ __forceinline__ void mas_restrict(const MasView& M, const MasSmem& S, int cta, GetR getR, int firstLevel = 1)
{
    const int nT = blockDim.x;
    // 8 lanes per (node, component): one child (a row for the leaves) per lane, then a 3-step shuffle reduction
    // (firstLevel = 2: the caller has already restricted the leaves, see mas_restrict_leaf)
    for (int l = firstLevel; l <= M.L; ++l) {
        const int done = S.lvOff[l - 1], end = S.lvOff[l];
        const int items = 16 * (end - done);
        for (int w = threadIdx.x; w < ((items + 31) & ~31); w += nT) {       // whole warps beyond the range skip
            const bool valid = w < items;
            const int k = w & 7, pair = w >> 3, node = done + (valid ? pair >> 1 : 0), comp = pair & 1;
            const int4 B = S.nodeB[node];
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            if (valid && k < B.z) {
                if (l == 1) {
                    const float4 vi = S.vinfo[B.y + k];
                    const double2 rr = getR(B.y + k);
                    const double rv = comp ? rr.y : rr.x;
                    a0 = mas_unpack(vi.x) * rv; a1 = mas_unpack(vi.y) * rv; a2 = mas_unpack(vi.z) * rv;
                } else {
                    const double4 X = S.nodeX[B.y + k];
                    const double* rc = S.rc + (size_t)(B.y + k) * kMasDof + 3 * comp;
                    a0 = rc[0]; a1 = X.x * rc[0] + X.z * rc[1]; a2 = X.y * rc[0] + X.z * rc[2];
                }
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (valid && k == 0) {
                double* o = S.rc + (size_t)node * kMasDof + 3 * comp;
                o[0] = a0; o[1] = a1; o[2] = a2;
            }
        }
        __syncthreads();
    }
    // publish the coarse residual of the own coarse nodes (the last local level)
    for (int i = threadIdx.x; i < S.nOwnC * kMasDof; i += nT) M.rcC[(size_t)S.cBeg * kMasDof + i] = S.rc[(size_t)S.lvOff[M.L - 1] * kMasDof + i];
}

// ---- level 1 fused into the caller's row loop: leaves are exactly 8 consecutive rows, so the 8 lanes that hold a leaf's
// residuals reduce them with three shuffle steps.  Must be called by ALL lanes of the warp (valid = the lane has a row).
__device__ __forceinline__ void mas_restrict_leaf(const MasSmem& S, int lr, bool valid, double2 r)
{
    double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (valid) {
        const float4 vi = S.vinfo[lr];
        a[0] = mas_unpack(vi.x) * r.x; a[1] = mas_unpack(vi.y) * r.x; a[2] = mas_unpack(vi.z) * r.x;
        a[3] = mas_unpack(vi.x) * r.y; a[4] = mas_unpack(vi.y) * r.y; a[5] = mas_unpack(vi.z) * r.y;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
#pragma unroll
        for (int q = 0; q < 6; ++q) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
    if (valid && (lr & 7) == 0) {
        double* o = S.rc + (size_t)(lr >> 3) * kMasDof;
#pragma unroll
        for (int q = 0; q < 6; ++q) o[q] = a[q];
    }
}

// ---- CTA-local group solves y = D_l^-1 rc_l for every local level below L (4 threads per output, the inverse
// rows streamed from L2/HBM with three independent 16-byte loads per thread); y lands in S.e.  No barrier inside.
__device__ __forceinline__ void mas_local_solves(const MasView& M, const MasSmem& S)
{
    const int nT = blockDim.x;
    const int items = S.nSolve * kMasDof * 4;
    for (int w = threadIdx.x; w < ((items + 31) & ~31); w += nT) {           // whole warps beyond the range skip
        const bool valid = w < items;
        const int node = valid ? w / (kMasDof * 4) : 0, rem = w % (kMasDof * 4), q = rem >> 2, part = rem & 3;
        const int4 A = S.nodeA[node];
        const int4 B = S.nodeB[node];
        const float4* row = reinterpret_cast<const float4*>(M.inv + (size_t)B.x + (size_t)(A.z + q) * kMasBlk);
        const double* rc = S.rc + (size_t)A.x * kMasDof;
        double y = 0.0;
        if (valid) {
            float4 a[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) a[j] = __ldg(row + part + 4 * j);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const double* r4 = rc + 4 * (part + 4 * j);
                y += mas_unpack(a[j].x) * r4[0] + mas_unpack(a[j].y) * r4[1] + mas_unpack(a[j].z) * r4[2] + mas_unpack(a[j].w) * r4[3];
            }
        }
        y += __shfl_xor_sync(0xffffffffu, y, 1);
        y += __shfl_xor_sync(0xffffffffu, y, 2);
        if (valid && part == 0) S.e[(size_t)node * kMasDof + q] = y;
    }
}

// ---- rows of the exact coarse solve: eC[row] = cinv[row, :] . rc for the CTA's rows, LPR lanes per row
__device__ __forceinline__ void mas_lds4(double (&r)[4], unsigned addr)
{
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r[0]), "=d"(r[1]) : "r"(addr));
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(r[2]), "=d"(r[3]) : "r"(addr));
}
template <int LPR, int UNR>
__device__ __forceinline__ void mas_coarse_rows(const float* __restrict__ cbase, int ldC, int ld4, int rows, const double* rcAll, double* eC, int nT)
{
    const int sub = threadIdx.x & (LPR - 1);
    const unsigned rcS = (unsigned)__cvta_generic_to_shared(rcAll) + 32u * sub;
    for (int row0 = 0; row0 < rows; row0 += nT / LPR) {
        const int row = row0 + threadIdx.x / LPR;
        const bool valid = row < rows;
        const float4* a = reinterpret_cast<const float4*>(cbase + (size_t)(valid ? row : 0) * ldC) + sub;
        double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
        if (valid) {
            int c4 = sub;
            for (; c4 + (UNR - 1) * LPR < ld4; c4 += UNR * LPR) {
                float4 u[UNR];
#pragma unroll
                for (int j = 0; j < UNR; ++j) u[j] = a[(c4 - sub) + j * LPR];
                const unsigned rs = rcS + 32u * (c4 - sub);
#pragma unroll
                for (int j = 0; j < UNR; ++j) {
                    double r[4];
                    mas_lds4(r, rs + 32u * LPR * j);
                    y0 = fma(mas_unpack(u[j].x), r[0], y0); y1 = fma(mas_unpack(u[j].y), r[1], y1);
                    y2 = fma(mas_unpack(u[j].z), r[2], y2); y3 = fma(mas_unpack(u[j].w), r[3], y3);
                }
            }
            for (; c4 < ld4; c4 += LPR) {
                const float4 u = a[c4 - sub];
                double r[4];
                mas_lds4(r, rcS + 32u * (c4 - sub));
                y0 = fma(mas_unpack(u.x), r[0], y0); y1 = fma(mas_unpack(u.y), r[1], y1);
                y2 = fma(mas_unpack(u.z), r[2], y2); y3 = fma(mas_unpack(u.w), r[3], y3);
            }
        }
        double y = (y0 + y1) + (y2 + y3);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
        if (valid && sub == 0) eC[row] = y;
    }
}

// ---- after the barrier: the exact coarse solve for the own coarse nodes (one warp per output row, the inverse
// rows streamed from L2 with 16-byte loads), then the local down sweep (prolongation adds).  Leaves the coarse
// correction coefficients of every local node in S.e; the caller adds m * (e0 + lx e1 + ly e2, e3 + lx e4 + ly e5)
// of the row's leaf to the block-Jacobi part of z.
template <int UNR>
__device__ __forceinline__ void mas_down(const MasView& M, const MasSmem& S, int cta, long long* stamp = nullptr)
{
    const int nT = blockDim.x, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < M.nC; i += nT) S.rcAll[i] = __ldcg(M.rcC + i);
    __syncthreads();
    if (stamp) stamp[0] = clock64();
    {   // LPR lanes per output row (32 when the CTA's rows then still fit one pass, else 16).  The pass is bound by
        // INSTRUCTION ISSUE, not by memory (measured: 7k cycles for 83 KB at 10k faces, the same with the loads or
        // the shared-memory reads removed), so the main loop is straight-line code: UNR 16-byte loads of the inverse
        // row per lane issued together, the residual read with 16-byte shared loads at immediate offsets, four
        // independent FMA chains, no predicates; a predicated tail takes the rest of the row.
        const int rows = S.nOwnC * kMasDof, ld4 = M.ldC >> 2;
        double* eC = S.e + (size_t)S.lvOff[M.L - 1] * kMasDof;
        const float* cbase = M.cinvInSmem ? S.cinv : M.cinv + (size_t)S.cBeg * kMasDof * M.ldC;
        if (rows * 32 <= nT) mas_coarse_rows<32, UNR>(cbase, M.ldC, ld4, rows, S.rcAll, eC, nT);
        else mas_coarse_rows<16, UNR>(cbase, M.ldC, ld4, rows, S.rcAll, eC, nT);
    }
    __syncthreads();
    if (stamp) stamp[1] = clock64();
    // local down sweep: e = y + prolongation of the parent's e; levels are contiguous and ascending in the node list
    for (int l = M.L - 1; l >= 1; --l) {
        const int beg = S.lvOff[l - 1], end = S.lvOff[l];
        for (int w = threadIdx.x; w < kMasDof * (end - beg); w += nT) {
            const int node = beg + w / kMasDof, q = w % kMasDof, qq = q % 3;
            const int4 A = S.nodeA[node];
            const double4 X = S.nodeX[node];
            const double* ep = S.e + (size_t)A.w * kMasDof + 3 * (q / 3);
            S.e[(size_t)node * kMasDof + q] += qq == 0 ? ep[0] + X.x * ep[1] + X.y * ep[2] : X.z * ep[qq];
        }
        __syncthreads();
    }
}
#endif  // __CUDACC__

}  // namespace ocb
