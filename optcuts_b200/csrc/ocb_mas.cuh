// optcuts_b200 — multilevel additive Schwarz (MAS) preconditioner of the PCG solve: device-side view and the
// CTA-level apply (included by ocb_pcg.cu; set up by ocb_mas.cu).
//
//   M^-1 = blockdiag2x2(A)^-1 + sum_{l=1..L} P_l D_l^-1 P_l^T
//
// The solver rows are ordered by recursive coordinate bisection of the UV positions so that a persistent CTA's
// contiguous row range is a compact patch, split into leaves of <= 8 vertices.  Level-l nodes carry 6 DOFs: the
// affine displacement fields (1, x, y) x (u, v) on the node, x/y in node-local coordinates.  8 consecutive nodes
// form a group (= a node of level l+1); D_l is the block diagonal of the Galerkin matrix P_l^T A P_l over the
// groups (48x48 blocks, inverted on the device, kept in fp32).  Levels 1..Lloc live entirely inside one CTA
// (level Lloc has exactly one node per CTA); the levels above are evaluated redundantly by every CTA from the
// `grid x 6` restricted residuals, which cross CTAs through global memory at the grid barrier that also carries
// |r|^2 — no other communication.  Everything an iteration needs except the group inverses (node tables, per-row
// basis values, the inverse rows of the CTA's ancestor chain) is staged in shared memory once per solve, and the
// CTA-local group solves run between the arrive and the wait of that barrier.  Measured on the reference's
// matrices (tools/mas_proto.py): 2466 -> 276 CG iterations at 10k faces (Tutte state), 11 500 -> 760 at 160k.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ocb {

static constexpr int kMasLeaf = 8;          // vertices per leaf
static constexpr int kMasGroup = 8;         // nodes per group
static constexpr int kMasDof = 6;           // DOFs per node
static constexpr int kMasBlk = kMasGroup * kMasDof;   // 48
static constexpr int kMasMaxLevels = 12;
static constexpr int kMasChainRow = kMasDof * kMasBlk;   // 288 floats: the 6 inverse rows of one chain level

// one level of the hierarchy as the set-up kernels see it (global node indices)
struct MasLevel {
    int nNodes, nGroups;
    const int32_t* childBeg;   // nNodes + 1: rows (level 1) or nodes of the level below
    const int32_t* groupBeg;   // nGroups + 1: nodes of THIS level per group (= childBeg of the level above)
    const int32_t* parent;     // nNodes: group of the node
    const double4* geom;       // nNodes: {cx, cy, s, -}
    float* inv;                // nGroups x 48 x 48
};

// Host-precomputed tables, CTA-local indices.  A CTA's local nodes are numbered level 1 first, then level 2, ...
// up to its single level-Lloc node; "top" nodes are the nodes of levels Lloc..L of the whole hierarchy.
struct MasView {
    int L, Lloc, grid, nCh;    // L == 0: preconditioner disabled (block-Jacobi only); nCh = L - Lloc + 1
    int topNodes;              // nodes of levels Lloc..L
    int nCtaNodes;             // nodes of level Lloc (= CTAs that own rows)
    int maxLocalNodes;         // max over CTAs of the nodes of levels 1..Lloc
    int rowsPer;               // rows per CTA (vinfo staging)
    const int32_t* ctaNodeOff; // grid + 1: first entry of every CTA in nodeA/nodeB/nodeX
    const int32_t* ctaSolve;   // grid: local nodes below level Lloc (they take part in a CTA-local group solve)
    const int32_t* ctaLvOff;   // grid x (kMasMaxLevels + 1): first local node of level l at [l - 1]; [Lloc] = all local nodes
    const int32_t* ctaLeafBeg; // grid + 1: first (global) leaf of every CTA
    const int4* nodeA;         // {local index of the group's first child, nk = 6 * children of the group, 6 * slot in the group, local index of the parent}
    const int4* nodeB;         // {float offset of the group's inverse, first child (local node; local row for leaves), children, level}
    const double4* nodeX;      // {tx, ty, rho, -}: transfer of the node's coefficients to its parent's basis
    const int2* topUp;         // per top node: {first child (top index), children}; level Lloc: unused
    const double4* topX;       // per top node: transfer to its parent (levels < L)
    const int32_t* topLevelOff;// nCh + 1: first top index of levels Lloc, Lloc+1, ..., L
    const int4* chainM;        // grid x kMasMaxLevels, entry j = level L - j: {top index of the group's first child, nk, float offset of the 6 inverse rows, top index of the CTA's ancestor}
    const float* inv;          // all group inverses (48 x 48 each, symmetric)
    const float4* vinfo;       // per row: {m, m*lx, m*ly, leaf id (int bits)}, m = 0 for fixed vertices
    double* rcCta;             // grid x 6: restricted residual of every CTA node (exchange buffer)
};

__host__ __device__ inline size_t mas_smem_bytes(int maxLocalNodes, int topNodes, int rowsPer, int nCh)
{
    size_t b = 0;
    b += ((size_t)maxLocalNodes * kMasDof + kMasBlk) * 8 * 2;      // rc (+ zero pad), e
    b += (size_t)maxLocalNodes * (16 + 16 + 32);                   // nodeA, nodeB, nodeX
    b += (size_t)rowsPer * 16;                                     // vinfo
    b += ((size_t)topNodes * kMasDof + kMasBlk) * 8;               // rcTop (+ zero pad)
    b += (size_t)topNodes * (8 + 32);                              // topUp, topX
    b += (size_t)nCh * (kMasChainRow * 4 + 16);                    // chain inverse rows + chainM
    b += 32 * 8 + 2 * (kMasMaxLevels + 2) * 4 + 64;                // chain scratch, topLevelOff, lvOff, slack
    return (b + 15) / 16 * 16;
}

#ifdef __CUDACC__
struct MasSmem {
    double* rc; double* e; double* rcTop; double* chain;
    int4* nodeA; int4* nodeB; double4* nodeX; float4* vinfo; int2* topUp; double4* topX; float* chainInv; int4* chainM; int* topLevelOff;
    int* lvOff;
    int nLoc, nSolve, leaf0;
};
__device__ __forceinline__ MasSmem mas_carve(unsigned char* base, const MasView& M)
{
    MasSmem S; size_t o = 0;
    S.nodeX = reinterpret_cast<double4*>(base + o);  o += (size_t)M.maxLocalNodes * 32;
    S.topX = reinterpret_cast<double4*>(base + o);   o += (size_t)M.topNodes * 32;
    S.nodeA = reinterpret_cast<int4*>(base + o);     o += (size_t)M.maxLocalNodes * 16;
    S.nodeB = reinterpret_cast<int4*>(base + o);     o += (size_t)M.maxLocalNodes * 16;
    S.vinfo = reinterpret_cast<float4*>(base + o);   o += (size_t)M.rowsPer * 16;
    S.chainM = reinterpret_cast<int4*>(base + o);    o += (size_t)M.nCh * 16;
    S.chainInv = reinterpret_cast<float*>(base + o); o += (size_t)M.nCh * kMasChainRow * 4;
    S.rc = reinterpret_cast<double*>(base + o);      o += ((size_t)M.maxLocalNodes * kMasDof + kMasBlk) * 8;
    S.e = reinterpret_cast<double*>(base + o);       o += ((size_t)M.maxLocalNodes * kMasDof + kMasBlk) * 8;
    S.rcTop = reinterpret_cast<double*>(base + o);   o += ((size_t)M.topNodes * kMasDof + kMasBlk) * 8;
    S.chain = reinterpret_cast<double*>(base + o);   o += 32 * 8;
    S.topUp = reinterpret_cast<int2*>(base + o);     o += (size_t)M.topNodes * 8;
    S.topLevelOff = reinterpret_cast<int*>(base + o); o += (kMasMaxLevels + 2) * 4;
    S.lvOff = reinterpret_cast<int*>(base + o);
    S.nLoc = 0; S.nSolve = 0; S.leaf0 = 0;
    return S;
}

// stage the CTA's tables in shared memory; call once per solve (after ocb_factorize: the chain rows are VALUES)
__device__ __forceinline__ void mas_init(const MasView& M, MasSmem& S, int cta, int rowBeg, int rowEnd)
{
    const int nT = blockDim.x, t = threadIdx.x;
    const int n0 = M.ctaNodeOff[cta];
    S.nLoc = M.ctaNodeOff[cta + 1] - n0;
    S.nSolve = M.ctaSolve[cta];
    S.leaf0 = M.ctaLeafBeg[cta];
    for (int i = t; i <= kMasMaxLevels; i += nT) S.lvOff[i] = M.ctaLvOff[(size_t)cta * (kMasMaxLevels + 1) + i];
    for (int i = t; i < S.nLoc; i += nT) { S.nodeA[i] = M.nodeA[n0 + i]; S.nodeB[i] = M.nodeB[n0 + i]; S.nodeX[i] = M.nodeX[n0 + i]; }
    for (int i = t; i < rowEnd - rowBeg; i += nT) S.vinfo[i] = M.vinfo[rowBeg + i];
    for (int i = t; i < M.topNodes; i += nT) { S.topUp[i] = M.topUp[i]; S.topX[i] = M.topX[i]; }
    for (int i = t; i <= M.nCh; i += nT) S.topLevelOff[i] = M.topLevelOff[i];
    for (int i = t; i < M.nCh; i += nT) S.chainM[i] = M.chainM[(size_t)cta * kMasMaxLevels + i];
    for (int i = t; i < M.nCh * kMasChainRow; i += nT) {
        const int j = i / kMasChainRow;
        S.chainInv[i] = M.inv[(size_t)M.chainM[(size_t)cta * kMasMaxLevels + j].z + (i - j * kMasChainRow)];
    }
    for (int i = t; i < M.maxLocalNodes * kMasDof + kMasBlk; i += nT) { S.rc[i] = 0.0; S.e[i] = 0.0; }
    for (int i = t; i < M.topNodes * kMasDof + kMasBlk; i += nT) S.rcTop[i] = 0.0;
    __syncthreads();
}

// ---- restriction: r (own rows, via getR(localRow) -> double2) -> rc of every local level; publishes the CTA
// node's 6 values.  The caller then ARRIVES at the grid barrier, runs mas_local_solves, and waits.
template <class GetR>
__device__ __forceinline__ void mas_restrict(const MasView& M, const MasSmem& S, int cta, GetR getR)
{
    const int nT = blockDim.x;
    // 8 lanes per (node, component): one child (a row for the leaves) per lane, then a 3-step shuffle reduction
    for (int l = 1; l <= M.Lloc; ++l) {
        const int done = S.lvOff[l - 1], end = S.lvOff[l];
        const int items = 16 * (end - done);
        for (int w0 = 0; w0 < items; w0 += nT) {
            const int w = w0 + threadIdx.x;
            const bool valid = w < items;
            const int k = w & 7, pair = w >> 3, node = done + (valid ? pair >> 1 : 0), comp = pair & 1;
            const int4 B = S.nodeB[node];
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            if (valid && k < B.z) {
                if (l == 1) {
                    const float4 vi = S.vinfo[B.y + k];
                    const double2 rr = getR(B.y + k);
                    const double rv = comp ? rr.y : rr.x;
                    a0 = (double)vi.x * rv; a1 = (double)vi.y * rv; a2 = (double)vi.z * rv;
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
    if (threadIdx.x < kMasDof && cta < M.nCtaNodes) M.rcCta[(size_t)cta * kMasDof + threadIdx.x] = S.rc[(size_t)(S.nLoc - 1) * kMasDof + threadIdx.x];
}

// ---- CTA-local group solves y = D_l^-1 rc_l for every local level below Lloc (4 threads per output, the inverse
// rows streamed from L2/HBM with three independent 16-byte loads per thread); y lands in S.e.  No barrier inside.
__device__ __forceinline__ void mas_local_solves(const MasView& M, const MasSmem& S)
{
    const int nT = blockDim.x;
    const int items = S.nSolve * kMasDof * 4;
    for (int w0 = 0; w0 < items; w0 += nT) {
        const int w = w0 + threadIdx.x;
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
                y += (double)a[j].x * r4[0] + (double)a[j].y * r4[1] + (double)a[j].z * r4[2] + (double)a[j].w * r4[3];
            }
        }
        y += __shfl_xor_sync(0xffffffffu, y, 1);
        y += __shfl_xor_sync(0xffffffffu, y, 2);
        if (valid && part == 0) S.e[(size_t)node * kMasDof + q] = y;
    }
}

// ---- after the barrier: top levels (redundantly in every CTA, shared memory only), then the local down sweep
// (prolongation adds).  Leaves the coarse correction coefficients of every local node in S.e; the caller adds
// m * (e0 + lx e1 + ly e2, e3 + lx e4 + ly e5) of the row's leaf to the block-Jacobi part of z.
__device__ __forceinline__ void mas_down(const MasView& M, const MasSmem& S, int cta)
{
    const int nT = blockDim.x;
    for (int i = threadIdx.x; i < M.nCtaNodes * kMasDof; i += nT) S.rcTop[i] = __ldcg(M.rcCta + i);
    __syncthreads();
    for (int j = 1; j < M.nCh; ++j) {                 // top up-sweep: levels Lloc+1 .. L
        const int n0 = S.topLevelOff[j], n1 = S.topLevelOff[j + 1];
        const int items = 16 * (n1 - n0);
        for (int w0 = 0; w0 < items; w0 += nT) {
            const int w = w0 + threadIdx.x;
            const bool valid = w < items;
            const int k = w & 7, pair = w >> 3, node = n0 + (valid ? pair >> 1 : 0), comp = pair & 1;
            const int2 U = S.topUp[node];
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            if (valid && k < U.y) {
                const double4 X = S.topX[U.x + k];
                const double* rc = S.rcTop + (size_t)(U.x + k) * kMasDof + 3 * comp;
                a0 = rc[0]; a1 = X.x * rc[0] + X.z * rc[1]; a2 = X.y * rc[0] + X.z * rc[2];
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (valid && k == 0) {
                double* o = S.rcTop + (size_t)node * kMasDof + 3 * comp;
                o[0] = a0; o[1] = a1; o[2] = a2;
            }
        }
        __syncthreads();
    }
    // ancestor chain of this CTA, top down: warp 0, 4 lanes per DOF, shared memory only
    if (threadIdx.x < 32 && cta < M.nCtaNodes) {
        const int lane = threadIdx.x, q = lane >> 2, part = lane & 3;
        double* ch = S.chain;                          // e of the level above at ch[0..5], new at ch[8..13]
        if (lane < kMasDof) ch[lane] = 0.0;
        __syncwarp();
        for (int j = 0; j < M.nCh; ++j) {              // level L - j
            const int4 C = S.chainM[j];
            double y = 0.0;
            if (q < kMasDof) {
                const float* row = S.chainInv + j * kMasChainRow + q * kMasBlk;
                const double* rc = S.rcTop + (size_t)C.x * kMasDof;
                for (int k = part; k < C.y; k += 4) y += (double)row[k] * rc[k];
            }
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            if (q < kMasDof && part == 0) {
                double e = y;
                if (j > 0) {
                    const double4 X = S.topX[C.w];
                    const double* ep = ch + 3 * (q / 3);
                    const int qq = q % 3;
                    e += qq == 0 ? ep[0] + X.x * ep[1] + X.y * ep[2] : X.z * ep[qq];
                }
                ch[8 + q] = e;
            }
            __syncwarp();
            if (lane < kMasDof) ch[lane] = ch[8 + lane];
            __syncwarp();
        }
        if (lane < kMasDof) S.e[(size_t)(S.nLoc - 1) * kMasDof + lane] = ch[lane];
    }
    __syncthreads();
    // local down sweep: e = y + prolongation of the parent's e; levels are contiguous and ascending in the node list
    for (int l = M.Lloc - 1; l >= 1; --l) {
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
