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
// |r|^2 — no other communication.  Measured on the reference's matrices (tools/mas_proto.py): 2466 -> 264 CG
// iterations at 10k faces (Tutte state), 11 500 -> 755 at 160k faces.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ocb {

static constexpr int kMasLeaf = 8;          // vertices per leaf
static constexpr int kMasGroup = 8;         // nodes per group
static constexpr int kMasDof = 6;           // DOFs per node
static constexpr int kMasBlk = kMasGroup * kMasDof;   // 48
static constexpr int kMasMaxLevels = 12;

struct MasLevel {
    int nNodes, nGroups;
    const int32_t* childBeg;   // nNodes + 1: rows (level 1) or nodes of the level below
    const int32_t* groupBeg;   // nGroups + 1: nodes of THIS level per group (= childBeg of the level above)
    const int32_t* parent;     // nNodes: group of the node
    const double4* geom;       // nNodes: {cx, cy, s, -}
    const int32_t* ctaBeg;     // grid + 1: first node of every CTA (levels <= Lloc), else nullptr
    const float* inv;          // nGroups x 48 x 48, symmetric
};

struct MasView {
    int L, Lloc, grid;         // L == 0: preconditioner disabled (block-Jacobi only)
    MasLevel lv[kMasMaxLevels];          // lv[l-1] = level l
    const float4* vinfo;       // per row: {m, m*lx, m*ly, leaf id (int bits)}, m = 0 for fixed vertices
    double* rcCta;             // grid x 6: restricted residual of every CTA node (exchange buffer)
    int topNodes;              // nodes of levels Lloc..L
    int maxLocalNodes;         // max over CTAs of the nodes of levels 1..Lloc
};

__host__ __device__ inline size_t mas_smem_bytes(int maxLocalNodes, int topNodes)
{
    // rc + e for the local nodes, rc for the top nodes, chain scratch, per-level offsets
    return (size_t)maxLocalNodes * kMasDof * 8 * 2 + (size_t)topNodes * kMasDof * 8 + 64 * 8 + 4 * kMasMaxLevels * 4;
}

#ifdef __CUDACC__
struct MasSmem {
    double* rc; double* e; double* rcTop; double* chain; int* off;     // off[l-1]: first local node of level l in rc/e; off[12+l-1]: top offset
};
__device__ __forceinline__ MasSmem mas_carve(unsigned char* base, const MasView& M)
{
    MasSmem S; size_t o = 0;
    S.rc = reinterpret_cast<double*>(base + o);     o += (size_t)M.maxLocalNodes * kMasDof * 8;
    S.e = reinterpret_cast<double*>(base + o);      o += (size_t)M.maxLocalNodes * kMasDof * 8;
    S.rcTop = reinterpret_cast<double*>(base + o);  o += (size_t)M.topNodes * kMasDof * 8;
    S.chain = reinterpret_cast<double*>(base + o);  o += 64 * 8;
    S.off = reinterpret_cast<int*>(base + o);
    return S;
}

// per-CTA offsets of the local levels in the shared arrays and of the top levels in rcTop; call once, all threads
__device__ __forceinline__ void mas_init(const MasView& M, const MasSmem& S, int cta)
{
    if (threadIdx.x == 0) {
        int o = 0;
        for (int l = 1; l <= M.Lloc; ++l) { S.off[l - 1] = o; o += M.lv[l - 1].ctaBeg[cta + 1] - M.lv[l - 1].ctaBeg[cta]; }
        o = 0;
        for (int l = M.Lloc; l <= M.L; ++l) { S.off[kMasMaxLevels + l - 1] = o; o += M.lv[l - 1].nNodes; }
    }
    __syncthreads();
}

// restriction of a child's 3 coefficients (one component) into its parent's basis
__device__ __forceinline__ void mas_restrict3(const double4 gc, const double4 gp, const double* rc, double* acc)
{
    const double is = 1.0 / gp.z, tx = (gc.x - gp.x) * is, ty = (gc.y - gp.y) * is, rho = gc.z * is;
    acc[0] += rc[0];
    acc[1] += tx * rc[0] + rho * rc[1];
    acc[2] += ty * rc[0] + rho * rc[2];
}
// prolongation of the parent's coefficients (one component) into the child's basis
__device__ __forceinline__ double mas_prolong1(const double4 gc, const double4 gp, const double* ep, int q)
{
    const double is = 1.0 / gp.z;
    if (q == 0) return ep[0] + (gc.x - gp.x) * is * ep[1] + (gc.y - gp.y) * is * ep[2];
    return gc.z * is * ep[q];
}

// ---- up sweep: r (own rows, via getR(localRow) -> double2) -> local levels; publishes the CTA node's 6 values.
// Ends with the data in S.rc; the caller must cross a grid-wide barrier before mas_down.
template <class GetR>
__device__ __forceinline__ void mas_up(const MasView& M, const MasSmem& S, int cta, int rowBeg, GetR getR)
{
    const int nT = blockDim.x;
    {   // level 1: one thread per (leaf, component)
        const MasLevel& V = M.lv[0];
        const int n0 = V.ctaBeg[cta], n1 = V.ctaBeg[cta + 1];
        for (int w = threadIdx.x; w < 2 * (n1 - n0); w += nT) {
            const int leaf = n0 + (w >> 1), comp = w & 1;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            const int r1 = V.childBeg[leaf + 1];
            for (int row = V.childBeg[leaf]; row < r1; ++row) {
                const float4 vi = __ldg(M.vinfo + row);
                const double2 rr = getR(row - rowBeg);
                const double rv = comp ? rr.y : rr.x;
                a0 += (double)vi.x * rv; a1 += (double)vi.y * rv; a2 += (double)vi.z * rv;
            }
            double* o = S.rc + (size_t)(S.off[0] + (leaf - n0)) * kMasDof + 3 * comp;
            o[0] = a0; o[1] = a1; o[2] = a2;
        }
    }
    __syncthreads();
    for (int l = 2; l <= M.Lloc; ++l) {
        const MasLevel& V = M.lv[l - 1];
        const MasLevel& C = M.lv[l - 2];
        const int n0 = V.ctaBeg[cta], n1 = V.ctaBeg[cta + 1], c0 = C.ctaBeg[cta];
        for (int w = threadIdx.x; w < 2 * (n1 - n0); w += nT) {
            const int node = n0 + (w >> 1), comp = w & 1;
            const double4 gp = V.geom[node];
            double acc[3] = {0.0, 0.0, 0.0};
            const int ce = V.childBeg[node + 1];
            for (int ch = V.childBeg[node]; ch < ce; ++ch)
                mas_restrict3(C.geom[ch], gp, S.rc + (size_t)(S.off[l - 2] + (ch - c0)) * kMasDof + 3 * comp, acc);
            double* o = S.rc + (size_t)(S.off[l - 1] + (node - n0)) * kMasDof + 3 * comp;
            o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
        }
        __syncthreads();
    }
    if (threadIdx.x < kMasDof && cta < M.lv[M.Lloc - 1].nNodes)
        M.rcCta[(size_t)cta * kMasDof + threadIdx.x] = S.rc[(size_t)S.off[M.Lloc - 1] * kMasDof + threadIdx.x];
}

// ---- after the barrier: top levels (redundantly in every CTA), then the local down sweep.  Leaves the coarse
// correction coefficients of every local LEAF in S.e[(off[0] + leaf) * 6 ..]; the caller adds
// m * (e0 + lx e1 + ly e2, e3 + lx e4 + ly e5) to the block-Jacobi part of z.
__device__ __forceinline__ void mas_down(const MasView& M, const MasSmem& S, int cta)
{
    const int nT = blockDim.x;
    const int* offTop = S.off + kMasMaxLevels;
    // top up-sweep
    const int nCtaNodes = M.lv[M.Lloc - 1].nNodes;          // CTAs that own rows (trailing CTAs may be empty)
    for (int i = threadIdx.x; i < nCtaNodes * kMasDof; i += nT) S.rcTop[i] = __ldcg(M.rcCta + i);
    __syncthreads();
    for (int l = M.Lloc + 1; l <= M.L; ++l) {
        const MasLevel& V = M.lv[l - 1];
        const MasLevel& C = M.lv[l - 2];
        for (int w = threadIdx.x; w < 2 * V.nNodes; w += nT) {
            const int node = w >> 1, comp = w & 1;
            const double4 gp = V.geom[node];
            double acc[3] = {0.0, 0.0, 0.0};
            const int ce = V.childBeg[node + 1];
            for (int ch = V.childBeg[node]; ch < ce; ++ch)
                mas_restrict3(C.geom[ch], gp, S.rcTop + (size_t)(offTop[l - 2] + ch) * kMasDof + 3 * comp, acc);
            double* o = S.rcTop + (size_t)(offTop[l - 1] + node) * kMasDof + 3 * comp;
            o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
        }
        __syncthreads();
    }
    // ancestor chain of this CTA, top down: warp 0, 4 lanes per DOF
    if (threadIdx.x < 32 && cta < nCtaNodes) {
        const int lane = threadIdx.x;
        // ancestors: anc[l] for l = Lloc..L
        int anc[kMasMaxLevels + 2];
        anc[M.Lloc] = cta;
        for (int l = M.Lloc; l < M.L; ++l) anc[l + 1] = M.lv[l - 1].parent[anc[l]];
        double* ch = S.chain;                    // e of the level above (6) at ch[0..5], new at ch[8..13]
        if (lane < kMasDof) ch[lane] = 0.0;
        __syncwarp();
        for (int l = M.L; l >= M.Lloc; --l) {
            const MasLevel& V = M.lv[l - 1];
            const int a = anc[l];
            const int g = V.parent[a];
            const int gb = V.groupBeg[g], nch = V.groupBeg[g + 1] - gb;
            const int slot = a - gb;
            // 4 lanes per DOF: lane = q * 4 + part
            const int q = lane >> 2, part = lane & 3;
            double y = 0.0;
            if (q < kMasDof) {
                const float* row = V.inv + (size_t)g * kMasBlk * kMasBlk + (size_t)(slot * kMasDof + q) * kMasBlk;
                const double* rc = S.rcTop + (size_t)(offTop[l - 1] + gb) * kMasDof;
                for (int k = part; k < nch * kMasDof; k += 4) y += (double)__ldg(row + k) * rc[k];
            }
            y += __shfl_xor_sync(0xffffffffu, y, 1);
            y += __shfl_xor_sync(0xffffffffu, y, 2);
            if (q < kMasDof && part == 0) {
                double e = y;
                if (l < M.L) e += mas_prolong1(V.geom[a], M.lv[l].geom[g], ch + 3 * (q / 3), q % 3);
                ch[8 + q] = e;
            }
            __syncwarp();
            if (lane < kMasDof) ch[lane] = ch[8 + lane];
            __syncwarp();
        }
        if (lane < kMasDof) S.e[(size_t)S.off[M.Lloc - 1] * kMasDof + lane] = ch[lane];
    }
    __syncthreads();
    // local down sweep
    for (int l = M.Lloc - 1; l >= 1; --l) {
        const MasLevel& V = M.lv[l - 1];
        const MasLevel& U = M.lv[l];
        const int n0 = V.ctaBeg[cta], n1 = V.ctaBeg[cta + 1], u0 = U.ctaBeg[cta];
        for (int w = threadIdx.x; w < kMasDof * (n1 - n0); w += nT) {
            const int node = n0 + w / kMasDof, q = w % kMasDof;
            const int g = V.parent[node];
            const int gb = V.groupBeg[g], nch = V.groupBeg[g + 1] - gb;
            const float* row = V.inv + (size_t)g * kMasBlk * kMasBlk + (size_t)((node - gb) * kMasDof + q) * kMasBlk;
            const double* rc = S.rc + (size_t)(S.off[l - 1] + (gb - n0)) * kMasDof;
            double y0 = 0.0, y1 = 0.0;
            const int nk = nch * kMasDof;
            for (int k = 0; k + 1 < nk; k += 2) { y0 += (double)__ldg(row + k) * rc[k]; y1 += (double)__ldg(row + k + 1) * rc[k + 1]; }
            const double* ep = S.e + (size_t)(S.off[l] + (g - u0)) * kMasDof + 3 * (q / 3);
            S.e[(size_t)(S.off[l - 1] + (node - n0)) * kMasDof + q] = y0 + y1 + mas_prolong1(V.geom[node], U.geom[g], ep, q % 3);
        }
        __syncthreads();
    }
}
#endif  // __CUDACC__

}  // namespace ocb
