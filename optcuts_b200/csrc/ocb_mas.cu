// optcuts_b200 — multilevel additive Schwarz preconditioner: hierarchy construction (host, at pattern time) and
// the per-factorisation set-up kernels (Galerkin coarse matrices, group-block inversion).  See ocb_mas.cuh for the
// operator and the apply; this replaces the numeric factorisation of Eigen::SimplicialLDLT
// (EigenLibSolver.cpp:80-93) as the per-Newton-iteration solver set-up.
#include "ocb_internal.cuh"
#include "ocb_mas.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace ocb {

#define KCHECK(c) do { (c)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return cuda_fail((c), _e, __func__); } while (0)

// ------------------------------------------------------------------------------------------------
// host: recursive coordinate bisection into consecutive parts of prescribed sizes
namespace {
struct Rcb {
    const double* xy;
    int32_t* idx;
    void split(int beg, int end, const int* pre, int nParts) const      // pre: prefix sums of the part sizes (nParts + 1)
    {
        if (nParts <= 1 || end - beg <= 1) return;
        const int h = nParts / 2;
        const int nl = pre[h] - pre[0];
        if (nl > 0 && nl < end - beg) {
            double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
            for (int i = beg; i < end; ++i) {
                const double* p = xy + 2 * (size_t)idx[i];
                for (int a = 0; a < 2; ++a) { if (p[a] < lo[a]) lo[a] = p[a]; if (p[a] > hi[a]) hi[a] = p[a]; }
            }
            const int ax = (hi[1] - lo[1] > hi[0] - lo[0]) ? 1 : 0;
            const double* q = xy;
            std::nth_element(idx + beg, idx + beg + nl, idx + end, [q, ax](int32_t a, int32_t b) {
                const double va = q[2 * (size_t)a + ax], vb = q[2 * (size_t)b + ax];
                return va < vb || (va == vb && a < b);
            });
        }
        split(beg, beg + nl, pre, h);
        split(beg + nl, end, pre + h, nParts - h);
    }
};

static void even_prefix(int n, int k, std::vector<int>& pre)
{
    pre.resize((size_t)k + 1);
    pre[0] = 0;
    for (int i = 0; i < k; ++i) pre[i + 1] = pre[i] + n / k + (i < n % k ? 1 : 0);
}
}  // namespace

// Solver order + hierarchy.  xy: 2 doubles per INTERNAL vertex (non-finite values are treated as 0).
// Fills c->hVertOf (row -> internal vertex), c->hRowOf and c->masH (without the level patterns).
int mas_build_hierarchy(ocb_ctx* c, const double* xyIn, int grid)
{
    const int n = c->nVtot;
    MasHost& H = c->masH;
    H = MasHost();
    std::vector<double> xy(2 * (size_t)n);
    for (size_t i = 0; i < xy.size(); ++i) xy[i] = std::isfinite(xyIn[i]) ? xyIn[i] : 0.0;
    if (grid < 1) grid = 1;
    const int rowsPer = (n + grid - 1) / grid;
    // stage 1: CTA chunks of exactly rowsPer rows (the last one shorter; trailing CTAs may be empty)
    std::vector<int> ctaPre((size_t)grid + 1, 0);
    for (int b = 0; b < grid; ++b) ctaPre[b + 1] = std::min(n, (b + 1) * rowsPer);
    int nonEmpty = 0;
    for (int b = 0; b < grid; ++b) if (ctaPre[b + 1] > ctaPre[b]) nonEmpty = b + 1;
    c->hVertOf.resize((size_t)n);
    for (int i = 0; i < n; ++i) c->hVertOf[i] = i;
    Rcb R{xy.data(), c->hVertOf.data()};
    R.split(0, n, ctaPre.data(), nonEmpty);
    // stage 2: leaves inside every chunk
    H.grid = grid;
    H.lv.clear();
    H.lv.emplace_back();
    {
        MasHost::Level& L1 = H.lv[0];
        L1.childBeg.push_back(0);
        L1.ctaBeg.assign((size_t)grid + 1, 0);
        std::vector<int> pre;
        for (int b = 0; b < grid; ++b) {
            const int beg = ctaPre[b], m = ctaPre[b + 1] - beg;
            if (m > 0) {
                const int nl = (m + kMasLeaf - 1) / kMasLeaf;
                even_prefix(m, nl, pre);
                R.split(beg, beg + m, pre.data(), nl);
                for (int k = 1; k <= nl; ++k) L1.childBeg.push_back(beg + pre[k]);
            }
            L1.ctaBeg[b + 1] = (int32_t)L1.childBeg.size() - 1;
        }
    }
    c->hRowOf.assign((size_t)n, 0);
    for (int r = 0; r < n; ++r) c->hRowOf[c->hVertOf[r]] = r;
    // geometry of the leaves + per-row info
    auto bbox_to_geom = [](const double* lo, const double* hi, double* g) {
        g[0] = 0.5 * (lo[0] + hi[0]); g[1] = 0.5 * (lo[1] + hi[1]);
        double s = 0.5 * std::max(hi[0] - lo[0], hi[1] - lo[1]);
        g[2] = s > 0.0 ? s : 1.0; g[3] = 0.0;
    };
    std::vector<double> lo, hi;              // 2 per node of the current level
    {
        MasHost::Level& L1 = H.lv[0];
        const int nl = (int)L1.childBeg.size() - 1;
        lo.assign(2 * (size_t)nl, 1e300); hi.assign(2 * (size_t)nl, -1e300);
        L1.geom.resize(4 * (size_t)nl);
        H.vinfo.resize(4 * (size_t)n);
        for (int k = 0; k < nl; ++k) {
            for (int r = L1.childBeg[k]; r < L1.childBeg[k + 1]; ++r) {
                const double* p = xy.data() + 2 * (size_t)c->hVertOf[r];
                for (int a = 0; a < 2; ++a) { lo[2 * k + a] = std::min(lo[2 * k + a], p[a]); hi[2 * k + a] = std::max(hi[2 * k + a], p[a]); }
            }
            double* g = L1.geom.data() + 4 * (size_t)k;
            bbox_to_geom(&lo[2 * k], &hi[2 * k], g);
            for (int r = L1.childBeg[k]; r < L1.childBeg[k + 1]; ++r) {
                const int v = c->hVertOf[r];
                const double* p = xy.data() + 2 * (size_t)v;
                const float m = c->hFixed[v] ? 0.0f : 1.0f;
                float* o = H.vinfo.data() + 4 * (size_t)r;
                o[0] = m; o[1] = m * (float)((p[0] - g[0]) / g[2]); o[2] = m * (float)((p[1] - g[1]) / g[2]);
                int32_t id = k; std::memcpy(o + 3, &id, 4);
            }
        }
    }
    // upper levels
    auto finish_level = [&](MasHost::Level& cur, MasHost::Level& nxt, const std::vector<int>& groupPre) {
        // groupPre: prefix over nodes of cur, one entry per node of nxt (+1)
        const int ng = (int)groupPre.size() - 1;
        nxt.childBeg.assign(groupPre.begin(), groupPre.end());
        cur.parent.resize(cur.childBeg.size() - 1);
        std::vector<double> lo2(2 * (size_t)ng, 1e300), hi2(2 * (size_t)ng, -1e300);
        nxt.geom.resize(4 * (size_t)ng);
        for (int g = 0; g < ng; ++g) {
            for (int k = groupPre[g]; k < groupPre[g + 1]; ++k) {
                cur.parent[k] = g;
                for (int a = 0; a < 2; ++a) { lo2[2 * g + a] = std::min(lo2[2 * g + a], lo[2 * k + a]); hi2[2 * g + a] = std::max(hi2[2 * g + a], hi[2 * k + a]); }
            }
            bbox_to_geom(&lo2[2 * g], &hi2[2 * g], nxt.geom.data() + 4 * (size_t)g);
        }
        lo.swap(lo2); hi.swap(hi2);
    };
    H.Lloc = 0;
    for (int l = 1; l < kMasMaxLevels; ++l) {
        MasHost::Level& cur = H.lv[l - 1];
        bool single = true;
        for (int b = 0; b < grid; ++b) if (cur.ctaBeg[b + 1] - cur.ctaBeg[b] > 1) single = false;
        if (single) { H.Lloc = l; break; }
        H.lv.emplace_back();
        MasHost::Level& c2 = H.lv[l - 1];
        MasHost::Level& nxt = H.lv[l];
        std::vector<int> groupPre(1, 0), pre;
        nxt.ctaBeg.assign((size_t)grid + 1, 0);
        for (int b = 0; b < grid; ++b) {
            const int k0 = c2.ctaBeg[b], k = c2.ctaBeg[b + 1] - k0;
            if (k > 0) {
                const int ng = (k + kMasGroup - 1) / kMasGroup;
                even_prefix(k, ng, pre);
                for (int g = 1; g <= ng; ++g) groupPre.push_back(k0 + pre[g]);
            }
            nxt.ctaBeg[b + 1] = (int32_t)groupPre.size() - 1;
        }
        finish_level(c2, nxt, groupPre);
    }
    if (H.Lloc == 0) return set_err(c, OCB_ERR_STATE, "MAS hierarchy: too many local levels");
    // a CTA without rows owns no node; the level above must still see `grid`-indexed CTA nodes: only non-empty CTAs carry one
    int l = H.Lloc;
    for (; l < kMasMaxLevels; ++l) {
        const int nn = (int)H.lv[l - 1].childBeg.size() - 1;
        if (nn <= kMasGroup) break;
        H.lv.emplace_back();
        MasHost::Level& cur = H.lv[l - 1];
        MasHost::Level& nxt = H.lv[l];
        const int ng = (nn + kMasGroup - 1) / kMasGroup;
        std::vector<int> groupPre;
        even_prefix(nn, ng, groupPre);
        finish_level(cur, nxt, groupPre);
    }
    H.L = l;
    {
        MasHost::Level& top = H.lv[H.L - 1];
        const int nn = (int)top.childBeg.size() - 1;
        if (nn > kMasGroup) return set_err(c, OCB_ERR_STATE, "MAS hierarchy: too many levels");
        top.parent.assign((size_t)nn, 0);
    }
    H.topNodes = 0;
    for (int k = H.Lloc; k <= H.L; ++k) H.topNodes += (int)H.lv[k - 1].childBeg.size() - 1;
    H.maxLocalNodes = 0;
    for (int b = 0; b < grid; ++b) {
        int s = 0;
        for (int k = 1; k <= H.Lloc; ++k) s += H.lv[k - 1].ctaBeg[b + 1] - H.lv[k - 1].ctaBeg[b];
        H.maxLocalNodes = std::max(H.maxLocalNodes, s);
    }
    H.enabled = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// host: node adjacency of every level from the solver-order BSR pattern, then one upload of everything
int mas_install(ocb_ctx* c)
{
    MasHost& H = c->masH;
    MasDev& D = c->masD;
    if (!H.enabled) { D.view.L = 0; return 0; }
    const int n = c->nVtot;
    std::vector<int32_t> nodeOf((size_t)n);       // row -> node of the current level
    for (int r = 0; r < n; ++r) { int32_t id; std::memcpy(&id, &H.vinfo[4 * (size_t)r + 3], 4); nodeOf[r] = id; }
    const std::vector<int32_t>* fineRowPtr = &c->hSRowPtr; const std::vector<int32_t>* fineColIdx = &c->hSColIdx;
    std::vector<int32_t> buf;
    for (int l = 1; l <= H.L; ++l) {
        MasHost::Level& V = H.lv[l - 1];
        const int nn = (int)V.childBeg.size() - 1;
        V.rowPtr.assign((size_t)nn + 1, 0); V.colIdx.clear();
        for (int k = 0; k < nn; ++k) {
            buf.clear();
            buf.push_back(k);
            for (int ch = V.childBeg[k]; ch < V.childBeg[k + 1]; ++ch)
                for (int b = (*fineRowPtr)[ch]; b < (*fineRowPtr)[ch + 1]; ++b) buf.push_back(nodeOf[(*fineColIdx)[b]]);
            std::sort(buf.begin(), buf.end());
            buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
            V.colIdx.insert(V.colIdx.end(), buf.begin(), buf.end());
            V.rowPtr[k + 1] = (int32_t)V.colIdx.size();
        }
        // next level: "rows" are this level's nodes, nodeOf = their parent
        nodeOf.assign(V.parent.begin(), V.parent.end());
        fineRowPtr = &V.rowPtr; fineColIdx = &V.colIdx;
    }
    // pack
    std::vector<int32_t> ints; std::vector<double> geom;
    struct Off { size_t childBeg, parent, ctaBeg, rowPtr, colIdx, geom, val, inv; int nNodes, nGroups, nnz; };
    std::vector<Off> off((size_t)H.L);
    size_t valTot = 0, invTot = 0;
    auto put = [&ints](const std::vector<int32_t>& v) { size_t o = ints.size(); ints.insert(ints.end(), v.begin(), v.end()); while (ints.size() & 3) ints.push_back(0); return o; };
    for (int l = 1; l <= H.L; ++l) {
        MasHost::Level& V = H.lv[l - 1];
        Off& o = off[l - 1];
        o.nNodes = (int)V.childBeg.size() - 1;
        o.nGroups = l < H.L ? (int)H.lv[l].childBeg.size() - 1 : 1;
        o.nnz = (int)V.colIdx.size();
        o.childBeg = put(V.childBeg); o.parent = put(V.parent);
        o.ctaBeg = l <= H.Lloc ? put(V.ctaBeg) : 0;
        o.rowPtr = put(V.rowPtr); o.colIdx = put(V.colIdx);
        o.geom = geom.size(); geom.insert(geom.end(), V.geom.begin(), V.geom.end());
        o.val = valTot; valTot += 36 * (size_t)o.nnz;
        o.inv = invTot; invTot += (size_t)kMasBlk * kMasBlk * o.nGroups;
    }
    // the top level's group list: [0, nNodes]
    std::vector<int32_t> topGroup(2, 0); topGroup[1] = off[H.L - 1].nNodes;
    const size_t topGroupOff = put(topGroup);
    OCB_CUDA(c, D.ints.reserve(ints.size() + 4, c->stream));
    OCB_CUDA(c, D.geom.reserve(geom.size() + 4, c->stream));
    OCB_CUDA(c, D.val.reserve(valTot + 4, c->stream));
    OCB_CUDA(c, D.inv.reserve(invTot + 4, c->stream));
    OCB_CUDA(c, D.vinfo.reserve(4 * (size_t)n + 4, c->stream));
    OCB_CUDA(c, D.rcCta.reserve((size_t)H.grid * kMasDof + 8, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(D.ints.p, ints.data(), ints.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(D.geom.p, geom.data(), geom.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(D.vinfo.p, H.vinfo.data(), H.vinfo.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaStreamSynchronize(c->stream));       // the staging vectors die with this scope
    MasView& W = D.view;
    W = MasView();
    W.L = H.L; W.Lloc = H.Lloc; W.grid = H.grid; W.topNodes = H.topNodes; W.maxLocalNodes = H.maxLocalNodes;
    W.vinfo = reinterpret_cast<const float4*>(D.vinfo.p);
    W.rcCta = D.rcCta.p;
    D.lvRowPtr.assign((size_t)H.L, nullptr); D.lvColIdx.assign((size_t)H.L, nullptr); D.lvVal.assign((size_t)H.L, nullptr);
    D.lvNnz.assign((size_t)H.L, 0);
    D.valTotal = valTot;
    D.groupTotal = 0;
    for (int l = 1; l <= H.L; ++l) {
        const Off& o = off[l - 1];
        MasLevel& V = W.lv[l - 1];
        V.nNodes = o.nNodes; V.nGroups = o.nGroups;
        V.childBeg = D.ints.p + o.childBeg;
        V.parent = D.ints.p + o.parent;
        V.ctaBeg = l <= H.Lloc ? D.ints.p + o.ctaBeg : nullptr;
        V.groupBeg = l < H.L ? D.ints.p + off[l].childBeg : D.ints.p + topGroupOff;
        V.geom = reinterpret_cast<const double4*>(D.geom.p + o.geom);
        V.inv = D.inv.p + o.inv;
        D.lvRowPtr[l - 1] = D.ints.p + o.rowPtr; D.lvColIdx[l - 1] = D.ints.p + o.colIdx; D.lvVal[l - 1] = D.val.p + o.val;
        D.lvNnz[l - 1] = o.nnz;
        D.groupTotal += o.nGroups;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device set-up
__device__ __forceinline__ int find_col(const int32_t* __restrict__ colIdx, int lo, int hi, int col)
{
    --hi;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1, cm = colIdx[mid];
        if (cm == col) return mid;
        if (cm < col) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

// A_1 = P_1^T A P_1 on the leaf adjacency: one thread per fine block row
__global__ void __launch_bounds__(256)
mas_galerkin_fine_kernel(int nRows, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, const double* __restrict__ val,
                         const float4* __restrict__ vinfo, const int32_t* __restrict__ rowPtr1, const int32_t* __restrict__ colIdx1,
                         double* __restrict__ val1)
{
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nRows; i += gridDim.x * 256) {
        const float4 vi = vinfo[i];
        if (vi.x == 0.0f) continue;
        const int a = __float_as_int(vi.w);
        const double fi[3] = {(double)vi.x, (double)vi.y, (double)vi.z};
        const int lo1 = rowPtr1[a], hi1 = rowPtr1[a + 1];
        for (int b = rowPtr[i]; b < rowPtr[i + 1]; ++b) {
            const int j = colIdx[b];
            const float4 vj = vinfo[j];
            if (vj.x == 0.0f) continue;
            const int s = find_col(colIdx1, lo1, hi1, __float_as_int(vj.w));
            if (s < 0) continue;
            const double fj[3] = {(double)vj.x, (double)vj.y, (double)vj.z};
            const double A[2][2] = {{val[4 * (size_t)b], val[4 * (size_t)b + 1]}, {val[4 * (size_t)b + 2], val[4 * (size_t)b + 3]}};
            double* o = val1 + 36 * (size_t)s;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                for (int qi = 0; qi < 3; ++qi)
#pragma unroll
                    for (int cj = 0; cj < 2; ++cj)
#pragma unroll
                        for (int qj = 0; qj < 3; ++qj)
                            atomicAdd(o + (ci * 3 + qi) * 6 + cj * 3 + qj, fi[qi] * A[ci][cj] * fj[qj]);
        }
    }
}

// A_{l+1} = R A_l R^T: one thread per node row of level l
__global__ void __launch_bounds__(128)
mas_coarsen_kernel(int nNodes, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, const double* __restrict__ val,
                   const int32_t* __restrict__ parent, const double4* __restrict__ geom, const double4* __restrict__ geomUp,
                   const int32_t* __restrict__ rowPtrUp, const int32_t* __restrict__ colIdxUp, double* __restrict__ valUp)
{
    for (int a = blockIdx.x * 128 + threadIdx.x; a < nNodes; a += gridDim.x * 128) {
        const int pa = parent[a];
        const double4 ga = geom[a], gpa = geomUp[pa];
        const double isa = 1.0 / gpa.z;
        const double Ra[3][3] = {{1.0, 0.0, 0.0}, {(ga.x - gpa.x) * isa, ga.z * isa, 0.0}, {(ga.y - gpa.y) * isa, 0.0, ga.z * isa}};
        const int lo = rowPtrUp[pa], hi = rowPtrUp[pa + 1];
        for (int blk = rowPtr[a]; blk < rowPtr[a + 1]; ++blk) {
            const int b = colIdx[blk];
            const int pb = parent[b];
            const int s = find_col(colIdxUp, lo, hi, pb);
            if (s < 0) continue;
            const double4 gb = geom[b], gpb = geomUp[pb];
            const double isb = 1.0 / gpb.z;
            const double Rb[3][3] = {{1.0, 0.0, 0.0}, {(gb.x - gpb.x) * isb, gb.z * isb, 0.0}, {(gb.y - gpb.y) * isb, 0.0, gb.z * isb}};
            const double* M = val + 36 * (size_t)blk;
            double* o = valUp + 36 * (size_t)s;
#pragma unroll
            for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                for (int cj = 0; cj < 2; ++cj) {
                    double S[3][3], T[3][3];
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) S[i][j] = M[(ci * 3 + i) * 6 + cj * 3 + j];
                    // T = S Rb^T
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        T[i][0] = S[i][0];
                        T[i][1] = S[i][0] * Rb[1][0] + S[i][1] * Rb[1][1];
                        T[i][2] = S[i][0] * Rb[2][0] + S[i][2] * Rb[2][2];
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        atomicAdd(o + (ci * 3 + 0) * 6 + cj * 3 + j, T[0][j]);
                        atomicAdd(o + (ci * 3 + 1) * 6 + cj * 3 + j, Ra[1][0] * T[0][j] + Ra[1][1] * T[1][j]);
                        atomicAdd(o + (ci * 3 + 2) * 6 + cj * 3 + j, Ra[2][0] * T[0][j] + Ra[2][2] * T[2][j]);
                    }
                }
        }
    }
}

// group blocks D_l[g] (<= 48x48) gathered from A_l and inverted in shared memory (Gauss-Jordan, SPD, no pivoting;
// a DOF whose pivot collapses -- a node of fixed vertices only, collinear vertices -- is dropped); all levels in
// one launch: CTA -> (level, group) through the prefix table.
struct MasInvertArgs {
    int L;
    int groupPre[kMasMaxLevels + 1];
    const int32_t* rowPtr[kMasMaxLevels]; const int32_t* colIdx[kMasMaxLevels]; const double* val[kMasMaxLevels];
    const int32_t* parent[kMasMaxLevels]; const int32_t* groupBeg[kMasMaxLevels]; float* inv[kMasMaxLevels];
};
__global__ void __launch_bounds__(256)
mas_invert_kernel(MasInvertArgs P)
{
    __shared__ double D[kMasBlk][kMasBlk + 1];
    __shared__ double d0[kMasBlk];
    __shared__ int dead[kMasBlk];
    int l = 0;
    while (l + 1 < P.L && (int)blockIdx.x >= P.groupPre[l + 1]) ++l;
    const int g = blockIdx.x - P.groupPre[l];
    const int gb = P.groupBeg[l][g], nch = P.groupBeg[l][g + 1] - gb;
    const int nd = nch * kMasDof;
    for (int e = threadIdx.x; e < kMasBlk * kMasBlk; e += 256) D[e / kMasBlk][e % kMasBlk] = 0.0;
    __syncthreads();
    for (int sa = 0; sa < nch; ++sa) {
        const int a = gb + sa;
        const int b0 = P.rowPtr[l][a], nb = P.rowPtr[l][a + 1] - b0;
        for (int e = threadIdx.x; e < nb * 36; e += 256) {
            const int blk = b0 + e / 36, ij = e % 36;
            const int b = P.colIdx[l][blk];
            if (P.parent[l][b] != g) continue;
            D[sa * kMasDof + ij / 6][(b - gb) * kMasDof + ij % 6] = P.val[l][36 * (size_t)blk + ij];
        }
    }
    __syncthreads();
    if (threadIdx.x < kMasBlk) {
        const int k = threadIdx.x;
        if (k >= nd) D[k][k] = 1.0;
        d0[k] = D[k][k];
        dead[k] = 0;
    }
    __syncthreads();
    for (int k = 0; k < nd; ++k) {
        const double p = D[k][k];
        const bool bad = !(d0[k] > 0.0) || !(p > 1e-10 * d0[k]);
        __syncthreads();
        if (bad) {
            if (threadIdx.x < kMasBlk) { D[k][threadIdx.x] = 0.0; D[threadIdx.x][k] = 0.0; }
            if (threadIdx.x == 0) dead[k] = 1;
            __syncthreads();
            continue;
        }
        const double ip = 1.0 / p;
        if (threadIdx.x < kMasBlk && threadIdx.x != k) D[k][threadIdx.x] *= ip;
        __syncthreads();
        for (int e = threadIdx.x; e < nd * nd; e += 256) {
            const int i = e / nd, j = e % nd;
            if (i != k && j != k) D[i][j] -= D[i][k] * D[k][j];
        }
        __syncthreads();
        if (threadIdx.x < kMasBlk) {
            if (threadIdx.x != k) D[threadIdx.x][k] *= -ip;
            else D[k][k] = ip;
        }
        __syncthreads();
    }
    float* out = P.inv[l] + (size_t)g * kMasBlk * kMasBlk;
    for (int e = threadIdx.x; e < kMasBlk * kMasBlk; e += 256) {
        const int i = e / kMasBlk, j = e % kMasBlk;
        const bool ok = i < nd && j < nd && !dead[i] && !dead[j];
        out[e] = ok ? (float)(0.5 * (D[i][j] + D[j][i])) : 0.0f;
    }
}

int launch_mas_setup(ocb_ctx* c)
{
    MasHost& H = c->masH;
    MasDev& D = c->masD;
    if (!H.enabled) return 0;
    ProfScope prof(c, K_MAS_SETUP);
    OCB_CUDA(c, cudaMemsetAsync(D.val.p, 0, D.valTotal * sizeof(double), c->stream));
    const int n = c->nVtot;
    int grid = (n + 255) / 256; if (grid > c->numSMs * 8) grid = c->numSMs * 8; if (grid < 1) grid = 1;
    mas_galerkin_fine_kernel<<<grid, 256, 0, c->stream>>>(n, c->rowPtr.p, c->colIdx.p, c->val.p, D.view.vinfo, D.lvRowPtr[0], D.lvColIdx[0], D.lvVal[0]);
    KCHECK(c);
    for (int l = 1; l < H.L; ++l) {
        const MasLevel& V = D.view.lv[l - 1];
        int g = (V.nNodes + 127) / 128; if (g > c->numSMs * 8) g = c->numSMs * 8; if (g < 1) g = 1;
        mas_coarsen_kernel<<<g, 128, 0, c->stream>>>(V.nNodes, D.lvRowPtr[l - 1], D.lvColIdx[l - 1], D.lvVal[l - 1], V.parent, V.geom,
                                                    D.view.lv[l].geom, D.lvRowPtr[l], D.lvColIdx[l], D.lvVal[l]);
        KCHECK(c);
    }
    MasInvertArgs A;
    A.L = H.L;
    A.groupPre[0] = 0;
    for (int l = 1; l <= H.L; ++l) {
        const MasLevel& V = D.view.lv[l - 1];
        A.groupPre[l] = A.groupPre[l - 1] + V.nGroups;
        A.rowPtr[l - 1] = D.lvRowPtr[l - 1]; A.colIdx[l - 1] = D.lvColIdx[l - 1]; A.val[l - 1] = D.lvVal[l - 1];
        A.parent[l - 1] = V.parent; A.groupBeg[l - 1] = V.groupBeg; A.inv[l - 1] = const_cast<float*>(V.inv);
    }
    mas_invert_kernel<<<A.groupPre[H.L], 256, 0, c->stream>>>(A);
    KCHECK(c);
    return 0;
}

}  // namespace ocb
