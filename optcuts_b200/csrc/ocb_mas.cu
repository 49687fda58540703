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
struct RcbPt { double x, y; int32_t id; };
struct Rcb {
    RcbPt* pt;
    // bounding box handed down the recursion (the cut value closes the children's boxes): no extra pass per level
    void split(int beg, int end, const int* pre, int nParts, double lox, double hix, double loy, double hiy) const      // pre: prefix sums of the part sizes (nParts + 1)
    {
        if (nParts <= 1 || end - beg <= 1) return;
        const int h = nParts / 2;
        const int nl = pre[h] - pre[0];
        double cut = 0.0; bool alongY = false, didCut = false;
        if (nl > 0 && nl < end - beg) {
            alongY = (hiy - loy) > (hix - lox);
            if (alongY) std::nth_element(pt + beg, pt + beg + nl, pt + end, [](const RcbPt& a, const RcbPt& b) { return a.y < b.y || (a.y == b.y && a.id < b.id); });
            else        std::nth_element(pt + beg, pt + beg + nl, pt + end, [](const RcbPt& a, const RcbPt& b) { return a.x < b.x || (a.x == b.x && a.id < b.id); });
            cut = alongY ? pt[beg + nl].y : pt[beg + nl].x;
            didCut = true;
        }
        if (didCut && alongY) { split(beg, beg + nl, pre, h, lox, hix, loy, cut); split(beg + nl, end, pre + h, nParts - h, lox, hix, cut, hiy); }
        else if (didCut)      { split(beg, beg + nl, pre, h, lox, cut, loy, hiy); split(beg + nl, end, pre + h, nParts - h, cut, hix, loy, hiy); }
        else                  { split(beg, beg + nl, pre, h, lox, hix, loy, hiy); split(beg + nl, end, pre + h, nParts - h, lox, hix, loy, hiy); }
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
    HostTimer* _t1 = new HostTimer("    h:rcb");
    std::vector<RcbPt> pts((size_t)n);
    double blo[2] = {1e300, 1e300}, bhi[2] = {-1e300, -1e300};
    // start from the previous solver order when there is one: between Newton iterations the vertices barely move, and
    // selecting on nearly-partitioned data costs a fraction of the swaps
    {
        std::vector<uint8_t> seen((size_t)n, 0);
        int w = 0;
        for (size_t r = 0; r < c->hVertOf.size(); ++r) { const int v = c->hVertOf[r]; if (v >= 0 && v < n && !seen[v]) { seen[v] = 1; pts[w++].id = v; } }
        for (int v = 0; v < n; ++v) if (!seen[v]) pts[w++].id = v;
    }
    for (int i = 0; i < n; ++i) {
        const int v = pts[i].id;
        pts[i].x = xy[2 * (size_t)v]; pts[i].y = xy[2 * (size_t)v + 1];
        blo[0] = std::min(blo[0], pts[i].x); bhi[0] = std::max(bhi[0], pts[i].x); blo[1] = std::min(blo[1], pts[i].y); bhi[1] = std::max(bhi[1], pts[i].y);
    }
    Rcb R{pts.data()};
    // two-stage: CTA chunks first, then the leaves of every chunk; the chunk boxes are recomputed (cheap, once)
    R.split(0, n, ctaPre.data(), nonEmpty, blo[0], bhi[0], blo[1], bhi[1]);
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
                double lo2[2] = {1e300, 1e300}, hi2[2] = {-1e300, -1e300};
                for (int i = beg; i < beg + m; ++i) {
                    lo2[0] = std::min(lo2[0], pts[i].x); hi2[0] = std::max(hi2[0], pts[i].x); lo2[1] = std::min(lo2[1], pts[i].y); hi2[1] = std::max(hi2[1], pts[i].y);
                }
                R.split(beg, beg + m, pre.data(), nl, lo2[0], hi2[0], lo2[1], hi2[1]);
                for (int k = 1; k <= nl; ++k) L1.childBeg.push_back(beg + pre[k]);
            }
            L1.ctaBeg[b + 1] = (int32_t)L1.childBeg.size() - 1;
        }
    }
    delete _t1;
    HostTimer _t2("    h:levels");
    c->hVertOf.resize((size_t)n);
    c->hRowOf.assign((size_t)n, 0);
    for (int r = 0; r < n; ++r) { c->hVertOf[r] = pts[r].id; c->hRowOf[pts[r].id] = r; }
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
    if (!H.enabled) { D.view = MasView(); D.view.L = 0; return 0; }
    const int n = c->nVtot;
    std::vector<int32_t> nodeOf((size_t)n);       // row -> node of the current level
    for (int r = 0; r < n; ++r) { int32_t id; std::memcpy(&id, &H.vinfo[4 * (size_t)r + 3], 4); nodeOf[r] = id; }
    const std::vector<int32_t>* fineRowPtr = &c->hSRowPtr; const std::vector<int32_t>* fineColIdx = &c->hSColIdx;
    std::vector<int32_t> stamp;
    for (int l = 1; l <= H.L; ++l) {
        MasHost::Level& V = H.lv[l - 1];
        const int nn = (int)V.childBeg.size() - 1;
        V.rowPtr.assign((size_t)nn + 1, 0); V.colIdx.clear();
        V.colIdx.reserve((size_t)nn * 10);
        stamp.assign((size_t)nn, -1);
        for (int k = 0; k < nn; ++k) {
            const size_t w0 = V.colIdx.size();
            V.colIdx.push_back(k); stamp[k] = k;
            for (int ch = V.childBeg[k]; ch < V.childBeg[k + 1]; ++ch)
                for (int b = (*fineRowPtr)[ch]; b < (*fineRowPtr)[ch + 1]; ++b) {
                    const int q = nodeOf[(*fineColIdx)[b]];
                    if (stamp[q] != k) { stamp[q] = k; V.colIdx.push_back(q); }
                }
            int32_t* a = V.colIdx.data() + w0;
            const int m = (int)(V.colIdx.size() - w0);
            for (int i = 1; i < m; ++i) { const int32_t v = a[i]; int j = i - 1; while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; } a[j + 1] = v; }
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
    D.lv.assign((size_t)H.L, MasLevel());
    D.lvRowPtr.assign((size_t)H.L, nullptr); D.lvColIdx.assign((size_t)H.L, nullptr); D.lvVal.assign((size_t)H.L, nullptr);
    D.lvNnz.assign((size_t)H.L, 0);
    D.valTotal = valTot;
    D.groupTotal = 0;
    for (int l = 1; l <= H.L; ++l) {
        const Off& o = off[l - 1];
        MasLevel& V = D.lv[l - 1];
        V.nNodes = o.nNodes; V.nGroups = o.nGroups;
        V.childBeg = D.ints.p + o.childBeg;
        V.parent = D.ints.p + o.parent;
        V.groupBeg = l < H.L ? D.ints.p + off[l].childBeg : D.ints.p + topGroupOff;
        V.geom = reinterpret_cast<const double4*>(D.geom.p + o.geom);
        V.inv = D.inv.p + o.inv;
        D.lvRowPtr[l - 1] = D.ints.p + o.rowPtr; D.lvColIdx[l - 1] = D.ints.p + o.colIdx; D.lvVal[l - 1] = D.val.p + o.val;
        D.lvNnz[l - 1] = o.nnz;
        D.groupTotal += o.nGroups;
    }
    // ---- the apply's tables (CTA-local indices, see MasView)
    const int grid = H.grid, Lloc = H.Lloc, L = H.L, nCh = L - Lloc + 1;
    const int rowsPer = (n + grid - 1) / grid;
    auto nodesOf = [&](int l) { return (int)H.lv[l - 1].childBeg.size() - 1; };
    auto groupBegOf = [&](int l, int g) { return l < L ? H.lv[l].childBeg[g] : (g == 0 ? 0 : nodesOf(L)); };
    auto xfer = [&](int l, int k, double* X) {         // node k of level l -> its parent (level l + 1)
        X[0] = X[1] = 0.0; X[2] = 1.0; X[3] = 0.0;
        if (l >= L) return;
        const double* gc = H.lv[l - 1].geom.data() + 4 * (size_t)k;
        const double* gp = H.lv[l].geom.data() + 4 * (size_t)H.lv[l - 1].parent[k];
        X[0] = (gc[0] - gp[0]) / gp[2]; X[1] = (gc[1] - gp[1]) / gp[2]; X[2] = gc[2] / gp[2];
    };
    const int nCtaNodes = nodesOf(Lloc);
    std::vector<int32_t> tI; std::vector<double> tD;
    std::vector<int32_t> ctaNodeOff((size_t)grid + 1, 0), ctaSolve((size_t)grid, 0), ctaLvOff((size_t)grid * (kMasMaxLevels + 1), 0), ctaLeafBeg(H.lv[0].ctaBeg);
    std::vector<int32_t> nodeA, nodeB; std::vector<double> nodeX;
    { size_t tot = 0; for (int l = 1; l <= Lloc; ++l) tot += (size_t)nodesOf(l); nodeA.reserve(4 * tot); nodeB.reserve(4 * tot); nodeX.reserve(4 * tot); }
    for (int b = 0; b < grid; ++b) {
        int32_t* lvOff = ctaLvOff.data() + (size_t)b * (kMasMaxLevels + 1);
        int o = 0;
        for (int l = 1; l <= Lloc; ++l) { lvOff[l - 1] = o; o += H.lv[l - 1].ctaBeg[b + 1] - H.lv[l - 1].ctaBeg[b]; }
        for (int l = Lloc; l <= kMasMaxLevels; ++l) lvOff[l] = o;
        ctaNodeOff[b + 1] = ctaNodeOff[b] + o;
        ctaSolve[b] = lvOff[Lloc - 1];
        const int rowBeg = std::min(n, b * rowsPer);
        for (int l = 1; l <= Lloc; ++l) {
            const MasHost::Level& V = H.lv[l - 1];
            const int n0 = V.ctaBeg[b];
            for (int k = n0; k < V.ctaBeg[b + 1]; ++k) {
                int32_t A[4] = {0, 0, 0, 0}, B[4] = {0, 0, 0, l};
                if (l < Lloc) {
                    const int g = V.parent[k], gb = groupBegOf(l, g), ge = groupBegOf(l, g + 1);
                    A[0] = lvOff[l - 1] + (gb - n0); A[1] = kMasDof * (ge - gb); A[2] = kMasDof * (k - gb);
                    A[3] = lvOff[l] + (g - H.lv[l].ctaBeg[b]);
                    B[0] = (int32_t)(off[l - 1].inv + (size_t)g * kMasBlk * kMasBlk);
                }
                B[1] = l == 1 ? V.childBeg[k] - rowBeg : lvOff[l - 2] + (V.childBeg[k] - H.lv[l - 2].ctaBeg[b]);
                B[2] = V.childBeg[k + 1] - V.childBeg[k];
                nodeA.insert(nodeA.end(), A, A + 4); nodeB.insert(nodeB.end(), B, B + 4);
                double X[4]; xfer(l, k, X);
                nodeX.insert(nodeX.end(), X, X + 4);
            }
        }
    }
    std::vector<int32_t> topLevelOff((size_t)nCh + 1, 0);
    for (int j = 0; j < nCh; ++j) topLevelOff[j + 1] = topLevelOff[j] + nodesOf(Lloc + j);
    const int topNodes = topLevelOff[nCh];
    std::vector<int32_t> topUp(2 * (size_t)topNodes, 0); std::vector<double> topX(4 * (size_t)topNodes, 0.0);
    for (int j = 0; j < nCh; ++j) {
        const int l = Lloc + j;
        for (int k = 0; k < nodesOf(l); ++k) {
            const int ti = topLevelOff[j] + k;
            if (j > 0) { topUp[2 * ti] = topLevelOff[j - 1] + H.lv[l - 1].childBeg[k]; topUp[2 * ti + 1] = H.lv[l - 1].childBeg[k + 1] - H.lv[l - 1].childBeg[k]; }
            xfer(l, k, topX.data() + 4 * (size_t)ti);
        }
    }
    std::vector<int32_t> chainM((size_t)grid * kMasMaxLevels * 4, 0);
    for (int b = 0; b < nCtaNodes; ++b) {
        int anc[kMasMaxLevels + 2];
        anc[Lloc] = b;
        for (int l = Lloc; l < L; ++l) anc[l + 1] = H.lv[l - 1].parent[anc[l]];
        for (int j = 0; j < nCh; ++j) {
            const int l = L - j, a = anc[l], g = H.lv[l - 1].parent[a], gb = groupBegOf(l, g), ge = groupBegOf(l, g + 1);
            int32_t* C = chainM.data() + ((size_t)b * kMasMaxLevels + j) * 4;
            C[0] = topLevelOff[l - Lloc] + gb; C[1] = kMasDof * (ge - gb);
            C[2] = (int32_t)(off[l - 1].inv + (size_t)g * kMasBlk * kMasBlk + (size_t)(a - gb) * kMasDof * kMasBlk);
            C[3] = topLevelOff[l - Lloc] + a;
        }
    }
    auto putI = [&tI](const std::vector<int32_t>& v) { size_t o = tI.size(); tI.insert(tI.end(), v.begin(), v.end()); while (tI.size() & 3) tI.push_back(0); return o; };
    auto putD = [&tD](const std::vector<double>& v) { size_t o = tD.size(); tD.insert(tD.end(), v.begin(), v.end()); while (tD.size() & 3) tD.push_back(0.0); return o; };
    const size_t oNodeOff = putI(ctaNodeOff), oSolve = putI(ctaSolve), oLvOff = putI(ctaLvOff), oLeafBeg = putI(ctaLeafBeg), oNodeA = putI(nodeA),
                 oNodeB = putI(nodeB), oTopUp = putI(topUp), oTopLevelOff = putI(topLevelOff), oChainM = putI(chainM);
    const size_t oNodeX = putD(nodeX), oTopX = putD(topX);
    OCB_CUDA(c, D.tabI.reserve(tI.size() + 4, c->stream));
    OCB_CUDA(c, D.tabD.reserve(tD.size() + 4, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(D.tabI.p, tI.data(), tI.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(D.tabD.p, tD.data(), tD.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    // (pageable sources: cudaMemcpyAsync returns once the data is staged, the vectors may go out of scope)
    MasView& W = D.view;
    W = MasView();
    W.L = L; W.Lloc = Lloc; W.grid = grid; W.nCh = nCh; W.topNodes = topNodes; W.nCtaNodes = nCtaNodes;
    W.maxLocalNodes = H.maxLocalNodes; W.rowsPer = rowsPer;
    W.ctaNodeOff = D.tabI.p + oNodeOff; W.ctaSolve = D.tabI.p + oSolve; W.ctaLvOff = D.tabI.p + oLvOff; W.ctaLeafBeg = D.tabI.p + oLeafBeg;
    W.nodeA = reinterpret_cast<const int4*>(D.tabI.p + oNodeA); W.nodeB = reinterpret_cast<const int4*>(D.tabI.p + oNodeB);
    W.nodeX = reinterpret_cast<const double4*>(D.tabD.p + oNodeX);
    W.topUp = reinterpret_cast<const int2*>(D.tabI.p + oTopUp); W.topX = reinterpret_cast<const double4*>(D.tabD.p + oTopX);
    W.topLevelOff = D.tabI.p + oTopLevelOff; W.chainM = reinterpret_cast<const int4*>(D.tabI.p + oChainM);
    W.inv = D.inv.p; W.vinfo = reinterpret_cast<const float4*>(D.vinfo.p); W.rcCta = D.rcCta.p;
    H.topNodes = topNodes;
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
    mas_galerkin_fine_kernel<<<grid, 256, 0, c->stream>>>(n, c->rowPtr.p, c->colIdx.p, c->val.p, reinterpret_cast<const float4*>(D.vinfo.p), D.lvRowPtr[0], D.lvColIdx[0], D.lvVal[0]);
    KCHECK(c);
    for (int l = 1; l < H.L; ++l) {
        const MasLevel& V = D.lv[l - 1];
        int g = (V.nNodes + 127) / 128; if (g > c->numSMs * 8) g = c->numSMs * 8; if (g < 1) g = 1;
        mas_coarsen_kernel<<<g, 128, 0, c->stream>>>(V.nNodes, D.lvRowPtr[l - 1], D.lvColIdx[l - 1], D.lvVal[l - 1], V.parent, V.geom,
                                                    D.lv[l].geom, D.lvRowPtr[l], D.lvColIdx[l], D.lvVal[l]);
        KCHECK(c);
    }
    MasInvertArgs A;
    A.L = H.L;
    A.groupPre[0] = 0;
    for (int l = 1; l <= H.L; ++l) {
        const MasLevel& V = D.lv[l - 1];
        A.groupPre[l] = A.groupPre[l - 1] + V.nGroups;
        A.rowPtr[l - 1] = D.lvRowPtr[l - 1]; A.colIdx[l - 1] = D.lvColIdx[l - 1]; A.val[l - 1] = D.lvVal[l - 1];
        A.parent[l - 1] = V.parent; A.groupBeg[l - 1] = V.groupBeg; A.inv[l - 1] = V.inv;
    }
    mas_invert_kernel<<<A.groupPre[H.L], 256, 0, c->stream>>>(A);
    KCHECK(c);
    return 0;
}

}  // namespace ocb
