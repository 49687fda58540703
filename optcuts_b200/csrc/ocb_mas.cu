// optcuts_b200 — multilevel additive Schwarz preconditioner: hierarchy construction (host, at pattern time) and
// the per-factorisation set-up kernels (Galerkin coarse matrices, group-block inversion).  See ocb_mas.cuh for the
// operator and the apply; this replaces the numeric factorisation of Eigen::SimplicialLDLT
// (EigenLibSolver.cpp:80-93) as the per-Newton-iteration solver set-up.
#include "ocb_internal.cuh"
#include "ocb_mas.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace ocb {

#define KCHECK(c) do { (c)->launches++; cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) return cuda_fail((c), _e, __func__); } while (0)

// ------------------------------------------------------------------------------------------------
// host: order of the points along a Hilbert curve (20 bits per axis).  O(n): table-driven curve index (4 bits of x and
// y per step) + LSD radix sort.  Consecutive runs along the curve are compact patches, so CTA chunks, leaves and
// groups are all plain consecutive ranges; measured against recursive coordinate bisection (tools/mas_proto.py
// --curve): 154 vs 174 CG iterations at 10k faces, 282 vs 286 at 160k, at 1/4 (10k) to 1/15 (1M) of the host time.
namespace {
struct HilbertTable {
    uint16_t t[4][256];         // [state][x4 << 4 | y4] -> digits (8 bits) | next state << 8
    HilbertTable()
    {
        for (int st = 0; st < 4; ++st)
            for (int xy = 0; xy < 256; ++xy) {
                int swap = st & 1, compl_ = st >> 1, d = 0;
                for (int b = 3; b >= 0; --b) {
                    int xb = ((xy >> 4) >> b) & 1, yb = ((xy & 15) >> b) & 1;
                    xb ^= compl_; yb ^= compl_;
                    if (swap) { const int tmp = xb; xb = yb; yb = tmp; }
                    d = (d << 2) | ((3 * xb) ^ yb);
                    if (yb == 0) { compl_ ^= xb; swap ^= 1; }
                }
                t[st][xy] = (uint16_t)(d | ((swap | (compl_ << 1)) << 8));
            }
    }
};
static inline uint64_t hilbert_index20(const HilbertTable& H, uint32_t x, uint32_t y)     // x, y < 2^20
{
    uint64_t d = 0;
    int st = 0;
    for (int sh = 16; sh >= 0; sh -= 4) {
        const uint16_t e = H.t[st][(((x >> sh) & 15) << 4) | ((y >> sh) & 15)];
        d = (d << 8) | (e & 255);
        st = e >> 8;
    }
    return d;
}
// order[] = point ids sorted by curve index (ties by id); keys are packed with the id, 4 radix passes of 11 bits
static void hilbert_sort(const double* xy, int n, std::vector<int32_t>& order)
{
    static const HilbertTable H;
    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 2; ++a) { lo[a] = std::min(lo[a], xy[2 * (size_t)i + a]); hi[a] = std::max(hi[a], xy[2 * (size_t)i + a]); }
    const double ext = std::max(hi[0] - lo[0], hi[1] - lo[1]);
    const double scale = ext > 0.0 ? 1048575.0 / ext : 0.0;
    std::vector<uint64_t> a((size_t)n), b((size_t)n);
    for (int i = 0; i < n; ++i) {
        const uint32_t x = (uint32_t)((xy[2 * (size_t)i] - lo[0]) * scale), y = (uint32_t)((xy[2 * (size_t)i + 1] - lo[1]) * scale);
        a[i] = (hilbert_index20(H, x > 1048575u ? 1048575u : x, y > 1048575u ? 1048575u : y) << 24) | (uint64_t)i;      // n < 2^24 (checked by the caller)
    }
    uint32_t cnt[2048];
    for (int pass = 0; pass < 4; ++pass) {
        const int sh = 24 + 11 * pass;                             // 40-bit curve index above the 24-bit id: digits at bits 24, 35, 46, 57
        std::memset(cnt, 0, sizeof(cnt));
        for (int i = 0; i < n; ++i) ++cnt[(a[i] >> sh) & 2047];
        uint32_t run = 0;
        for (int k = 0; k < 2048; ++k) { const uint32_t c = cnt[k]; cnt[k] = run; run += c; }
        for (int i = 0; i < n; ++i) b[cnt[(a[i] >> sh) & 2047]++] = a[i];
        a.swap(b);
    }
    order.resize((size_t)n);
    for (int i = 0; i < n; ++i) order[i] = (int32_t)(a[i] & 0xFFFFFFu);
}

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
    HostTimer* _t1 = new HostTimer("    h:curve");
    if (n >= (1 << 24)) return set_err(c, OCB_ERR_ARG, "MAS hierarchy: more than 2^24 vertices");
    std::vector<int32_t> order;
    hilbert_sort(xy.data(), n, order);
    // CTA chunks and leaves are consecutive runs of the curve order
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
                // leaves of EXACTLY 8 rows (the last one of a CTA shorter): local row lr belongs to local leaf lr / 8, so the
                // PCG kernel restricts the residual with an 8-lane shuffle reduction right where it updates it
                const int nl = (m + kMasLeaf - 1) / kMasLeaf;
                for (int k = 1; k <= nl; ++k) L1.childBeg.push_back(beg + std::min(m, k * kMasLeaf));
            }
            L1.ctaBeg[b + 1] = (int32_t)L1.childBeg.size() - 1;
        }
    }
    delete _t1;
    HostTimer _t2("    h:levels");
    c->hVertOf.resize((size_t)n);
    c->hRowOf.assign((size_t)n, 0);
    for (int r = 0; r < n; ++r) { c->hVertOf[r] = order[r]; c->hRowOf[order[r]] = r; }
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
                const double m = c->hFixed[v] ? 0.0 : 1.0;
                float* o = H.vinfo.data() + 4 * (size_t)r;
                // packed like the inverses (high word of the fp64 value, see mas_pack): no conversion instruction per use
                auto pack = [](double v) { uint64_t b; std::memcpy(&b, &v, 8); b += 0x80000000ull; const uint32_t h = (uint32_t)(b >> 32); float f; std::memcpy(&f, &h, 4); return f; };
                o[0] = pack(m); o[1] = pack(m * (p[0] - g[0]) / g[2]); o[2] = pack(m * (p[1] - g[1]) / g[2]);
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
    // coarse level: the first level small enough for the exact dense inverse; everything above it is dropped
    static const int coarseEnv = []() { const char* e = getenv("OCB_MAS_COARSE_MAX"); return e ? atoi(e) : 0; }();
    const int coarseMax = coarseEnv >= kMasDof ? std::min(coarseEnv, kMasCoarseMax) : mas_coarse_cap(n);
    int Lc = H.Lloc;
    for (int l = 1; l <= H.Lloc; ++l) if (((int)H.lv[l - 1].childBeg.size() - 1) * kMasDof <= coarseMax) { Lc = l; break; }
    if (((int)H.lv[Lc - 1].childBeg.size() - 1) * kMasDof > kMasCoarseMax) return set_err(c, OCB_ERR_STATE, "MAS hierarchy: coarse level too large");
    H.lv.resize((size_t)Lc);
    H.L = H.Lloc = Lc;
    H.lv[Lc - 1].parent.assign(H.lv[Lc - 1].childBeg.size() - 1, 0);
    H.topNodes = 0;
    H.maxLocalNodes = 0;
    for (int b = 0; b < grid; ++b) {
        int s = 0;
        for (int k = 1; k <= H.L; ++k) s += H.lv[k - 1].ctaBeg[b + 1] - H.lv[k - 1].ctaBeg[b];
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
            V.colIdx.push_back(k); stamp[k] = k;
            for (int ch = V.childBeg[k]; ch < V.childBeg[k + 1]; ++ch)
                for (int b = (*fineRowPtr)[ch]; b < (*fineRowPtr)[ch + 1]; ++b) {
                    const int q = nodeOf[(*fineColIdx)[b]];
                    if (stamp[q] != k) { stamp[q] = k; V.colIdx.push_back(q); }
                }
            V.rowPtr[k + 1] = (int32_t)V.colIdx.size();      // rows stay unsorted: every device look-up is a linear scan
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
        o.nGroups = l < H.L ? (int)H.lv[l].childBeg.size() - 1 : 0;
        o.nnz = (int)V.colIdx.size();
        o.childBeg = put(V.childBeg); o.parent = put(V.parent);
        o.ctaBeg = l <= H.Lloc ? put(V.ctaBeg) : 0;
        o.rowPtr = put(V.rowPtr); o.colIdx = put(V.colIdx);
        o.geom = geom.size(); geom.insert(geom.end(), V.geom.begin(), V.geom.end());
        o.val = valTot; valTot += 36 * (size_t)o.nnz;
        o.inv = invTot; invTot += (size_t)kMasBlk * kMasBlk * o.nGroups;
    }
    OCB_CUDA(c, D.ints.reserve(ints.size() + 4, c->stream));
    OCB_CUDA(c, D.geom.reserve(geom.size() + 4, c->stream));
    OCB_CUDA(c, D.val.reserve(valTot + 4, c->stream));
    OCB_CUDA(c, D.inv.reserve(invTot + 4, c->stream));
    OCB_CUDA(c, D.vinfo.reserve(4 * (size_t)n + 4, c->stream));
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
        V.groupBeg = l < H.L ? D.ints.p + off[l].childBeg : nullptr;
        V.geom = reinterpret_cast<const double4*>(D.geom.p + o.geom);
        V.inv = D.inv.p + o.inv;
        D.lvRowPtr[l - 1] = D.ints.p + o.rowPtr; D.lvColIdx[l - 1] = D.ints.p + o.colIdx; D.lvVal[l - 1] = D.val.p + o.val;
        D.lvNnz[l - 1] = o.nnz;
        D.groupTotal += o.nGroups;
    }
    // ---- the apply's tables (CTA-local indices, see MasView)
    const int grid = H.grid, L = H.L;
    const int rowsPer = (n + grid - 1) / grid;
    auto nodesOf = [&](int l) { return (int)H.lv[l - 1].childBeg.size() - 1; };
    auto xfer = [&](int l, int k, double* X) {         // node k of level l -> its parent (level l + 1)
        X[0] = X[1] = 0.0; X[2] = 1.0; X[3] = 0.0;
        if (l >= L) return;
        const double* gc = H.lv[l - 1].geom.data() + 4 * (size_t)k;
        const double* gp = H.lv[l].geom.data() + 4 * (size_t)H.lv[l - 1].parent[k];
        X[0] = (gc[0] - gp[0]) / gp[2]; X[1] = (gc[1] - gp[1]) / gp[2]; X[2] = gc[2] / gp[2];
    };
    std::vector<int32_t> tI; std::vector<double> tD;
    std::vector<int32_t> ctaNodeOff((size_t)grid + 1, 0), ctaSolve((size_t)grid, 0), ctaLvOff((size_t)grid * (kMasMaxLevels + 1), 0);
    const std::vector<int32_t>& ctaLeafBeg = H.lv[0].ctaBeg;
    const std::vector<int32_t>& ctaCBeg = H.lv[L - 1].ctaBeg;
    std::vector<int32_t> nodeA, nodeB; std::vector<double> nodeX;
    { size_t tot = 0; for (int l = 1; l <= L; ++l) tot += (size_t)nodesOf(l); nodeA.reserve(4 * tot); nodeB.reserve(4 * tot); nodeX.reserve(4 * tot); }
    for (int b = 0; b < grid; ++b) {
        int32_t* lvOff = ctaLvOff.data() + (size_t)b * (kMasMaxLevels + 1);
        int o = 0;
        for (int l = 1; l <= L; ++l) { lvOff[l - 1] = o; o += H.lv[l - 1].ctaBeg[b + 1] - H.lv[l - 1].ctaBeg[b]; }
        for (int l = L; l <= kMasMaxLevels; ++l) lvOff[l] = o;
        ctaNodeOff[b + 1] = ctaNodeOff[b] + o;
        ctaSolve[b] = lvOff[L - 1];
        const int rowBeg = std::min(n, b * rowsPer);
        for (int l = 1; l <= L; ++l) {
            const MasHost::Level& V = H.lv[l - 1];
            const int n0 = V.ctaBeg[b];
            for (int k = n0; k < V.ctaBeg[b + 1]; ++k) {
                int32_t A[4] = {0, 0, 0, 0}, B[4] = {0, 0, 0, l};
                if (l < L) {
                    const int g = V.parent[k], gb = H.lv[l].childBeg[g], ge = H.lv[l].childBeg[g + 1];
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
    auto putI = [&tI](const std::vector<int32_t>& v) { size_t o = tI.size(); tI.insert(tI.end(), v.begin(), v.end()); while (tI.size() & 3) tI.push_back(0); return o; };
    auto putD = [&tD](const std::vector<double>& v) { size_t o = tD.size(); tD.insert(tD.end(), v.begin(), v.end()); while (tD.size() & 3) tD.push_back(0.0); return o; };
    const size_t oNodeOff = putI(ctaNodeOff), oSolve = putI(ctaSolve), oLvOff = putI(ctaLvOff), oLeafBeg = putI(ctaLeafBeg), oCBeg = putI(ctaCBeg),
                 oNodeA = putI(nodeA), oNodeB = putI(nodeB);
    const size_t oNodeX = putD(nodeX);
    // dense coarse system: X / Y ping-pong (fp64), the fp32 inverse, the original diagonal, two pivot-block buffers
    const int nC = kMasDof * nodesOf(L);
    const int ldC = (nC + kMasCoarseBlk - 1) / kMasCoarseBlk * kMasCoarseBlk;
    OCB_CUDA(c, D.tabI.reserve(tI.size() + 4, c->stream));
    OCB_CUDA(c, D.tabD.reserve(tD.size() + 4, c->stream));
    OCB_CUDA(c, D.dense.reserve(2 * (size_t)ldC * ldC + 2 * (size_t)ldC + 2 * kMasCoarseBlk * kMasCoarseBlk + 16 + kMasCoarseMax / kMasCoarseBlk, c->stream));
    OCB_CUDA(c, D.cinv.reserve((size_t)ldC * ldC + 8, c->stream));
    OCB_CUDA(c, D.rcCta.reserve((size_t)ldC + 8, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(D.tabI.p, tI.data(), tI.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    OCB_CUDA(c, cudaMemcpyAsync(D.tabD.p, tD.data(), tD.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    // (pageable sources: cudaMemcpyAsync returns once the data is staged, the vectors may go out of scope)
    MasView& W = D.view;
    W = MasView();
    W.L = L; W.nC = nC; W.ldC = ldC;
    W.maxLocalNodes = H.maxLocalNodes; W.rowsPer = rowsPer;
    W.maxOwnC = 0;
    for (int b = 0; b < grid; ++b) W.maxOwnC = std::max(W.maxOwnC, (int)(ctaCBeg[b + 1] - ctaCBeg[b]));
    W.cinvInSmem = 0;                                  // decided per launch (it depends on the kernel's shared-memory mode)
    W.ctaNodeOff = D.tabI.p + oNodeOff; W.ctaSolve = D.tabI.p + oSolve; W.ctaLvOff = D.tabI.p + oLvOff; W.ctaLeafBeg = D.tabI.p + oLeafBeg;
    W.ctaCBeg = D.tabI.p + oCBeg;
    W.nodeA = reinterpret_cast<const int4*>(D.tabI.p + oNodeA); W.nodeB = reinterpret_cast<const int4*>(D.tabI.p + oNodeB);
    W.nodeX = reinterpret_cast<const double4*>(D.tabD.p + oNodeX);
    W.inv = D.inv.p; W.cinv = D.cinv.p; W.vinfo = reinterpret_cast<const float4*>(D.vinfo.p); W.rcC = D.rcCta.p;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// device set-up
__device__ __forceinline__ int find_col(const int32_t* __restrict__ colIdx, int lo, int hi, int col)
{
    for (int b = lo; b < hi; ++b) if (colIdx[b] == col) return b;     // short, unsorted rows
    return -1;
}

// sum of one value over the 8 lanes of a group in LANE ORDER (every lane gets the same, order-exact sum)
__device__ __forceinline__ double group8_sum(double v, unsigned mask, int base)
{
    double acc = __shfl_sync(mask, v, base);
#pragma unroll
    for (int l = 1; l < 8; ++l) acc += __shfl_sync(mask, v, base + l);
    return acc;
}

// A_1 = P_1^T A P_1 on the leaf adjacency, as a GATHER: 8 lanes per leaf row, lane = one fine row of the leaf (leaves hold
// <= 8 rows).  For every leaf block of the row, each lane sums the contributions of ITS fine row's blocks that fall into it
// (stored order), then the 8 partial 6x6 blocks are added in lane order = ascending fine row: no atomics, no memset, a
// fixed order -- the preconditioner (and with it the CG iteration count and the search direction) is reproducible to the
// bit -- and all the loads of a leaf are in flight at once (a serial walk of the 8 rows cost 80 us at 10k faces).
__global__ void __launch_bounds__(256)
mas_galerkin_fine_kernel(int nLeaves, const int32_t* __restrict__ childBeg, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx,
                         const double* __restrict__ val, const float4* __restrict__ vinfo, const int32_t* __restrict__ rowPtr1,
                         const int32_t* __restrict__ colIdx1, double* __restrict__ val1)
{
    const float* vleaf = reinterpret_cast<const float*>(vinfo);
    const int lane8 = threadIdx.x & 7, base = (threadIdx.x & 31) & ~7;
    const unsigned mask = 0xffu << base;
    for (long w = (blockIdx.x * 256L + threadIdx.x) >> 3; w < nLeaves; w += (gridDim.x * 256L) >> 3) {
        const int a = (int)w;
        const int i = childBeg[a] + lane8;
        const bool have = i < childBeg[a + 1];
        float4 vi = make_float4(0.f, 0.f, 0.f, 0.f);
        int b0 = 0, b1 = 0;
        if (have) { vi = vinfo[i]; b0 = rowPtr[i]; b1 = rowPtr[i + 1]; }
        const bool live = have && vi.x != 0.0f;
        const double fi[3] = {mas_unpack(vi.x), mas_unpack(vi.y), mas_unpack(vi.z)};
        for (int s = rowPtr1[a]; s < rowPtr1[a + 1]; ++s) {
            const int a2 = colIdx1[s];
            double acc[36];
#pragma unroll
            for (int q = 0; q < 36; ++q) acc[q] = 0.0;
            if (live)
                for (int b = b0; b < b1; ++b) {
                    const int j = colIdx[b];
                    if (__float_as_int(vleaf[4 * (size_t)j + 3]) != a2) continue;
                    const float4 vj = vinfo[j];
                    if (vj.x == 0.0f) continue;
                    const double fj[3] = {mas_unpack(vj.x), mas_unpack(vj.y), mas_unpack(vj.z)};
                    const double2 A0 = *reinterpret_cast<const double2*>(val + 4 * (size_t)b), A1 = *reinterpret_cast<const double2*>(val + 4 * (size_t)b + 2);
                    const double A[2][2] = {{A0.x, A0.y}, {A1.x, A1.y}};
#pragma unroll
                    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                        for (int qi = 0; qi < 3; ++qi)
#pragma unroll
                            for (int cj = 0; cj < 2; ++cj)
#pragma unroll
                                for (int qj = 0; qj < 3; ++qj)
                                    acc[(ci * 3 + qi) * 6 + cj * 3 + qj] += fi[qi] * A[ci][cj] * fj[qj];
                }
            double* o = val1 + 36 * (size_t)s;
#pragma unroll
            for (int q = 0; q < 36; ++q) {
                const double t = group8_sum(acc[q], mask, base);
                if ((q & 7) == lane8) o[q] = t;                 // the 36 stores spread over the 8 lanes
            }
        }
    }
}

// A_{l+1} = R A_l R^T, gathered the same way: 8 lanes per node row of level l + 1, lane = one child of the node (groups hold
// <= 8 nodes); per block of the row each lane sums R_a S R_b^T over its child's blocks (stored order) whose column's parent
// is the block's column, then the partial blocks are added in lane order = ascending child
__global__ void __launch_bounds__(128)
mas_coarsen_kernel(int nUp, const int32_t* __restrict__ groupBeg, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx,
                   const double* __restrict__ val, const int32_t* __restrict__ parent, const double4* __restrict__ geom,
                   const double4* __restrict__ geomUp, const int32_t* __restrict__ rowPtrUp, const int32_t* __restrict__ colIdxUp,
                   double* __restrict__ valUp)
{
    const int lane8 = threadIdx.x & 7, base = (threadIdx.x & 31) & ~7;
    const unsigned mask = 0xffu << base;
    for (long w = (blockIdx.x * 128L + threadIdx.x) >> 3; w < nUp; w += (gridDim.x * 128L) >> 3) {
        const int pa = (int)w;
        const int a = groupBeg[pa] + lane8;
        const bool live = a < groupBeg[pa + 1];
        const double4 gpa = geomUp[pa];
        const double isa = 1.0 / gpa.z;
        double Ra10 = 0.0, Ra11 = 0.0, Ra20 = 0.0, Ra22 = 0.0;
        int k0 = 0, k1 = 0;
        if (live) {
            const double4 ga = geom[a];
            Ra10 = (ga.x - gpa.x) * isa; Ra11 = ga.z * isa; Ra20 = (ga.y - gpa.y) * isa; Ra22 = ga.z * isa;
            k0 = rowPtr[a]; k1 = rowPtr[a + 1];
        }
        for (int s = rowPtrUp[pa]; s < rowPtrUp[pa + 1]; ++s) {
            const int pb = colIdxUp[s];
            const double4 gpb = geomUp[pb];
            const double isb = 1.0 / gpb.z;
            double acc[36];
#pragma unroll
            for (int q = 0; q < 36; ++q) acc[q] = 0.0;
            for (int blk = k0; blk < k1; ++blk) {
                const int b = colIdx[blk];
                if (parent[b] != pb) continue;
                const double4 gb = geom[b];
                const double Rb10 = (gb.x - gpb.x) * isb, Rb11 = gb.z * isb, Rb20 = (gb.y - gpb.y) * isb, Rb22 = gb.z * isb;
                const double* M = val + 36 * (size_t)blk;
#pragma unroll
                for (int ci = 0; ci < 2; ++ci)
#pragma unroll
                    for (int cj = 0; cj < 2; ++cj) {
                        double T[3][3];
                        // T = S Rb^T
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double S0 = M[(ci * 3 + i) * 6 + cj * 3], S1 = M[(ci * 3 + i) * 6 + cj * 3 + 1], S2 = M[(ci * 3 + i) * 6 + cj * 3 + 2];
                            T[i][0] = S0;
                            T[i][1] = S0 * Rb10 + S1 * Rb11;
                            T[i][2] = S0 * Rb20 + S2 * Rb22;
                        }
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            acc[(ci * 3 + 0) * 6 + cj * 3 + j] += T[0][j];
                            acc[(ci * 3 + 1) * 6 + cj * 3 + j] += Ra10 * T[0][j] + Ra11 * T[1][j];
                            acc[(ci * 3 + 2) * 6 + cj * 3 + j] += Ra20 * T[0][j] + Ra22 * T[2][j];
                        }
                    }
            }
            double* o = valUp + 36 * (size_t)s;
#pragma unroll
            for (int q = 0; q < 36; ++q) {
                const double t = group8_sum(acc[q], mask, base);
                if ((q & 7) == lane8) o[q] = t;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// coarse level: dense Galerkin matrix and its exact inverse
// Diagonal equilibration of the coarse matrix (option mas_equilibrate): scale[k] = 1 / sqrt(A_kk) (1 where A_kk <= 0: a node of fixed
// vertices only).  The Gauss-Jordan elimination below has no pivoting; on a badly scaled Galerkin matrix (a nearly degenerate
// triangle puts 1e10 next to 1e-3 on the diagonal) it loses the soft part to rounding and the inverse comes out indefinite; the
// equilibrated matrix has a unit diagonal and the elimination error is relative to each DOF's own stiffness.
__global__ void __launch_bounds__(256)
mas_dense_scale_kernel(int nNodes, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, const double* __restrict__ val, int ld, double* __restrict__ scale)
{
    for (int k = blockIdx.x * 256 + threadIdx.x; k < ld; k += gridDim.x * 256) {
        double d = 1.0;
        if (k < nNodes * kMasDof) {
            const int a = k / kMasDof, i = k % kMasDof;
            d = 0.0;
            for (int blk = rowPtr[a]; blk < rowPtr[a + 1]; ++blk) if (colIdx[blk] == a) d = val[36 * (size_t)blk + 7 * i];
        }
        scale[k] = d > 0.0 ? 1.0 / sqrt(d) : 1.0;
    }
}
__global__ void __launch_bounds__(256)
mas_dense_fill_kernel(int nNodes, const int32_t* __restrict__ rowPtr, const int32_t* __restrict__ colIdx, const double* __restrict__ val,
                      int ld, int nC, double* __restrict__ X, const double* __restrict__ scale)
{
    const int total = nNodes * 36;
    for (int w = blockIdx.x * 256 + threadIdx.x; w < total; w += gridDim.x * 256) {
        const int a = w / 36, ij = w % 36, i = ij / 6, j = ij % 6;
        for (int blk = rowPtr[a]; blk < rowPtr[a + 1]; ++blk) {
            const int r = kMasDof * a + i, cc = kMasDof * colIdx[blk] + j;
            X[(size_t)r * ld + cc] = scale ? val[36 * (size_t)blk + ij] * (scale[r] * scale[cc]) : val[36 * (size_t)blk + ij];
        }
    }
    for (int k = nC + blockIdx.x * 256 + threadIdx.x; k < ld; k += gridDim.x * 256) X[(size_t)k * ld + k] = 1.0;     // padding
}

// In-place inverse of the SPD coarse matrix by blocked Gauss-Jordan (48x48 tiles, no pivoting), one cooperative
// launch over the whole GPU: step k turns buffer X into buffer Y tile by tile,
//     Y_kk = P,  Y_kj = P X_kj,  Y_ik = -X_ik P,  Y_ij = X_ij - X_ik P X_kj      (P = X_kk^-1)
// and the CTA that produces Y_(k+1)(k+1) inverts it right away (look-ahead) so that one grid barrier per step is
// all the synchronisation there is.  A DOF whose pivot collapses (a coarse node of fixed vertices only, collinear
// leaf) is dropped: zero row and column in the inverse.  Finally the symmetrised result is stored in fp32.
// A DOF is dropped (zero row and column of the inverse) when its pivot fell below this fraction of its original diagonal:
// it is (nearly) in the span of the DOFs eliminated before it -- collinear vertices make a leaf's affine functions
// dependent -- and its inverse entries would exceed what the fp32 storage resolves.  With 1e-10 such pivots were rounding
// noise on a knife edge: ~1 solve in 20 got an indefinite preconditioner (tools/gpu_freerun_check.py).
static constexpr double kMasPivotTol = 1e-6;
// tile and its padded shared-memory stride.  NEXT: a stride of 52 (= 4 mod 16 doubles) makes the 64-bit fragment loads of
// tile_gemm conflict-free (a half-warp = 4 rows x 4 consecutive doubles then touches 16 distinct 8-byte bank pairs: 2
// wavefronts per load instead of 4 with 50, by the bank arithmetic); not yet measured on the GPU, so 50 stays.
static constexpr int kCB = kMasCoarseBlk, kCBs = kMasCoarseBlk + 2;
static constexpr int kDenseThreads = 576;                                // 18 warps: warp w owns the 8x8 sub-tiles 2w and 2w+1 (6x6 grid)
struct DenseInvArgs { int nb, ld, nC; double* X; double* Y; double* P; double* diag0; float* out; long long* dbg; int* ticket; const double* scale; };

// The tile products run on the FP64 tensor cores (mma.sync m8n8k4: the only tensor path for doubles; nothing else on
// this path is a dense contraction).  A thread owns 4 elements of a 48x48 tile, in the accumulator-fragment layout:
// row = 8 * (warp / 3) + lane / 4, columns c0 + {0, 1, 8, 9} with c0 = 8 * ((2 * warp) % 6) + 2 * (lane % 4).
struct TileMap { int r, c0; };
__device__ __forceinline__ TileMap tile_map()
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TileMap m; m.r = 8 * (warp / 3) + (lane >> 2); m.c0 = 8 * ((2 * warp) % 6) + 2 * (lane & 3);
    return m;
}
__device__ __forceinline__ int tile_col(const TileMap& m, int e) { return m.c0 + (e & 1) + 8 * (e >> 1); }

__device__ __forceinline__ void tile_load_regs(double (&c)[4], const double* __restrict__ G, int ld, const TileMap& m)
{
    // through L2: the tile was written by another CTA in the previous step
    const double2 a = __ldcg(reinterpret_cast<const double2*>(G + (size_t)m.r * ld + m.c0)), b = __ldcg(reinterpret_cast<const double2*>(G + (size_t)m.r * ld + m.c0 + 8));
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
}
__device__ __forceinline__ void tile_store_smem(double (*S)[kCBs], const double (&c)[4], const TileMap& m)
{
    *reinterpret_cast<double2*>(&S[m.r][m.c0]) = make_double2(c[0], c[1]);
    *reinterpret_cast<double2*>(&S[m.r][m.c0 + 8]) = make_double2(c[2], c[3]);
}
__device__ __forceinline__ void tile_load(double (*S)[kCBs], const double* __restrict__ G, int ld, const TileMap& m)
{
    double c[4];
    tile_load_regs(c, G, ld, m);
    tile_store_smem(S, c, m);
}
// c += sign * A * B  (48x48 tiles in shared memory)
__device__ __forceinline__ void tile_gemm(const double (*A)[kCBs], const double (*B)[kCBs], double (&c)[4], double sign)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowA = 8 * (warp / 3) + (lane >> 2), kq = lane & 3;
    const int colB0 = 8 * ((2 * warp) % 6) + (lane >> 2);
    double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
    for (int kk = 0; kk < kCB / 4; ++kk) {
        const double a = A[rowA][4 * kk + kq];
        const double b0 = B[4 * kk + kq][colB0], b1 = B[4 * kk + kq][colB0 + 8];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b0));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d2), "+d"(d3) : "d"(a), "d"(b1));
    }
    c[0] += sign * d0; c[1] += sign * d1; c[2] += sign * d2; c[3] += sign * d3;
}
// Gauss-Jordan inverse of one SPD tile (in shared memory T on entry and exit).  The sequential chain of 48 pivots is
// what bounds the whole blocked inversion, so it runs on 5 warps only (a named barrier over 160 threads is cheaper than
// one over 18 warps, and no warp waits for an issue slot): thread t < 144 keeps row t % 48, columns 16 (t / 48) .. + 15 in
// REGISTERS; per pivot the pivot row and column cross through a double-buffered shared-memory line, one barrier per
// pivot.  buf: 2 x (48 row + 48 column) doubles.  A DOF whose pivot collapses is dropped (zero row and column).
// The mapping is SEGMENT-major on purpose: the lanes of a warp then read the same 16 bytes of the pivot row (a pure
// broadcast).  Row-major (row t / 3, segment t % 3) put the three segments of a warp 128 bytes apart, i.e. on the same
// banks: 3-way conflicts on every one of the 8 loads per pivot, 602 cycles per pivot against 285 now
// (tools/micro/tile_invert_bench.cu, which also shows that 2x2 block pivots and 9 / 18 warps are slower).
__device__ __forceinline__ void tile_invert_smem(double (*T)[kCBs], double* buf, const double* d0s)
{
    __syncthreads();
    const int t = threadIdx.x;
    if (t < 160) {
        const bool act = t < 144;
        const int r = act ? t % kCB : 0, c0 = act ? 16 * (t / kCB) : 0;
        double d[16];
#pragma unroll
        for (int e = 0; e < 16; e += 2) { const double2 v = *reinterpret_cast<const double2*>(&T[r][c0 + e]); d[e] = v.x; d[e + 1] = v.y; }
        // The pivot loop is unrolled over the 16 columns of a segment so that every register index is a compile-time
        // constant (a run-time index would push d[] into local memory; rotating the registers instead was measured 2x
        // slower).  What bounds a pivot is the chain of DEPENDENT fp64 operations between two barriers (~700 cycles with
        // the IEEE reciprocal), so the reciprocal is the hardware approximation plus ONE Newton step (relative error
        // ~1e-12: this is a preconditioner, stored in fp32 anyway).
        for (int kq = 0; kq < kCB / 16; ++kq) {
            const bool mine = act && c0 == 16 * kq;           // this thread's segment holds the pivot columns of this round
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                const int k = 16 * kq + kk;
                double* rowb = buf + (k & 1) * 2 * kCB;
                double* colb = rowb + kCB;
                if (act && r == k) {
#pragma unroll
                    for (int e = 0; e < 16; e += 2) *reinterpret_cast<double2*>(rowb + c0 + e) = make_double2(d[e], d[e + 1]);
                }
                if (mine) colb[r] = d[kk];
                asm volatile("bar.sync 1, 160;" ::: "memory");
                // Everything below is straight-line code ordered for the in-order issue: the reciprocal goes first, the
                // collapsed-pivot test runs in its shadow and is applied with selects, and the two candidates for the row
                // coefficient are formed in parallel.  The reciprocal is the hardware approximation plus ONE Newton step
                // (relative error ~1e-12).  The bare approximation (1e-6) is NOT enough: the coarse Galerkin matrices are
                // ill-conditioned, the inverse came out indefinite on the third bimba iteration and CG stalled
                // (tools/gpu_freerun_check.py).
                const double p = rowb[k], dk0 = d0s[k], f = colb[r];
                double ip;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(ip) : "d"(p));
                ip = fma(ip, fma(-p, ip, 1.0), ip);
                const bool bad = !(dk0 > 0.0) || !(p > kMasPivotTol * dk0);
                // one FMA per element for every row: the pivot row's own values ARE the row buffer, so its scaling d * g is
                // d + (g - 1) * rowb; a collapsed pivot zeroes its row (coefficient -1) and leaves the other rows alone
                const double gOther = -f * ip, gPivot = ip - 1.0;
                const double coef = r != k ? (bad ? 0.0 : gOther) : (bad ? -1.0 : gPivot);
                const double g = bad ? 0.0 : (r != k ? gOther : ip);       // new value of the pivot-column element
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const double2 rj = *reinterpret_cast<const double2*>(rowb + c0 + e);
                    d[e] += coef * rj.x; d[e + 1] += coef * rj.y;
                }
                if (mine) d[kk] = g;
            }
        }
        if (act) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) *reinterpret_cast<double2*>(&T[r][c0 + e]) = make_double2(d[e], d[e + 1]);
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void tile_load_from_smem(double (&c)[4], const double (*S)[kCBs], const TileMap& m)
{
    const double2 a = *reinterpret_cast<const double2*>(&S[m.r][m.c0]), b = *reinterpret_cast<const double2*>(&S[m.r][m.c0 + 8]);
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
}

__global__ void __launch_bounds__(kDenseThreads, 1)
mas_dense_invert_kernel(DenseInvArgs A)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smraw[];
    double (*sP)[kCBs] = reinterpret_cast<double (*)[kCBs]>(smraw);
    double (*sA)[kCBs] = sP + kCB;
    double (*sB)[kCBs] = sA + kCB;
    double (*sR)[kCBs] = sB + kCB;
    double* d0s = reinterpret_cast<double*>(sR + kCB);
    double* ibuf = d0s + kCB;                       // 4 x 48
    const int tid = threadIdx.x;
    const TileMap m = tile_map();
    const int nb = A.nb, ld = A.ld;
    for (int k = blockIdx.x * kDenseThreads + tid; k < ld; k += gridDim.x * kDenseThreads) A.diag0[k] = A.X[(size_t)k * ld + k];
    grid.sync();
    if (blockIdx.x == 0) {
        if (tid < kCB) d0s[tid] = A.diag0[tid];
        double d[4];
        tile_load(sA, A.X, ld, m);
        tile_invert_smem(sA, ibuf, d0s);
        tile_load_from_smem(d, sA, m);
#pragma unroll
        for (int e = 0; e < 4; ++e) A.P[(size_t)m.r * kCB + tile_col(m, e)] = d[e];
    }
    double* X = A.X; double* Y = A.Y;
    long long tSync = 0, tTile = 0, tInv = 0;
    for (int k = 0; k < nb; ++k) {
        long long t0 = A.dbg ? clock64() : 0;
        grid.sync();
        if (A.dbg) { const long long t1 = clock64(); tSync += t1 - t0; t0 = t1; }
        tile_load(sP, A.P + (size_t)(k & 1) * kCB * kCB, kCB, m);
        __syncthreads();
        const int special = k + 1 < nb ? (k + 1) * nb + (k + 1) : -1;
        const bool owner = special >= 0 && (int)blockIdx.x == special % (int)gridDim.x;
        // tiles of this CTA, the look-ahead tile first; software-pipelined: the operands of the NEXT tile are already
        // on their way from L2 into registers while the two tensor-core products of the current one run
        // tiles are handed out by a ticket counter per step (the owner of the look-ahead tile spends a long time on
        // its inversion and must not sit on a fixed share of the trailing tiles)
        int q = owner ? -1 : 0;
        __shared__ int sTicket;
        auto advance = [&](int& qq) -> int {
            if (qq < 0) { qq = 0; return special; }
            for (;;) {
                __syncthreads();
                if (tid == 0) sTicket = atomicAdd(A.ticket + k, 1);
                __syncthreads();
                const int t = sTicket;
                if (t >= nb * nb) return -1;
                if (t != special) return t;
            }
        };
        auto issue = [&](int t, double (&pb)[4], double (&pa)[4], double (&pc)[4]) {
            const int i = t / nb, j = t % nb;
            if (j != k) tile_load_regs(pb, X + (size_t)k * kCB * ld + (size_t)j * kCB, ld, m);
            if (i != k) tile_load_regs(pa, X + (size_t)i * kCB * ld + (size_t)k * kCB, ld, m);
            if (i != k && j != k) tile_load_regs(pc, X + (size_t)i * kCB * ld + (size_t)j * kCB, ld, m);
        };
        int tcur = advance(q);
        double cb[4] = {0.0, 0.0, 0.0, 0.0}, ca[4] = {0.0, 0.0, 0.0, 0.0}, cc[4] = {0.0, 0.0, 0.0, 0.0};
        if (tcur >= 0) issue(tcur, cb, ca, cc);
        bool first = owner;
        while (tcur >= 0) {
            const int i = tcur / nb, j = tcur % nb;
            if (j != k) tile_store_smem(sB, cb, m);
            if (i != k) tile_store_smem(sA, ca, m);
            double c[4] = {cc[0], cc[1], cc[2], cc[3]};
            const int tnext = advance(q);
            if (tnext >= 0) issue(tnext, cb, ca, cc);
            __syncthreads();
            if (i == k && j == k) {
                c[0] = sP[m.r][m.c0]; c[1] = sP[m.r][m.c0 + 1]; c[2] = sP[m.r][m.c0 + 8]; c[3] = sP[m.r][m.c0 + 9];
            } else if (i == k) {
                c[0] = c[1] = c[2] = c[3] = 0.0;
                tile_gemm(sP, sB, c, 1.0);
            } else if (j == k) {
                c[0] = c[1] = c[2] = c[3] = 0.0;
                tile_gemm(sA, sP, c, -1.0);
            } else {
                double r[4] = {0.0, 0.0, 0.0, 0.0};
                tile_gemm(sP, sB, r, 1.0);
                tile_store_smem(sR, r, m);
                __syncthreads();
                tile_gemm(sA, sR, c, -1.0);
            }
            double* Yij = Y + (size_t)i * kCB * ld + (size_t)j * kCB + (size_t)m.r * ld + m.c0;
            *reinterpret_cast<double2*>(Yij) = make_double2(c[0], c[1]);
            *reinterpret_cast<double2*>(Yij + 8) = make_double2(c[2], c[3]);
            __syncthreads();                   // the shared tiles are reused by the next tile
            if (first) {                       // look-ahead: invert the next pivot tile now
                first = false;
                const long long ti0 = A.dbg ? clock64() : 0;
                if (tid < kCB) d0s[tid] = A.diag0[(k + 1) * kCB + tid];
                tile_store_smem(sR, c, m);
                tile_invert_smem(sR, ibuf, d0s);
                tile_load_from_smem(c, sR, m);
                double* Pn = A.P + (size_t)((k + 1) & 1) * kCB * kCB;
#pragma unroll
                for (int e = 0; e < 4; ++e) Pn[(size_t)m.r * kCB + tile_col(m, e)] = c[e];
                if (A.dbg) tInv += clock64() - ti0;
            }
            tcur = tnext;
        }
        if (A.dbg) tTile += clock64() - t0;
        double* T = X; X = Y; Y = T;
    }
    if (A.dbg && tid == 0) { A.dbg[3 * blockIdx.x] = tSync; A.dbg[3 * blockIdx.x + 1] = tTile; A.dbg[3 * blockIdx.x + 2] = tInv; }
    grid.sync();
    const size_t total = (size_t)ld * ld;
    for (size_t e = (size_t)blockIdx.x * kDenseThreads + tid; e < total; e += (size_t)gridDim.x * kDenseThreads) {
        const int i = (int)(e / ld), j = (int)(e % ld);
        const double sij = A.scale ? A.scale[i] * A.scale[j] : 1.0;         // equilibrated inversion: the inverse of S A S is S^-1 A^-1 S^-1
        A.out[e] = (i < A.nC && j < A.nC) ? mas_pack(0.5 * (__ldcg(X + e) + __ldcg(X + (size_t)j * ld + i)) * sij) : 0.0f;
    }
}

// group blocks D_l[g] (<= 48x48) gathered from A_l straight into registers and inverted with the register-resident
// Gauss-Jordan above (a DOF whose pivot collapses -- a node of fixed vertices only, collinear vertices -- is
// dropped); all levels below the coarse one in one launch: CTA -> (level, group) through the prefix table.
struct MasInvertArgs {
    int L;
    int groupPre[kMasMaxLevels + 1];
    const int32_t* rowPtr[kMasMaxLevels]; const int32_t* colIdx[kMasMaxLevels]; const double* val[kMasMaxLevels];
    const int32_t* parent[kMasMaxLevels]; const int32_t* groupBeg[kMasMaxLevels]; float* inv[kMasMaxLevels];
    int equilibrate;
};
__global__ void __launch_bounds__(kDenseThreads)
mas_invert_kernel(MasInvertArgs P)
{
    __shared__ double T[kCB][kCBs];
    __shared__ double d0s[kCB];
    __shared__ double ibuf[4 * kCB];
    int l = 0;
    while (l + 1 < P.L && (int)blockIdx.x >= P.groupPre[l + 1]) ++l;
    const int g = blockIdx.x - P.groupPre[l];
    const int gb = P.groupBeg[l][g], nch = P.groupBeg[l][g + 1] - gb;
    const int nd = nch * kMasDof;
    const TileMap m = tile_map();
    double d[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {                  // the thread's two column pairs lie in one 6x6 block each
        const int c = m.c0 + 8 * h;
        double v0 = 0.0, v1 = 0.0;
        if (m.r < nd && c < nd) {
            const int a = gb + m.r / kMasDof, b = gb + c / kMasDof;
            for (int blk = P.rowPtr[l][a]; blk < P.rowPtr[l][a + 1]; ++blk)
                if (P.colIdx[l][blk] == b) {
                    const double* src = P.val[l] + 36 * (size_t)blk + (m.r % kMasDof) * 6 + c % kMasDof;
                    v0 = src[0]; v1 = src[1];
                }
        } else {
            v0 = m.r == c ? 1.0 : 0.0; v1 = m.r == c + 1 ? 1.0 : 0.0;      // identity padding
        }
        d[2 * h] = v0; d[2 * h + 1] = v1;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) if (tile_col(m, e) == m.r) d0s[m.r] = d[e];
    __shared__ double sc[kCB];
    if (P.equilibrate) {                           // see mas_dense_scale_kernel: invert S D S with S = diag^-1/2, store S (S D S)^-1 S
        __syncthreads();
        if (threadIdx.x < kCB) { const double dd = d0s[threadIdx.x]; sc[threadIdx.x] = dd > 0.0 ? 1.0 / sqrt(dd) : 1.0; }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] *= sc[m.r] * sc[tile_col(m, e)];
        if (threadIdx.x < kCB && d0s[threadIdx.x] > 0.0) d0s[threadIdx.x] = 1.0;
    }
    tile_store_smem(T, d, m);
    tile_invert_smem(T, ibuf, d0s);
    float* out = P.inv[l] + (size_t)g * kMasBlk * kMasBlk;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = tile_col(m, e);
        const double sij = P.equilibrate ? sc[m.r] * sc[j] : 1.0;
        out[m.r * kMasBlk + j] = (m.r < nd && j < nd) ? mas_pack(0.5 * (T[m.r][j] + T[j][m.r]) * sij) : 0.0f;
    }
}

int launch_mas_setup(ocb_ctx* c)
{
    MasHost& H = c->masH;
    MasDev& D = c->masD;
    if (!H.enabled) return 0;
    ProfScope prof(c, K_MAS_SETUP);
    {   // every block of every level is written by exactly one thread (gather): no memset
        const MasLevel& V1 = D.lv[0];
        int grid = (int)((8L * V1.nNodes + 255) / 256); if (grid > c->numSMs * 16) grid = c->numSMs * 16; if (grid < 1) grid = 1;
        mas_galerkin_fine_kernel<<<grid, 256, 0, c->stream>>>(V1.nNodes, V1.childBeg, c->rowPtr.p, c->colIdx.p, c->val.p, reinterpret_cast<const float4*>(D.vinfo.p),
                                                             D.lvRowPtr[0], D.lvColIdx[0], D.lvVal[0]);
        KCHECK(c);
    }
    for (int l = 1; l < H.L; ++l) {
        const MasLevel& V = D.lv[l - 1];
        const int nUp = D.lv[l].nNodes;
        int g = (int)((8L * nUp + 127) / 128); if (g > c->numSMs * 16) g = c->numSMs * 16; if (g < 1) g = 1;
        mas_coarsen_kernel<<<g, 128, 0, c->stream>>>(nUp, V.groupBeg, D.lvRowPtr[l - 1], D.lvColIdx[l - 1], D.lvVal[l - 1], V.parent, V.geom,
                                                    D.lv[l].geom, D.lvRowPtr[l], D.lvColIdx[l], D.lvVal[l]);
        KCHECK(c);
    }
    if (H.L > 1) {                             // group inverses of the levels below the coarse one
        MasInvertArgs A;
        A.L = H.L - 1;
        A.groupPre[0] = 0;
        for (int l = 1; l < H.L; ++l) {
            const MasLevel& V = D.lv[l - 1];
            A.groupPre[l] = A.groupPre[l - 1] + V.nGroups;
            A.rowPtr[l - 1] = D.lvRowPtr[l - 1]; A.colIdx[l - 1] = D.lvColIdx[l - 1]; A.val[l - 1] = D.lvVal[l - 1];
            A.parent[l - 1] = V.parent; A.groupBeg[l - 1] = V.groupBeg; A.inv[l - 1] = V.inv;
        }
        A.equilibrate = c->masEquilibrate ? 1 : 0;
        mas_invert_kernel<<<A.groupPre[H.L - 1], kDenseThreads, 0, c->stream>>>(A);
        KCHECK(c);
    }
    {                                          // the coarse level: dense fill + exact inverse
        const MasView& W = D.view;
        const MasLevel& V = D.lv[H.L - 1];
        const size_t ld = (size_t)W.ldC;
        DenseInvArgs A;
        A.nb = W.ldC / kCB; A.ld = W.ldC; A.nC = W.nC;
        A.X = D.dense.p; A.Y = A.X + ld * ld; A.diag0 = A.Y + ld * ld; A.P = A.diag0 + ld; A.out = D.cinv.p;
        A.ticket = reinterpret_cast<int*>(A.P + 2 * kCB * kCB);
        double* scale = A.P + 2 * kCB * kCB + 8 + kMasCoarseMax / kMasCoarseBlk;
        A.scale = c->masEquilibrate ? scale : nullptr;
        OCB_CUDA(c, cudaMemsetAsync(A.ticket, 0, sizeof(int) * (size_t)(A.nb + 1), c->stream));
        OCB_CUDA(c, cudaMemsetAsync(A.X, 0, ld * ld * sizeof(double), c->stream));
        int g = (V.nNodes * 36 + 255) / 256; if (g > c->numSMs * 8) g = c->numSMs * 8; if (g < 1) g = 1;
        if (A.scale) { mas_dense_scale_kernel<<<(W.ldC + 255) / 256, 256, 0, c->stream>>>(V.nNodes, D.lvRowPtr[H.L - 1], D.lvColIdx[H.L - 1], D.lvVal[H.L - 1], W.ldC, scale); KCHECK(c); }
        mas_dense_fill_kernel<<<g, 256, 0, c->stream>>>(V.nNodes, D.lvRowPtr[H.L - 1], D.lvColIdx[H.L - 1], D.lvVal[H.L - 1], W.ldC, W.nC, A.X, A.scale);
        KCHECK(c);
        const size_t smem = 4 * (size_t)kCB * kCBs * sizeof(double) + 5 * kCB * sizeof(double);
        if (!D.denseAttr) {
            OCB_CUDA(c, cudaFuncSetAttribute(mas_dense_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            D.denseAttr = true;
        }
        int gi = A.nb * A.nb; if (gi > c->numSMs) gi = c->numSMs;
        static const bool dbgOn = []() { const char* e = getenv("OCB_MAS_DEBUG"); return e && atoi(e); }();
        A.dbg = nullptr;
        if (dbgOn) { cudaMalloc((void**)&A.dbg, 3 * sizeof(long long) * gi); cudaMemset(A.dbg, 0, 3 * sizeof(long long) * gi); }
        void* args[] = {&A};
        OCB_CUDA(c, cudaLaunchCooperativeKernel((void*)mas_dense_invert_kernel, dim3(gi), dim3(kDenseThreads), args, smem, c->stream));
        c->launches++;
        if (dbgOn) {
            std::vector<long long> h(3 * (size_t)gi);
            cudaStreamSynchronize(c->stream);
            cudaMemcpy(h.data(), A.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            cudaFree(A.dbg);
            long long mx[3] = {0, 0, 0}; double av[3] = {0, 0, 0};
            for (int b = 0; b < gi; ++b) for (int q = 0; q < 3; ++q) { mx[q] = std::max(mx[q], h[3 * b + q]); av[q] += (double)h[3 * b + q] / gi; }
            fprintf(stderr, "[ocb mas] dense inverse n %d (%d tiles^2) grid %d: cycles/step  grid.sync avg %.0f max %.0f | tiles avg %.0f max %.0f | invert (sum over CTAs) %.0f\n",
                    A.nC, A.nb, gi, av[0] / A.nb, (double)mx[0] / A.nb, av[1] / A.nb, (double)mx[1] / A.nb, av[2] * gi / A.nb);
        }
    }
    return 0;
}

}  // namespace ocb
