// Scaffold refresh fast path.  With bijectivity on, the reference rebuilds the global Scaffold after EVERY
// Newton iteration (Optimizer.cpp:236-239, 253-256): igl::boundary_loop and igl::components recompute the mesh's whole
// triangle-triangle / vertex-triangle adjacency each time (23 ms of the 27 ms per rebuild at 10k faces, profiles/
// r2_scaffold_phases.txt) although the mesh CONNECTIVITY only changes in a topology step, and Scaffold::mergeVNeighbor
// copies 5 000 std::sets for a LinSysSolver::set_pattern call that the device-resident solver ignores.
// shim/Makefile compiles a build-time copy of the reference's Scaffold.cpp in which the two libigl calls of the global
// branch (Scaffold.cpp:39, 94) go through the wrappers below: memoised on the CONTENT of F (plain Newton iterations do not
// change it) and, on a miss, recomputed by O(|F|) re-derivations that return exactly what libigl returns (same loops, same
// start vertices, same labels: OCB_SCAFFOLD_VERIFY=1 cross-checks every call; the bit-identical traces of
// tests/test_host_logic.py cover whole runs) -- so the air mesh's input, hence Triangle's output and the DOF numbering, are
// unchanged bit for bit -- and a copy of Optimizer.cpp in which the mergeVNeighbor calls are skipped while the Optimizer
// hooks keep the solve on the device (OCB_MERGE_VNEIGHBOR).
#ifndef CudaScaffoldFast_hpp
#define CudaScaffoldFast_hpp

#include <Eigen/Core>
#include <igl/boundary_loop.h>
#include <igl/components.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace OptCuts {

struct OcbConnectivityCache {
    std::mutex mu;
    Eigen::MatrixXi F_loops, F_comp;
    std::vector<std::vector<int>> loops;
    Eigen::VectorXi comp;
    long hits = 0, misses = 0;
    static OcbConnectivityCache& get(void) { static OcbConnectivityCache* c = new OcbConnectivityCache(); return *c; }
    static bool same(const Eigen::MatrixXi& a, const Eigen::MatrixXi& b) {
        return a.rows() == b.rows() && a.cols() == b.cols() && a.size() > 0 && std::memcmp(a.data(), b.data(), sizeof(int) * a.size()) == 0;
    }
};

// igl::boundary_loop(F, L) re-derived in O(|F|) for an oriented manifold triangle mesh (what a UV mesh is), SAME output:
// loops start at the smallest unvisited border vertex; from vertex v the walk takes, among v's incident faces in ascending
// face order, the first face whose edge leaving v (v -> next corner) has no opposite half-edge and whose target is unvisited
// (igl/boundary_loop.cpp:36-83 with triangle_triangle_adjacency / vertex_triangle_adjacency / is_border_vertex).
// libigl rebuilds two adjacency structures through sorted std::vectors of std::vectors for that; here: one open-addressing
// table of the half-edges and one counting sort of the corners.  OCB_SCAFFOLD_VERIFY=1 cross-checks against libigl.
inline void ocbFastBoundaryLoop(const Eigen::MatrixXi& F, std::vector<std::vector<int>>& L)
{
    L.clear();
    const int nF = static_cast<int>(F.rows());
    if (nF == 0) return;
    const int nV = F.maxCoeff() + 1;
    // half-edge table: key a * nV + b
    size_t cap = 1; while (cap < static_cast<size_t>(6) * nF) cap <<= 1;
    std::vector<long long> keys(cap, -1);
    auto slotOf = [&](long long key) { size_t h = (static_cast<unsigned long long>(key) * 0x9E3779B97F4A7C15ull) >> 20; return h & (cap - 1); };
    for (int f = 0; f < nF; ++f) for (int e = 0; e < 3; ++e) {
        const long long key = static_cast<long long>(F(f, e)) * nV + F(f, (e + 1) % 3);
        size_t h = slotOf(key);
        while (keys[h] != -1 && keys[h] != key) h = (h + 1) & (cap - 1);
        keys[h] = key;
    }
    auto has = [&](int a, int b) { const long long key = static_cast<long long>(a) * nV + b; size_t h = slotOf(key);
                                   while (keys[h] != -1) { if (keys[h] == key) return true; h = (h + 1) & (cap - 1); } return false; };
    // boundary half-edges (TT(f,e) < 0) and border vertices
    std::vector<unsigned char> bndEdge(static_cast<size_t>(3) * nF, 0), faceHasBnd(nF, 0), unvisited(nV, 0);
    for (int f = 0; f < nF; ++f) for (int e = 0; e < 3; ++e) {
        const int a = F(f, e), b = F(f, (e + 1) % 3);
        if (!has(b, a)) { bndEdge[3 * static_cast<size_t>(f) + e] = 1; faceHasBnd[f] = 1; unvisited[a] = 1; unvisited[b] = 1; }
    }
    // vertex -> incident faces in ascending face order (vertex_triangle_adjacency), border vertices only
    std::vector<int> ptr(nV + 1, 0);
    for (int f = 0; f < nF; ++f) for (int e = 0; e < 3; ++e) if (unvisited[F(f, e)]) ptr[F(f, e) + 1]++;
    for (int v = 0; v < nV; ++v) ptr[v + 1] += ptr[v];
    std::vector<int> fill(ptr.begin(), ptr.end() - 1), vf(ptr[nV]);
    for (int f = 0; f < nF; ++f) for (int e = 0; e < 3; ++e) if (unvisited[F(f, e)]) vf[fill[F(f, e)]++] = 3 * f + e;
    for (int start = 0; start < nV; ++start) {
        if (!unvisited[start]) continue;
        std::vector<int> l;
        unvisited[start] = 0;
        l.push_back(start);
        for (;;) {
            const int v = l.back();
            int next = -1;
            for (int q = ptr[v]; q < ptr[v + 1] && next < 0; ++q) {
                const int f = vf[q] / 3, vLoc = vf[q] % 3;
                if (!faceHasBnd[f]) continue;
                const int vNext = F(f, (vLoc + 1) % 3);
                if (unvisited[vNext] && bndEdge[3 * static_cast<size_t>(f) + vLoc]) next = vNext;
            }
            if (next < 0) break;
            l.push_back(next);
            unvisited[next] = 0;
        }
        L.push_back(l);
    }
}

// igl::components(F, C) re-derived with a union-find: label k = the k-th component met when the vertices are scanned in
// ascending order (igl/components.cpp: breadth-first searches started from the vertices in order)
inline void ocbFastComponents(const Eigen::MatrixXi& F, Eigen::VectorXi& C)
{
    const int nF = static_cast<int>(F.rows());
    const int nV = nF ? F.maxCoeff() + 1 : 0;
    std::vector<int> parent(nV);
    for (int v = 0; v < nV; ++v) parent[v] = v;
    auto find = [&](int v) { while (parent[v] != v) { parent[v] = parent[parent[v]]; v = parent[v]; } return v; };
    for (int f = 0; f < nF; ++f) for (int e = 0; e < 3; ++e) {
        const int a = find(F(f, e)), b = find(F(f, (e + 1) % 3));
        if (a != b) parent[a < b ? b : a] = a < b ? a : b;                 // the root is the component's smallest vertex
    }
    C.resize(nV);
    std::vector<int> label(nV, -1);
    int next = 0;
    for (int v = 0; v < nV; ++v) { const int r = find(v); if (label[r] < 0) label[r] = next++; C[v] = label[r]; }
}

inline bool ocbScaffoldVerify(void) { static const bool on = []() { const char* e = std::getenv("OCB_SCAFFOLD_VERIFY"); return e && std::atoi(e) != 0; }(); return on; }

inline void ocbCachedBoundaryLoop(const Eigen::MatrixXi& F, std::vector<std::vector<int>>& L)
{
    OcbConnectivityCache& C = OcbConnectivityCache::get();
    std::lock_guard<std::mutex> lock(C.mu);
    if (!OcbConnectivityCache::same(C.F_loops, F)) {
        ocbFastBoundaryLoop(F, C.loops);
        if (ocbScaffoldVerify()) {
            std::vector<std::vector<int>> ref;
            igl::boundary_loop(F, ref);
            if (ref != C.loops) { std::fprintf(stderr, "optcuts_b200: boundary loops differ from igl::boundary_loop\n"); std::abort(); }
        }
        C.F_loops = F;
        C.misses++;
    } else C.hits++;
    L = C.loops;
}

inline void ocbCachedComponents(const Eigen::MatrixXi& F, Eigen::VectorXi& compI_V)
{
    OcbConnectivityCache& C = OcbConnectivityCache::get();
    std::lock_guard<std::mutex> lock(C.mu);
    if (!OcbConnectivityCache::same(C.F_comp, F)) {
        ocbFastComponents(F, C.comp);
        if (ocbScaffoldVerify()) {
            Eigen::VectorXi ref;
            igl::components(F, ref);
            if (ref.size() != C.comp.size() || (ref - C.comp).cwiseAbs().maxCoeff() != 0) { std::fprintf(stderr, "optcuts_b200: components differ from igl::components\n"); std::abort(); }
        }
        C.F_comp = F;
    }
    compI_V = C.comp;
}

}  // namespace OptCuts
#endif
