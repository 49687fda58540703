/*
 * TEST INFRASTRUCTURE — CPU oracle, not on the product path.
 *
 * Plain-C restatement of the reference's algorithm for the OptCuts hot path, written from the
 * reference's behaviour; each function cites the reference file:line it follows.  It is PINNED
 * against the real reference (oracle/_ref/liboptcuts_ref.so, compiled from /root/reference) and
 * against the committed golden fixtures by tests/test_oracle_*.py.
 *
 * The linear solve restates Eigen::SimplicialLDLT (EigenLibSolver.cpp:71-107; vendored Eigen 3.3.4)
 * as a profile (skyline) LDL^T after a reverse Cuthill-McKee ordering: same factorisation family
 * (sparse LDL^T, no pivoting), different fill-reducing ordering (AMD there), hence results agree to
 * rounding (~1e-13 relative), not bitwise.  makePD restates Eigen::SelfAdjointEigenSolver as a
 * cyclic Jacobi eigen-solver on the 6x6 matrix (IglUtils.hpp:71-90).
 */
#include "sd_port.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define UVX(v) UV[(v)]
#define UVY(v) UV[nV + (v)]

/* TriMesh::computeFeatures, TriMesh.cpp:343-398 (per-triangle part) + igl::avg_edge_length */
int port_rest_features(int nV, int nF, const double* P, const int32_t* F, double thres, double* rest8, double* sc)
{
    double surf = 0.0, elen = 0.0; int bad = 0;
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        double a[3], b[3];
        for (int k = 0; k < 3; ++k) { a[k] = P[k * nV + i1] - P[k * nV + i0]; b[k] = P[k * nV + i2] - P[k * nV + i0]; }
        const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
        double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
        if (area == 0.0) bad = 1;
        double A2, e0, e1, d, k0, k1, kd;
        if (area < thres) {                               /* TriMesh.cpp:373-383 */
            area = thres; A2 = thres * thres;
            e0 = e1 = 4.0 / sqrt(3.0) * thres; d = e0 / 2.0;
            k0 = k1 = 2.0 / sqrt(3.0) / thres; kd = k0 / 2.0;
        } else {                                          /* :384-395 */
            A2 = area * area;
            e0 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; e1 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
            d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
            k0 = e0 / 2. / A2; k1 = e1 / 2. / A2; kd = d / 2. / A2;
        }
        surf += area;
        { double c[3]; for (int k = 0; k < 3; ++k) c[k] = b[k] - a[k];
          elen += sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) + sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]) + sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]); }
        rest8[t] = area; rest8[nF + t] = A2; rest8[2 * nF + t] = e0; rest8[3 * nF + t] = e1; rest8[4 * nF + t] = d;
        rest8[5 * nF + t] = k0; rest8[6 * nF + t] = k1; rest8[7 * nF + t] = kd;
    }
    sc[0] = surf; sc[1] = elen / (3.0 * nF); sc[2] = sqrt(surf / M_PI);
    return bad ? -1 : 0;
}

/* SymDirichletEnergy::getEnergyValPerElem, SymDirichletEnergy.cpp:24-46 */
void port_energy_per_elem(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, int uniform, double* out)
{
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const double ux = UVX(i1) - UVX(i0), uy = UVY(i1) - UVY(i0), vx = UVX(i2) - UVX(i0), vy = UVY(i2) - UVY(i0);
        const double area_U = 0.5 * (ux * vy - uy * vx);
        const double A2 = r8[nF + t], e0 = r8[2 * nF + t], e1 = r8[3 * nF + t], d = r8[4 * nF + t];
        const double w = uniform ? 1.0 : r8[t] / surf;
        out[t] = w * (1.0 + A2 / area_U / area_U) * (((vx * vx + vy * vy) * e0 + (ux * ux + uy * uy) * e1) / 4 / A2 - (vx * ux + vy * uy) * d / 2 / A2);
    }
}
/* Energy::computeEnergyVal, Energy.cpp:35-40 (Eigen's VectorXd::sum() pairwise order is not restated: plain sum) */
double port_energy(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, int uniform)
{
    double* e = (double*)malloc(sizeof(double) * (size_t)nF), s = 0.0;
    port_energy_per_elem(nV, nF, F, UV, r8, surf, uniform, e);
    for (int t = 0; t < nF; ++t) s += e[t];
    free(e);
    return s;
}

static void corner_gradients(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, int uniform, int t, double g[6])
{
    const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
    const double U1x = UVX(i0), U1y = UVY(i0), U2x = UVX(i1), U2y = UVY(i1), U3x = UVX(i2), U3y = UVY(i2);
    const double ux = U2x - U1x, uy = U2y - U1y, vx = U3x - U1x, vy = U3y - U1y;
    const double area_U = 0.5 * (ux * vy - uy * vx);
    const double A2 = r8[nF + t], e0 = r8[2 * nF + t], e1 = r8[3 * nF + t], d = r8[4 * nF + t];
    const double left = 1.0 + A2 / area_U / area_U;
    const double right = ((vx * vx + vy * vy) * e0 + (ux * ux + uy * uy) * e1) / 4 / A2 - (vx * ux + vy * uy) * d / 2 / A2;
    const double ar = A2 / area_U / area_U / area_U;
    const double w = uniform ? 1.0 : r8[t] / surf;
    /* opposite edges and their (y,-x) rotations, SymDirichletEnergy.cpp:283-297 */
    const double o1x = U3x - U2x, o1y = U3y - U2y, o2x = U1x - U3x, o2y = U1y - U3y, o3x = U2x - U1x, o3y = U2y - U1y;
    const double dL[6] = {ar * o1y, ar * -o1x, ar * o2y, ar * -o2x, ar * o3y, ar * -o3x};
    const double dR[6] = {((d - e0) * vx + (d - e1) * ux) / 2.0 / A2, ((d - e0) * vy + (d - e1) * uy) / 2.0 / A2,
                          (e1 * ux - d * vx) / 2.0 / A2, (e1 * uy - d * vy) / 2.0 / A2,
                          (e0 * vx - d * ux) / 2.0 / A2, (e0 * vy - d * uy) / 2.0 / A2};
    for (int k = 0; k < 6; ++k) g[k] = w * (dL[k] * right + dR[k] * left);
}
/* SymDirichletEnergy::computeGradient, SymDirichletEnergy.cpp:258-304 */
void port_gradient(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, int uniform,
                   const int32_t* fixed, int nFixed, double* g)
{
    memset(g, 0, sizeof(double) * 2 * (size_t)nV);
    for (int t = 0; t < nF; ++t) {
        double c[6]; corner_gradients(nV, nF, F, UV, r8, surf, uniform, t, c);
        for (int k = 0; k < 3; ++k) { const int v = F[k * nF + t]; g[2 * v] += c[2 * k]; g[2 * v + 1] += c[2 * k + 1]; }
    }
    for (int i = 0; i < nFixed; ++i) { g[2 * fixed[i]] = 0.0; g[2 * fixed[i] + 1] = 0.0; }
}

/* IglUtils::makePD<double,6>, IglUtils.hpp:71-90: clamp negative eigenvalues, untouched if lambda_min >= 0.
 * Eigen's tridiagonal QL is restated as a cyclic Jacobi sweep (same eigen-decomposition, to rounding). */
void port_make_pd6(double* M)
{
    double A[6][6], Q[6][6];
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { A[i][j] = 0.5 * (M[i * 6 + j] + M[j * 6 + i]); Q[i][j] = (i == j); }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = 0; i < 6; ++i) { dg += A[i][i] * A[i][i]; for (int j = i + 1; j < 6; ++j) off += A[i][j] * A[i][j]; }
        if (off <= 1e-36 * dg || off == 0.0) break;
        for (int p = 0; p < 5; ++p) for (int q = p + 1; q < 6; ++q) {
            const double apq = A[p][q];
            if (apq == 0.0) continue;
            const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(tt * tt + 1.0), s = tt * c;
            for (int r = 0; r < 6; ++r) { const double arp = A[r][p], arq = A[r][q]; A[r][p] = c * arp - s * arq; A[r][q] = s * arp + c * arq; }
            for (int r = 0; r < 6; ++r) { const double apr = A[p][r], aqr = A[q][r]; A[p][r] = c * apr - s * aqr; A[q][r] = s * apr + c * aqr; }
            for (int r = 0; r < 6; ++r) { const double qrp = Q[r][p], qrq = Q[r][q]; Q[r][p] = c * qrp - s * qrq; Q[r][q] = s * qrp + c * qrq; }
        }
    }
    double lmin = A[0][0];
    for (int i = 1; i < 6; ++i) if (A[i][i] < lmin) lmin = A[i][i];
    if (lmin >= 0.0) return;
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) {
        double s = 0.0;
        for (int e = 0; e < 6; ++e) { const double lam = A[e][e] < 0.0 ? 0.0 : A[e][e]; s += Q[i][e] * lam * Q[j][e]; }
        M[i * 6 + j] = s;
    }
}

/* SymDirichletEnergy::computeHessian element block, SymDirichletEnergy.cpp:440-525 (row-major 6x6) */
static void element_hessian(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, int uniform, int t, int project, double* H)
{
    const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
    const double U[3][2] = {{UVX(i0), UVY(i0)}, {UVX(i1), UVY(i1)}, {UVX(i2), UVY(i2)}};
    const double u[2] = {U[1][0] - U[0][0], U[1][1] - U[0][1]}, v[2] = {U[2][0] - U[0][0], U[2][1] - U[0][1]};
    const double area_U = 0.5 * (u[0] * v[1] - u[1] * v[0]);
    const double A2 = r8[nF + t], k0 = r8[5 * nF + t], k1 = r8[6 * nF + t], kd = r8[7 * nF + t];
    const double ar = A2 / area_U / area_U / area_U, mult = 3.0 / 2.0 * ar / area_U;
    const double w = uniform ? 1.0 : r8[t] / surf;
    const double left = 1.0 + A2 / area_U / area_U;
    const double right = ((v[0] * v[0] + v[1] * v[1]) * k0 + (u[0] * u[0] + u[1] * u[1]) * k1) / 2. - (v[0] * u[0] + v[1] * u[1]) * kd;
    double n[3][2], dL[3][2], dR[3][2];
    const double opp[3][2] = {{U[2][0] - U[1][0], U[2][1] - U[1][1]}, {U[0][0] - U[2][0], U[0][1] - U[2][1]}, {U[1][0] - U[0][0], U[1][1] - U[0][1]}};
    for (int k = 0; k < 3; ++k) { n[k][0] = opp[k][1]; n[k][1] = -opp[k][0]; dL[k][0] = ar * n[k][0]; dL[k][1] = ar * n[k][1]; }
    for (int c = 0; c < 2; ++c) {
        dR[0][c] = (kd - k0) * v[c] + (kd - k1) * u[c];
        dR[1][c] = k1 * u[c] - kd * v[c];
        dR[2][c] = k0 * v[c] - kd * u[c];
    }
    /* d2Right_kl (scalar times identity) and the sign of the +-ar*[[0,-1],[1,0]] term, :485-522 */
    const double d2R[3][3] = {{k0 + k1 - 2.0 * kd, kd - k1, kd - k0}, {kd - k1, k1, -kd}, {kd - k0, -kd, k0}};
    const double sgn[3][3] = {{0, 1, -1}, {-1, 0, 1}, {1, -1, 0}};   /* coefficient of dOrtho_div_dU = [[0,-1],[1,0]] */
    const double dO[2][2] = {{0.0, -1.0}, {1.0, 0.0}};
    for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
        const double d2L = mult * n[k][i] * n[l][j] + ar * sgn[k][l] * dO[i][j];
        H[(2 * k + i) * 6 + 2 * l + j] = w * (d2L * right + dL[k][i] * dR[l][j] + d2R[k][l] * left * (i == j) + dR[k][i] * dL[l][j]);
    }
    if (project) port_make_pd6(H);
}
void port_hessian_blocks(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, int uniform, int project, double* out)
{
    for (int t = 0; t < nF; ++t) element_hessian(nV, nF, F, UV, r8, surf, uniform, t, project, out + 36 * (size_t)t);
}
/* triplet stream: IglUtils::addBlockToMatrix (IglUtils.cpp:396-451) per triangle, then
 * addDiagonalToMatrix for the fixed vertices (SymDirichletEnergy.cpp:536-548).  V==NULL: count only. */
int64_t port_hessian_triplets(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, int uniform,
                              const int32_t* fixed, int nFixed, double* V, int32_t* I, int32_t* J)
{
    char* isFixed = (char*)calloc((size_t)nV, 1);
    for (int i = 0; i < nFixed; ++i) isFixed[fixed[i]] = 1;
    int64_t w = 0;
    for (int t = 0; t < nF; ++t) {
        int idx[3]; double H[36];
        for (int k = 0; k < 3; ++k) { idx[k] = F[k * nF + t]; if (isFixed[idx[k]]) idx[k] = -1; }
        if (V) element_hessian(nV, nF, F, UV, r8, surf, uniform, t, 1, H);
        for (int a = 0; a < 3; ++a) { if (idx[a] < 0) continue;
            for (int b = 0; b < 3; ++b) { if (idx[b] < 0) continue;
                for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) {
                    if (V) { V[w] = H[(2 * a + i) * 6 + 2 * b + j]; I[w] = 2 * idx[a] + i; J[w] = 2 * idx[b] + j; }
                    ++w;
                } } }
    }
    /* std::set iteration order = ascending */
    for (int v = 0; v < nV; ++v) if (isFixed[v]) for (int i = 0; i < 2; ++i) { if (V) { V[w] = 1.0; I[w] = J[w] = 2 * v + i; } ++w; }
    free(isFixed);
    return w;
}

/* SymDirichletEnergy::initStepSize, SymDirichletEnergy.cpp:551-610 */
double port_init_step_size(int nV, int nF, const int32_t* F, const double* UV, const double* p, double stepSize)
{
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const double ux = UVX(i1) - UVX(i0), uy = UVY(i1) - UVY(i0), vx = UVX(i2) - UVX(i0), vy = UVY(i2) - UVY(i0);
        const double dux = p[2 * i1] - p[2 * i0], duy = p[2 * i1 + 1] - p[2 * i0 + 1], dvx = p[2 * i2] - p[2 * i0], dvy = p[2 * i2 + 1] - p[2 * i0 + 1];
        const double a = dux * dvy - duy * dvx;
        const double b = ux * dvy - uy * dvx + dux * vy - duy * vx;
        const double c = ux * vy - uy * vx;
        const double delta = b * b - 4.0 * a * c;
        double bound = stepSize;
        if (a > 0.0) { if ((b < 0.0) && (delta >= 0.0)) bound = 2.0 * c / (-b + sqrt(delta)); }
        else if (a < 0.0) { if (b < 0.0) bound = 2.0 * c / (-b + sqrt(delta)); else bound = (-b - sqrt(delta)) / 2.0 / a; }
        else if (b < 0.0) bound = -c / b;
        if (bound < stepSize) stepSize = bound;
    }
    return stepSize;
}

/* LinSysSolver::set_pattern, LinSysSolver.hpp:37-135: 1-based upper-triangular CSR, 2 rows per vertex,
 * fixed vertices get a lone diagonal.  ia/ja NULL: return nnz only. */
int64_t port_set_pattern(int nV, const int32_t* adjPtr, const int32_t* adjIdx, const int32_t* fixed, int nFixed, int32_t* ia, int32_t* ja)
{
    char* isFixed = (char*)calloc((size_t)nV, 1);
    for (int i = 0; i < nFixed; ++i) isFixed[fixed[i]] = 1;
    int64_t w = 0;
    if (ia) ia[0] = 1;
    for (int r = 0; r < nV; ++r) {
        if (!isFixed[r]) {
            for (int row = 0; row < 2; ++row) {
                /* own block: x-row holds (2r,2r+1), y-row holds (2r+1) only (:85-106) */
                for (int q = row; q < 2; ++q) { if (ja) ja[w] = 2 * r + q + 1; ++w; }
                for (int k = adjPtr[r]; k < adjPtr[r + 1]; ++k) {
                    const int c = adjIdx[k];
                    if (isFixed[c] || c <= r) continue;
                    for (int q = 0; q < 2; ++q) { if (ja) ja[w] = 2 * c + q + 1; ++w; }
                }
                if (ia) ia[2 * r + row + 1] = (int32_t)(w + 1);
            }
        } else {
            for (int row = 0; row < 2; ++row) { if (ja) ja[w] = 2 * r + row + 1; ++w; if (ia) ia[2 * r + row + 1] = (int32_t)(w + 1); }
        }
    }
    free(isFixed);
    return w;
}

/* LinSysSolver::update_a, LinSysSolver.hpp:138-159: zero, accumulate triplets with i <= j.
 * (std::map lookup restated as a binary search in the sorted row.)  returns #triplets outside the pattern */
int port_update_a(int n, const int32_t* ia, const int32_t* ja, int64_t nT, const int32_t* I, const int32_t* J, const double* S, double* a)
{
    int miss = 0;
    memset(a, 0, sizeof(double) * (size_t)(ia[n] - 1));
    for (int64_t k = 0; k < nT; ++k) {
        const int i = I[k], j = J[k];
        if (i > j) continue;
        int lo = ia[i] - 1, hi = ia[i + 1] - 2, s = -1;
        while (lo <= hi) { const int mid = (lo + hi) / 2, c = ja[mid] - 1; if (c == j) { s = mid; break; } if (c < j) lo = mid + 1; else hi = mid - 1; }
        if (s < 0) { ++miss; continue; }
        a[s] += S[k];
    }
    return miss;
}

/* EigenLibSolver::analyze_pattern/factorize/solve (EigenLibSolver.cpp:71-107) restated as RCM + skyline LDL^T */
int port_ldlt_solve(int n, const int32_t* ia, const int32_t* ja, const double* a, const double* rhs, double* x)
{
    /* symmetric adjacency of scalar rows */
    int* deg = (int*)calloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n; ++i) for (int k = ia[i] - 1; k < ia[i + 1] - 1; ++k) { const int j = ja[k] - 1; if (j != i) { deg[i + 1]++; deg[j + 1]++; } }
    for (int i = 0; i < n; ++i) deg[i + 1] += deg[i];
    int* adj = (int*)malloc(sizeof(int) * (size_t)(deg[n] > 0 ? deg[n] : 1));
    int* fill = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) fill[i] = deg[i];
    for (int i = 0; i < n; ++i) for (int k = ia[i] - 1; k < ia[i + 1] - 1; ++k) { const int j = ja[k] - 1; if (j != i) { adj[fill[i]++] = j; adj[fill[j]++] = i; } }
    /* Cuthill-McKee BFS from a minimum-degree vertex of every component, then reversed */
    int* order = (int*)malloc(sizeof(int) * (size_t)n); int* perm = (int*)malloc(sizeof(int) * (size_t)n);
    char* seen = (char*)calloc((size_t)n, 1);
    int cnt = 0;
    for (;;) {
        int start = -1;
        for (int i = 0; i < n; ++i) if (!seen[i] && (start < 0 || deg[i + 1] - deg[i] < deg[start + 1] - deg[start])) start = i;
        if (start < 0) break;
        int head = cnt; order[cnt++] = start; seen[start] = 1;
        while (head < cnt) {
            const int v = order[head++];
            const int b = cnt;
            for (int k = deg[v]; k < deg[v + 1]; ++k) { const int u = adj[k]; if (!seen[u]) { seen[u] = 1; order[cnt++] = u; } }
            /* sort the newly added by degree (insertion sort, lists are short) */
            for (int p = b + 1; p < cnt; ++p) { const int u = order[p]; const int du = deg[u + 1] - deg[u]; int q = p - 1;
                while (q >= b && deg[order[q] + 1] - deg[order[q]] > du) { order[q + 1] = order[q]; --q; } order[q + 1] = u; }
        }
    }
    for (int i = 0; i < n; ++i) perm[order[n - 1 - i]] = i;       /* old -> new */
    /* skyline of the permuted LOWER triangle: row i stores columns first[i]..i */
    int* first = (int*)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) first[i] = i;
    for (int i = 0; i < n; ++i) for (int k = ia[i] - 1; k < ia[i + 1] - 1; ++k) {
        int r = perm[i], c = perm[ja[k] - 1]; if (r < c) { const int tmp = r; r = c; c = tmp; }
        if (c < first[r]) first[r] = c;
    }
    size_t* off = (size_t*)malloc(sizeof(size_t) * ((size_t)n + 1));
    off[0] = 0;
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + (size_t)(i - first[i] + 1);
    double* Lv = (double*)calloc(off[n], sizeof(double));
#define LL(i, j) Lv[off[i] + (size_t)((j) - first[i])]
    for (int i = 0; i < n; ++i) for (int k = ia[i] - 1; k < ia[i + 1] - 1; ++k) {
        int r = perm[i], c = perm[ja[k] - 1]; if (r < c) { const int tmp = r; r = c; c = tmp; }
        LL(r, c) += a[k];
    }
    /* in-place LDL^T: L(i,j) for j<i, D on the diagonal */
    int ok = 1;
    for (int i = 0; i < n; ++i) {
        for (int j = first[i]; j < i; ++j) {
            double s = LL(i, j);
            const int k0 = first[i] > first[j] ? first[i] : first[j];
            for (int k = k0; k < j; ++k) s -= LL(i, k) * LL(j, k);      /* LL(i,k) still holds L*D here */
            LL(i, j) = s;
        }
        double dsum = LL(i, i);
        for (int j = first[i]; j < i; ++j) { const double ld = LL(i, j); const double l = ld / LL(j, j); dsum -= ld * l; LL(i, j) = l; }
        LL(i, i) = dsum;
        if (!(dsum != 0.0)) ok = 0;
    }
    double* y = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) y[perm[i]] = rhs[i];
    for (int i = 0; i < n; ++i) { double s = y[i]; for (int j = first[i]; j < i; ++j) s -= LL(i, j) * y[j]; y[i] = s; }
    for (int i = 0; i < n; ++i) y[i] /= LL(i, i);
    for (int i = n - 1; i >= 0; --i) { const double yi = y[i]; for (int j = first[i]; j < i; ++j) y[j] -= LL(i, j) * yi; }
    for (int i = 0; i < n; ++i) x[i] = y[perm[i]];
#undef LL
    free(deg); free(adj); free(fill); free(order); free(perm); free(seen); free(first); free(off); free(Lv); free(y);
    return ok ? 0 : -1;
}

/* TriMesh::computeSeamSparsity, TriMesh.cpp:1542-1558 */
double port_seam_sparsity(int nV, int nCoh, const int32_t* cohE, const int32_t* bnd, const double* len, const double* UV,
                          double avgEdgeLen, double initSeamLen, int triSoup)
{
    const double thres = 1.0e-2; double s = 0.0;
    for (int c = 0; c < nCoh; ++c) {
        if (bnd[c]) continue;
        int take = !triSoup;
        if (!take) {
            const int a0 = cohE[c], a1 = cohE[nCoh + c], a2 = cohE[2 * nCoh + c], a3 = cohE[3 * nCoh + c];
            const double d0 = hypot(UVX(a0) - UVX(a2), UVY(a0) - UVY(a2)), d1 = hypot(UVX(a1) - UVX(a3), UVY(a1) - UVY(a3));
            take = (sqrt((UVX(a0) - UVX(a2)) * (UVX(a0) - UVX(a2)) + (UVY(a0) - UVY(a2)) * (UVY(a0) - UVY(a2))) / avgEdgeLen > thres) ||
                   (sqrt((UVX(a1) - UVX(a3)) * (UVX(a1) - UVX(a3)) + (UVY(a1) - UVY(a3)) * (UVY(a1) - UVY(a3))) / avgEdgeLen > thres);
            (void)d0; (void)d1;
        }
        if (take) s += len[c];
    }
    return s + initSeamLen;
}

/* SymDirichletEnergy::computeLocalGradient + computeDivGradPerVert, SymDirichletEnergy.cpp:108-149, 215-256 */
void port_divgrad(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf, double* out)
{
    double* mean = (double*)calloc(2 * (size_t)nV, sizeof(double)); int* cnt = (int*)calloc((size_t)nV, sizeof(int));
    double* dev = (double*)calloc((size_t)nV, sizeof(double)); double* loc = (double*)malloc(sizeof(double) * 6 * (size_t)nF);
    for (int t = 0; t < nF; ++t) { corner_gradients(nV, nF, F, UV, r8, surf, 0, t, loc + 6 * (size_t)t);
        for (int k = 0; k < 3; ++k) { const int v = F[k * nF + t]; mean[2 * v] += loc[6 * t + 2 * k]; mean[2 * v + 1] += loc[6 * t + 2 * k + 1]; cnt[v]++; } }
    for (int v = 0; v < nV; ++v) { mean[2 * v] /= cnt[v]; mean[2 * v + 1] /= cnt[v]; }
    for (int t = 0; t < nF; ++t) for (int k = 0; k < 3; ++k) { const int v = F[k * nF + t];
        const double dx = loc[6 * t + 2 * k] - mean[2 * v], dy = loc[6 * t + 2 * k + 1] - mean[2 * v + 1]; dev[v] += dx * dx + dy * dy; }
    for (int v = 0; v < nV; ++v) out[v] = (cnt[v] == 1) ? 0.0 : sqrt(dev[v] / (cnt[v] - 1.0));
    free(mean); free(cnt); free(dev); free(loc);
}

/* ---- one geometry step: Optimizer::solve(1) -> solve_oneStep -> lineSearch (Optimizer.cpp:203-261, 505-673),
 * energy/gradient/Hessian combination as Optimizer.cpp:764-843 + Scaffold.cpp:210-293 ---- */
static int cmp_i32(const void* a, const void* b) { return (*(const int32_t*)a > *(const int32_t*)b) - (*(const int32_t*)a < *(const int32_t*)b); }

static double total_energy(int nV, int nF, const int32_t* F, const double* UV, const double* r8, double surf,
                           int nVa, int nFa, const int32_t* Fa, const double* UVa, const double* r8a,
                           double p0, double w_scaf, double* Esd, double* Escaf)
{
    *Esd = port_energy(nV, nF, F, UV, r8, surf, 0);
    *Escaf = nFa > 0 ? port_energy(nVa, nFa, Fa, UVa, r8a, 1.0, 1) * (w_scaf / nFa) : 0.0;
    return p0 * *Esd + *Escaf;
}

int port_newton_step(int nV, int nF, const int32_t* F, double* UV, const double* r8, double surf, const int32_t* fixed, int nFixed,
                     int nVa, int nFa, const int32_t* Fa, double* UVa, const double* r8a, const int32_t* l2g, int nBnd,
                     const int32_t* fixedAir, int nFixedAir, double p0, double w_scaf, double targetGRes, int allowEDecRelTol,
                     double* searchDir_out, port_newton_result* out)
{
    const int scaf = nFa > 0;
    const int nVtot = nV + (scaf ? nVa - nBnd : 0), n = 2 * nVtot;
    memset(out, 0, sizeof(*out));
    /* gradient: Optimizer::computeGradient + Scaffold::augmentGradient */
    double* g = (double*)calloc((size_t)n, sizeof(double));
    port_gradient(nV, nF, F, UV, r8, surf, 0, fixed, nFixed, g);
    for (int i = 0; i < 2 * nV; ++i) g[i] = p0 * g[i];
    if (scaf) {
        double* ga = (double*)malloc(sizeof(double) * 2 * (size_t)nVa);
        port_gradient(nVa, nFa, Fa, UVa, r8a, 1.0, 1, fixedAir, nFixedAir, ga);
        const double ws = w_scaf / nFa;
        for (int v = 0; v < nVa; ++v) { g[2 * l2g[v]] += ws * ga[2 * v]; g[2 * l2g[v] + 1] += ws * ga[2 * v + 1]; }
        free(ga);
    }
    double sqn = 0.0; for (int i = 0; i < n; ++i) sqn += g[i] * g[i];
    out->sqn_g = sqn;
    if (sqn < targetGRes) { out->converged = 1; free(g); return 1; }
    /* merged adjacency + fixed set (Scaffold::mergeVNeighbor / mergeFixedV) */
    int32_t* deg = (int32_t*)calloc((size_t)nVtot + 1, sizeof(int32_t));
    for (int t = 0; t < nF; ++t) for (int k = 0; k < 3; ++k) deg[F[k * nF + t] + 1] += 2;
    for (int t = 0; t < nFa; ++t) for (int k = 0; k < 3; ++k) deg[l2g[Fa[k * nFa + t]] + 1] += 2;
    for (int v = 0; v < nVtot; ++v) deg[v + 1] += deg[v];
    int32_t* buf = (int32_t*)malloc(sizeof(int32_t) * (size_t)(deg[nVtot] + 1)); int32_t* fl = (int32_t*)malloc(sizeof(int32_t) * (size_t)nVtot);
    for (int v = 0; v < nVtot; ++v) fl[v] = deg[v];
    for (int t = 0; t < nF; ++t) { const int a = F[t], b = F[nF + t], c = F[2 * nF + t];
        buf[fl[a]++] = b; buf[fl[a]++] = c; buf[fl[b]++] = a; buf[fl[b]++] = c; buf[fl[c]++] = a; buf[fl[c]++] = b; }
    for (int t = 0; t < nFa; ++t) { const int a = l2g[Fa[t]], b = l2g[Fa[nFa + t]], c = l2g[Fa[2 * nFa + t]];
        buf[fl[a]++] = b; buf[fl[a]++] = c; buf[fl[b]++] = a; buf[fl[b]++] = c; buf[fl[c]++] = a; buf[fl[c]++] = b; }
    int32_t* adjPtr = (int32_t*)calloc((size_t)nVtot + 1, sizeof(int32_t)); int32_t* adjIdx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(deg[nVtot] + 1));
    int w = 0;
    for (int v = 0; v < nVtot; ++v) { qsort(buf + deg[v], (size_t)(fl[v] - deg[v]), sizeof(int32_t), cmp_i32);
        for (int k = deg[v]; k < fl[v]; ++k) if (k == deg[v] || buf[k] != buf[k - 1]) adjIdx[w++] = buf[k];
        adjPtr[v + 1] = w; }
    int32_t* fx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nFixed + nFixedAir + 1)); int nfx = 0;
    for (int i = 0; i < nFixed; ++i) fx[nfx++] = fixed[i];
    for (int i = 0; i < nFixedAir; ++i) { int dup = 0; for (int k = 0; k < nfx; ++k) if (fx[k] == l2g[fixedAir[i]]) dup = 1; if (!dup) fx[nfx++] = l2g[fixedAir[i]]; }
    const int64_t nnz = port_set_pattern(nVtot, adjPtr, adjIdx, fx, nfx, NULL, NULL);
    int32_t* ia = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1)); int32_t* ja = (int32_t*)malloc(sizeof(int32_t) * (size_t)nnz);
    double* a = (double*)malloc(sizeof(double) * (size_t)nnz);
    port_set_pattern(nVtot, adjPtr, adjIdx, fx, nfx, ia, ja);
    /* Hessian triplets: mesh * p0, air * w_scaf/|Fa| remapped (Optimizer::computeHessian, Scaffold::augmentProxyMatrix) */
    const int64_t nTm = port_hessian_triplets(nV, nF, F, UV, r8, surf, 0, fixed, nFixed, NULL, NULL, NULL);
    const int64_t nTa = scaf ? port_hessian_triplets(nVa, nFa, Fa, UVa, r8a, 1.0, 1, fixedAir, nFixedAir, NULL, NULL, NULL) : 0;
    double* TV = (double*)malloc(sizeof(double) * (size_t)(nTm + nTa)); int32_t* TI = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nTm + nTa)); int32_t* TJ = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nTm + nTa));
    port_hessian_triplets(nV, nF, F, UV, r8, surf, 0, fixed, nFixed, TV, TI, TJ);
    for (int64_t k = 0; k < nTm; ++k) TV[k] *= p0;
    if (scaf) {
        port_hessian_triplets(nVa, nFa, Fa, UVa, r8a, 1.0, 1, fixedAir, nFixedAir, TV + nTm, TI + nTm, TJ + nTm);
        const double ws = w_scaf / nFa;
        for (int64_t k = nTm; k < nTm + nTa; ++k) { TV[k] = ws * TV[k]; TI[k] = l2g[TI[k] / 2] * 2 + TI[k] % 2; TJ[k] = l2g[TJ[k] / 2] * 2 + TJ[k] % 2; }
    }
    port_update_a(n, ia, ja, nTm + nTa, TI, TJ, TV, a);
    double* rhs = (double*)malloc(sizeof(double) * (size_t)n); double* p = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) rhs[i] = -g[i];
    port_ldlt_solve(n, ia, ja, a, rhs, p);
    if (searchDir_out) memcpy(searchDir_out, p, sizeof(double) * (size_t)n);
    /* line search: Optimizer::lineSearch, Optimizer.cpp:575-652 */
    double step = 1.0;
    step = port_init_step_size(nV, nF, F, UV, p, step);
    double* pa = NULL;
    if (scaf) { pa = (double*)malloc(sizeof(double) * 2 * (size_t)nVa);
        for (int v = 0; v < nVa; ++v) { pa[2 * v] = p[2 * l2g[v]]; pa[2 * v + 1] = p[2 * l2g[v] + 1]; }
        step = port_init_step_size(nVa, nFa, Fa, UVa, pa, step); }
    step *= 0.99;
    double* UV0 = (double*)malloc(sizeof(double) * 2 * (size_t)nV); memcpy(UV0, UV, sizeof(double) * 2 * (size_t)nV);
    double* UVa0 = NULL; if (scaf) { UVa0 = (double*)malloc(sizeof(double) * 2 * (size_t)nVa); memcpy(UVa0, UVa, sizeof(double) * 2 * (size_t)nVa); }
    double Esd, Escaf, lastScaf;
    double Elast = total_energy(nV, nF, F, UV, r8, surf, nVa, nFa, Fa, UVa, r8a, p0, w_scaf, &Esd, &lastScaf);
    double E; int halv = 0, stopped = 0;
    for (;;) {
        for (int v = 0; v < nV; ++v) { UV[v] = UV0[v] + step * p[2 * v]; UV[nV + v] = UV0[nV + v] + step * p[2 * v + 1]; }
        if (scaf) for (int v = 0; v < nVa; ++v) { UVa[v] = UVa0[v] + step * pa[2 * v]; UVa[nVa + v] = UVa0[nVa + v] + step * pa[2 * v + 1]; }
        E = total_energy(nV, nF, F, UV, r8, surf, nVa, nFa, Fa, UVa, r8a, p0, w_scaf, &Esd, &Escaf);
        if (E > Elast) { step /= 2.0; ++halv; if (step == 0.0) { stopped = 1; break; } continue; }
        break;
    }
    double eDec = Elast - E;
    if (scaf) eDec += (-lastScaf + Escaf);
    if (allowEDecRelTol && (eDec / Elast < 1.0e-6 * step) && (step > 1.0e-3)) stopped = 1;
    out->alpha = step; out->E_new = E; out->E_scaf_new = Escaf; out->E_sd_new = Esd; out->lastEDec = eDec; out->E_last = Elast;
    out->n_halvings = halv; out->stopped = stopped;
    free(g); free(deg); free(buf); free(fl); free(adjPtr); free(adjIdx); free(fx); free(ia); free(ja); free(a);
    free(TV); free(TI); free(TJ); free(rhs); free(p); free(pa); free(UV0); free(UVa0);
    return stopped;
}
