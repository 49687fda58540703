// TEST HARNESS (CPU): compiles the PRODUCT's device-inline element math (optcuts_b200/csrc/ocb_element.cuh)
// for the host so its formulas can be checked against the oracle without a GPU.  Built by
// tests/test_element_math_host.py into a scratch directory; never shipped, never imported by the package.
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __ldg(p) (*(p))
struct double2 { double x, y; };
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::sqrt; using std::fabs; using std::copysign; using std::fma;
#define OCB_ELEMENT_HOST 1
#include "../../optcuts_b200/csrc/ocb_element.cuh"

using namespace ocb;
extern "C" {
long host_general_path(int reset) { const long n = host_general_path_count; if (reset) host_general_path_count = 0; return n; }
// F: nF x 3 col-major; UV: nV x 2 col-major; rest8: 8 x nF
void host_energy(int nV, int nF, const int32_t* F, const double* UV, const double* rest8, double surf, int uniform, double* out)
{
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const Vec2 U1 = mk(UV[i0], UV[nV + i0]), U2 = mk(UV[i1], UV[nV + i1]), U3 = mk(UV[i2], UV[nV + i2]);
        const double w = uniform ? 1.0 : rest8[t] / surf;
        double db;
        out[t] = sd_energy(U2 - U1, U3 - U1, rest8[nF + t], rest8[2 * nF + t], rest8[3 * nF + t], rest8[4 * nF + t], w, db);
    }
}
void host_gradient_corners(int nV, int nF, const int32_t* F, const double* UV, const double* rest8, double surf, int uniform, double* out /*nF x 6*/)
{
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const Vec2 U1 = mk(UV[i0], UV[nV + i0]), U2 = mk(UV[i1], UV[nV + i1]), U3 = mk(UV[i2], UV[nV + i2]);
        const double w = uniform ? 1.0 : rest8[t] / surf;
        Vec2 g[3];
        sd_gradient(U1, U2, U3, rest8[nF + t], rest8[2 * nF + t], rest8[3 * nF + t], rest8[4 * nF + t], w, g);
        for (int k = 0; k < 3; ++k) { out[6 * t + 2 * k] = g[k].x; out[6 * t + 2 * k + 1] = g[k].y; }
    }
}
void host_hessian_blocks(int nV, int nF, const int32_t* F, const double* UV, const double* rest8, double surf, int uniform, int project,
                         double* out /*nF x 36 row-major*/, int* clamped)
{
    const int bOf[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const Vec2 U1 = mk(UV[i0], UV[nV + i0]), U2 = mk(UV[i1], UV[nV + i1]), U3 = mk(UV[i2], UV[nV + i2]);
        const double w = uniform ? 1.0 : rest8[t] / surf;
        double Hb[6][2][2];
        sd_hessian(U1, U2, U3, rest8[nF + t], rest8[5 * nF + t], rest8[6 * nF + t], rest8[7 * nF + t], w, Hb);
        int c = project ? sd_project_psd(Hb) : 0;
        if (clamped) clamped[t] = c;
        for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j)
            out[36 * t + (2 * k + i) * 6 + 2 * l + j] = (k <= l) ? Hb[bOf[k][l]][i][j] : Hb[bOf[k][l]][j][i];
    }
}
// the projection alone on caller-supplied symmetric 6x6 matrices (row-major, n x 36, in place); returns the clamp counts
void host_project_psd(int n, double* H36, int* clamped)
{
    const int bOf[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int t = 0; t < n; ++t) {
        double* H = H36 + 36 * (size_t)t;
        double Hb[6][2][2];
        for (int k = 0; k < 3; ++k) for (int l = k; l < 3; ++l) for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j)
            Hb[bOf[k][l]][i][j] = H[(2 * k + i) * 6 + 2 * l + j];
        const int c = sd_project_psd(Hb);
        if (clamped) clamped[t] = c;
        for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j)
            H[(2 * k + i) * 6 + 2 * l + j] = (k <= l) ? Hb[bOf[k][l]][i][j] : Hb[bOf[k][l]][j][i];
    }
}
// the vertex-gather gradient of grad_gather_kernel (ocb_kernels.cu), restated for the host with the SAME element function
// (sd_corner) and the same order: per vertex, incident corners in ascending triangle order.  Per-corner values vs the
// per-triangle function (sd_gradient): returns the number of corners whose bits differ.
long host_corner_vs_triangle(int nV, int nF, const int32_t* F, const double* UV, const double* rest8, double surf, int uniform)
{
    long bad = 0;
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const Vec2 U1 = mk(UV[i0], UV[nV + i0]), U2 = mk(UV[i1], UV[nV + i1]), U3 = mk(UV[i2], UV[nV + i2]);
        const double w = uniform ? 1.0 : rest8[t] / surf;
        Vec2 g[3];
        sd_gradient(U1, U2, U3, rest8[nF + t], rest8[2 * nF + t], rest8[3 * nF + t], rest8[4 * nF + t], w, g);
        double db0;
        const double E0 = sd_energy(U2 - U1, U3 - U1, rest8[nF + t], rest8[2 * nF + t], rest8[3 * nF + t], rest8[4 * nF + t], w, db0);
        for (int k = 0; k < 3; ++k) {
            Vec2 gk; double E, db;
            sd_corner(U1, U2, U3, rest8[nF + t], rest8[2 * nF + t], rest8[3 * nF + t], rest8[4 * nF + t], w, k, gk, E, db);
            if (std::memcmp(&gk.x, &g[k].x, 8) || std::memcmp(&gk.y, &g[k].y, 8) || std::memcmp(&E, &E0, 8) || std::memcmp(&db, &db0, 8)) ++bad;
        }
    }
    return bad;
}
// out: 2 * nV (interleaved), sums in ascending triangle order per vertex, unscaled, fixed vertices NOT masked
void host_gather_gradient(int nV, int nF, const int32_t* F, const double* UV, const double* rest8, double surf, int uniform, double* out)
{
    for (int v = 0; v < 2 * nV; ++v) out[v] = 0.0;
    for (int t = 0; t < nF; ++t) {       // ascending t per vertex == the order a vertex's corner list is walked in
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        const int idx[3] = {i0, i1, i2};
        const Vec2 U1 = mk(UV[i0], UV[nV + i0]), U2 = mk(UV[i1], UV[nV + i1]), U3 = mk(UV[i2], UV[nV + i2]);
        const double w = uniform ? 1.0 : rest8[t] / surf;
        for (int k = 0; k < 3; ++k) {
            Vec2 gk; double E, db;
            sd_corner(U1, U2, U3, rest8[nF + t], rest8[2 * nF + t], rest8[3 * nF + t], rest8[4 * nF + t], w, k, gk, E, db);
            out[2 * idx[k]] += gk.x; out[2 * idx[k] + 1] += gk.y;
        }
    }
}
// divgrad_gather_kernel restated for the host (same element function, same two walks per vertex)
void host_divgrad(int nV, int nF, const int32_t* F, const double* UV, const double* rest8, double surf, double* out)
{
    for (int v = 0; v < nV; ++v) {
        double mx = 0.0, my = 0.0, dev = 0.0; int n = 0;
        for (int pass = 0; pass < 2; ++pass) {
            n = 0;
            for (int t = 0; t < nF; ++t) for (int k = 0; k < 3; ++k) {
                if (F[k * nF + t] != v) continue;
                ++n;
                const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
                Vec2 gk; double E, db;
                sd_corner(mk(UV[i0], UV[nV + i0]), mk(UV[i1], UV[nV + i1]), mk(UV[i2], UV[nV + i2]), rest8[nF + t], rest8[2 * nF + t], rest8[3 * nF + t],
                          rest8[4 * nF + t], rest8[t] / surf, k, gk, E, db);
                if (pass == 0) { mx += gk.x; my += gk.y; }
                else { const double dx = gk.x - mx, dy = gk.y - my; dev += dx * dx + dy * dy; }
            }
            if (pass == 0) { mx /= n; my /= n; }
        }
        out[v] = (n <= 1) ? 0.0 : sqrt(dev / (n - 1.0));
    }
}
double host_step_bound(int nV, int nF, const int32_t* F, const double* UV, const double* dir /*interleaved*/, double alpha0)
{
    double cur = alpha0;
    for (int t = 0; t < nF; ++t) {
        const int i0 = F[t], i1 = F[nF + t], i2 = F[2 * nF + t];
        cur = sd_step_bound(mk(UV[i0], UV[nV + i0]), mk(UV[i1], UV[nV + i1]), mk(UV[i2], UV[nV + i2]),
                            mk(dir[2 * i0], dir[2 * i0 + 1]), mk(dir[2 * i1], dir[2 * i1 + 1]), mk(dir[2 * i2], dir[2 * i2 + 1]), cur);
    }
    return cur;
}
}
