// optcuts_b200 — per-triangle Symmetric Dirichlet math (device inline functions).
//
// For a triangle with rest edges e0 = P2-P1, e1 = P3-P1, rest area A, UV edges u = U2-U1,
// v = U3-U1 and signed UV area a = (u x v)/2:
//     R = (|v|^2 |e0|^2 + |u|^2 |e1|^2)/(4A^2) - (u.v)(e0.e1)/(2A^2)  = ||J||_F^2
//     L = 1 + A^2/a^2                                                   = 1 + det(J)^-2
//     E_t = w L R
// Reference behaviour being reproduced (not its code): SymDirichletEnergy.cpp:24-46 (value),
// :258-304 (gradient), :429-549 (Hessian), IglUtils.hpp:71-90 (makePD), :551-610 (step bound).
#pragma once
#ifndef OCB_ELEMENT_HOST
#include <cuda_runtime.h>
#endif

namespace ocb {

struct Vec2 { double x, y; };
__device__ __forceinline__ Vec2 mk(double x, double y) { Vec2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ Vec2 operator-(Vec2 a, Vec2 b) { return mk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ Vec2 operator+(Vec2 a, Vec2 b) { return mk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ Vec2 operator*(double s, Vec2 a) { return mk(s * a.x, s * a.y); }
__device__ __forceinline__ double dot(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ double cross(Vec2 a, Vec2 b) { return a.x * b.y - a.y * b.x; }
__device__ __forceinline__ Vec2 perp(Vec2 a) { return mk(a.y, -a.x); }   // (x,y) -> (y,-x)

__device__ __forceinline__ Vec2 ld2(const double* x, int v) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(x) + v);
    return mk(t.x, t.y);
}
// position used by the line search: x0 + alpha * p with NO fma contraction, so that it is the
// same double the CPU computes (Optimizer.cpp:666-668)
__device__ __forceinline__ Vec2 ld2_step(const double* x0, const double* p, double alpha, int v) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(x0) + v);
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + v);
    return mk(__dadd_rn(a.x, __dmul_rn(alpha, b.x)), __dadd_rn(a.y, __dmul_rn(alpha, b.y)));
}

// ---- rest-frame features of one triangle (TriMesh::computeFeatures arithmetic, TriMesh.cpp:355-398): out8 = area, areaSq,
// |e0|^2, |e1|^2, e0.e1, and the three x / (2 A^2); triangles below `thres` (air-mesh degeneracy guard) get the
// equilateral surrogate (:373-383).  a = P2 - P1, b = P3 - P1 (3-D; the air mesh has z = 0).  Returns the raw area.
__device__ __forceinline__ double sd_rest_features(double ax, double ay, double az, double bx, double by, double bz, double thres, double out8[8])
{
    const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    const double raw = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
    double area = raw, A2, e0, e1, d, k0, k1, kd;
    if (area < thres) {
        const double sqrt3 = sqrt(3.0);
        area = thres; A2 = thres * thres;
        e0 = e1 = 4.0 / sqrt3 * thres; d = e0 / 2.0;
        k0 = k1 = 2.0 / sqrt3 / thres; kd = k0 / 2.0;
    } else {
        A2 = area * area;
        e0 = ax * ax + ay * ay + az * az; e1 = bx * bx + by * by + bz * bz; d = ax * bx + ay * by + az * bz;
        k0 = e0 / 2. / A2; k1 = e1 / 2. / A2; kd = d / 2. / A2;
    }
    out8[0] = area; out8[1] = A2; out8[2] = e0; out8[3] = e1; out8[4] = d; out8[5] = k0; out8[6] = k1; out8[7] = kd;
    return raw;
}

// ---- value (division order as the reference's value/gradient functions) -----------------------
__device__ __forceinline__ double sd_energy(Vec2 u, Vec2 v, double A2, double e0, double e1, double d,
                                            double w, double& dbArea)
{
    dbArea = cross(u, v);
    const double a = 0.5 * dbArea;
    const double L = 1.0 + A2 / a / a;
    const double R = (dot(v, v) * e0 + dot(u, u) * e1) / 4 / A2 - dot(v, u) * d / 2 / A2;
    return w * L * R;
}

// ---- gradient w.r.t. U1,U2,U3 -----------------------------------------------------------------
__device__ __forceinline__ void sd_gradient(Vec2 U1, Vec2 U2, Vec2 U3, double A2, double e0, double e1,
                                            double d, double w, Vec2 g[3])
{
    const Vec2 u = U2 - U1, v = U3 - U1;
    const double a = 0.5 * cross(u, v);
    const double L = 1.0 + A2 / a / a;
    const double R = (dot(v, v) * e0 + dot(u, u) * e1) / 4 / A2 - dot(v, u) * d / 2 / A2;
    const double ar = A2 / a / a / a;                       // dL/dU_k = ar * perp(opposite edge of k)
    const Vec2 n1 = perp(U3 - U2), n2 = perp(U1 - U3), n3 = perp(U2 - U1);
    const Vec2 r1 = mk(((d - e0) * v.x + (d - e1) * u.x) / 2.0 / A2, ((d - e0) * v.y + (d - e1) * u.y) / 2.0 / A2);
    const Vec2 r2 = mk((e1 * u.x - d * v.x) / 2.0 / A2, (e1 * u.y - d * v.y) / 2.0 / A2);
    const Vec2 r3 = mk((e0 * v.x - d * u.x) / 2.0 / A2, (e0 * v.y - d * u.y) / 2.0 / A2);
    // association order kept as w * ((ar*n) * R + r * L): with -fmad=false the per-element
    // contributions are then bit-identical to the CPU reference's
    g[0] = mk(w * ((ar * n1.x) * R + r1.x * L), w * ((ar * n1.y) * R + r1.y * L));
    g[1] = mk(w * ((ar * n2.x) * R + r2.x * L), w * ((ar * n2.y) * R + r2.y * L));
    g[2] = mk(w * ((ar * n3.x) * R + r3.x * L), w * ((ar * n3.y) * R + r3.y * L));
}

// ---- gradient w.r.t. ONE corner k (the vertex-gather assembly visits a triangle once per incident vertex) plus the
// element value.  Same operations in the same order as sd_gradient / sd_energy, so gk is bit-identical to g[k] and E to
// sd_energy's result: the corner's opposite-edge normal and the two coefficients of r_k = (cA v + cB u) / 2 / A^2 are
// selected without branches (k = 1, 2 use the commuted sums  -d v + e1 u  and  e0 v + (-d) u, which round like the
// reference's  e1 u - d v  and  e0 v - d u).
__device__ __forceinline__ void sd_corner(Vec2 U1, Vec2 U2, Vec2 U3, double A2, double e0, double e1, double d, double w,
                                          int k, Vec2& gk, double& E, double& dbArea)
{
    const Vec2 u = U2 - U1, v = U3 - U1;
    dbArea = cross(u, v);
    const double a = 0.5 * dbArea;
    const double L = 1.0 + A2 / a / a;
    const double R = (dot(v, v) * e0 + dot(u, u) * e1) / 4 / A2 - dot(v, u) * d / 2 / A2;
    const double ar = A2 / a / a / a;
    const Vec2 opp = (k == 0) ? (U3 - U2) : ((k == 1) ? (U1 - U3) : (U2 - U1));
    const Vec2 n = perp(opp);
    const double cA = (k == 0) ? (d - e0) : ((k == 1) ? -d : e0);
    const double cB = (k == 0) ? (d - e1) : ((k == 1) ? e1 : -d);
    const Vec2 r = mk((cA * v.x + cB * u.x) / 2.0 / A2, (cA * v.y + cB * u.y) / 2.0 / A2);
    gk = mk(w * ((ar * n.x) * R + r.x * L), w * ((ar * n.y) * R + r.y * L));
    E = w * L * R;
}

// ---- exact 6x6 Hessian, upper block triangle: Hb[b][i][j], b = 0:(1,1) 1:(1,2) 2:(1,3) 3:(2,2)
// 4:(2,3) 5:(3,3); uses the precomputed k0 = |e0|^2/(2A^2), k1 = |e1|^2/(2A^2), kd = e0.e1/(2A^2)
__device__ __forceinline__ void sd_hessian(Vec2 U1, Vec2 U2, Vec2 U3, double A2, double k0, double k1,
                                           double kd, double w, double Hb[6][2][2])
{
    const Vec2 u = U2 - U1, v = U3 - U1;
    const double a = 0.5 * cross(u, v);
    const double ar = A2 / a / a / a;
    const double m = 3.0 / 2.0 * ar / a;                     // d(ar)/da * (-1/2) ... coefficient of n_k n_l^T
    const double L = 1.0 + A2 / a / a;
    const double R = (dot(v, v) * k0 + dot(u, u) * k1) / 2. - dot(v, u) * kd;
    Vec2 n[3], r[3];
    n[0] = perp(U3 - U2); n[1] = perp(U1 - U3); n[2] = perp(U2 - U1);
    r[0] = mk((kd - k0) * v.x + (kd - k1) * u.x, (kd - k0) * v.y + (kd - k1) * u.y);
    r[1] = mk(k1 * u.x - kd * v.x, k1 * u.y - kd * v.y);
    r[2] = mk(k0 * v.x - kd * u.x, k0 * v.y - kd * u.y);
    // second derivative of R: c_kl * I
    const double c[6] = {k0 + k1 - 2.0 * kd, kd - k1, kd - k0, k1, -kd, k0};
    // derivative of perp(opp_k) w.r.t. U_l is s_kl * [[0,1],[-1,0]]
    const double s[6] = {0.0, -1.0, 1.0, 0.0, -1.0, 0.0};
    const int bk[6] = {0, 0, 0, 1, 1, 2}, bl[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const Vec2 nk = n[bk[b]], nl = n[bl[b]], rk = r[bk[b]], rl = r[bl[b]];
        const double q = ar * s[b];
        const double cL = c[b] * L;
        // ((R * d2L + dL_k r_l^T) + c L I) + r_k dL_l^T  — summation order of the reference's expression
        Hb[b][0][0] = w * ((((m * nk.x) * nl.x) * R + (ar * nk.x) * rl.x + cL) + rk.x * (ar * nl.x));
        Hb[b][0][1] = w * ((((m * nk.x) * nl.y + q) * R + (ar * nk.x) * rl.y) + rk.x * (ar * nl.y));
        Hb[b][1][0] = w * ((((m * nk.y) * nl.x - q) * R + (ar * nk.y) * rl.x) + rk.y * (ar * nl.x));
        Hb[b][1][1] = w * ((((m * nk.y) * nl.y) * R + (ar * nk.y) * rl.y + cL) + rk.y * (ar * nl.y));
    }
}

// ---- symmetric 4x4 eigen-decomposition, cyclic Jacobi, everything in registers ----------------
__device__ __forceinline__ void jacobi_rot(double A[4][4], double V[4][4], const int p, const int q)
{
    const double apq = A[p][q];
    if (apq == 0.0) return;
    const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
    const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = rsqrt(t * t + 1.0), s = t * c;
    A[p][p] -= t * apq;
    A[q][q] += t * apq;
    A[p][q] = A[q][p] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r != p && r != q) {
            const double arp = A[r][p], arq = A[r][q];
            A[r][p] = A[p][r] = c * arp - s * arq;
            A[r][q] = A[q][r] = s * arp + c * arq;
        }
        const double vrp = V[r][p], vrq = V[r][q];
        V[r][p] = c * vrp - s * vrq;
        V[r][q] = s * vrp + c * vrq;
    }
}
__device__ __forceinline__ void jacobi4(double A[4][4], double V[4][4])
{
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[0][3] * A[0][3] +
                           A[1][2] * A[1][2] + A[1][3] * A[1][3] + A[2][3] * A[2][3];
        const double dg = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2] + A[3][3] * A[3][3];
        if (off <= 1e-33 * dg || off == 0.0) break;
        jacobi_rot(A, V, 0, 1); jacobi_rot(A, V, 2, 3);
        jacobi_rot(A, V, 0, 2); jacobi_rot(A, V, 1, 3);
        jacobi_rot(A, V, 0, 3); jacobi_rot(A, V, 1, 2);
    }
}

#ifdef OCB_ELEMENT_HOST
static long host_general_path_count = 0;
#endif
// ---- 4x4 helpers of the projection ---------------------------------------------------------------
// positive-definiteness test: LDL^T pivots, early out on the first non-positive one
__device__ __forceinline__ bool pd4(const double M[4][4])
{
    const double d0 = M[0][0];
    if (!(d0 > 0.0)) return false;
    const double l1 = M[0][1] / d0, l2 = M[0][2] / d0, l3 = M[0][3] / d0;
    const double d1 = M[1][1] - l1 * M[0][1];
    if (!(d1 > 0.0)) return false;
    const double m12 = M[1][2] - l2 * M[0][1], m13 = M[1][3] - l3 * M[0][1];
    const double l21 = m12 / d1, l31 = m13 / d1;
    const double d2 = M[2][2] - l2 * M[0][2] - l21 * m12;
    if (!(d2 > 0.0)) return false;
    const double m23 = M[2][3] - l3 * M[0][2] - l31 * m12;
    const double d3 = M[3][3] - l3 * M[0][3] - l31 * m13 - (m23 / d2) * m23;
    return d3 > 0.0;
}

// LDL^T (no pivoting) of the symmetric S = M - shift I with pivots kept away from zero; returns the unit-lower
// factors l[6] = {l10, l20, l30, l21, l31, l32}, the pivots d[4] and their reciprocals
__device__ __forceinline__ void ldl4(const double M[4][4], double shift, double tiny, double l[6], double d[4], double id[4])
{
    const double s00 = M[0][0] - shift, s11 = M[1][1] - shift, s22 = M[2][2] - shift, s33 = M[3][3] - shift;
    d[0] = fabs(s00) < tiny ? tiny : s00; id[0] = 1.0 / d[0];
    l[0] = M[0][1] * id[0]; l[1] = M[0][2] * id[0]; l[2] = M[0][3] * id[0];
    const double a11 = fma(-l[0], M[0][1], s11), a12 = fma(-l[0], M[0][2], M[1][2]), a13 = fma(-l[0], M[0][3], M[1][3]);
    const double a22 = fma(-l[1], M[0][2], s22), a23 = fma(-l[1], M[0][3], M[2][3]), a33 = fma(-l[2], M[0][3], s33);
    d[1] = fabs(a11) < tiny ? tiny : a11; id[1] = 1.0 / d[1];
    l[3] = a12 * id[1]; l[4] = a13 * id[1];
    const double b22 = fma(-l[3], a12, a22), b23 = fma(-l[3], a13, a23), b33 = fma(-l[4], a13, a33);
    d[2] = fabs(b22) < tiny ? tiny : b22; id[2] = 1.0 / d[2];
    l[5] = b23 * id[2];
    const double c33 = fma(-l[5], b23, b33);
    d[3] = fabs(c33) < tiny ? tiny : c33; id[3] = 1.0 / d[3];
}

// y = M x, rho = x.y, res2 = |y - rho x|^2 for a unit vector x
__device__ __forceinline__ void rayleigh4(const double M[4][4], const double x[4], double& rho, double& res2)
{
    double y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = fma(M[i][3], x[3], fma(M[i][2], x[2], fma(M[i][1], x[1], M[i][0] * x[0])));
    rho = fma(x[3], y[3], fma(x[2], y[2], fma(x[1], y[1], x[0] * y[0])));
    res2 = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double r = fma(-rho, x[i], y[i]); res2 = fma(r, r, res2); }
}

// Smallest eigenpair of an indefinite symmetric 4x4 by Rayleigh-quotient iteration.  Start: with k the first
// non-positive LDL^T pivot, x = L^-T e_k has x^T M x = d_k <= 0.  While the residual is still large the shift is
// moved down to rho - |r| (a lower bound of the smallest eigenvalue once x is dominated by its eigenvector), which
// keeps the iteration from locking onto the nearest non-negative eigenvalue; after that the convergence is cubic,
// so the step taken from |r| <= 1e-5 |rho| ends at rounding level.  false: not converged / not negative.
__device__ __forceinline__ bool min_eigpair4(const double M[4][4], double& lam, double x[4])
{
    const double nrm = fabs(M[0][0]) + fabs(M[1][1]) + fabs(M[2][2]) + fabs(M[3][3]) +
                       fabs(M[0][1]) + fabs(M[0][2]) + fabs(M[0][3]) + fabs(M[1][2]) + fabs(M[1][3]) + fabs(M[2][3]);
    const double tiny = 1e-18 * nrm;
    if (!(nrm > 0.0) || !(nrm < 1e300)) return false;
    double l[6], d[4], id[4];
    ldl4(M, 0.0, tiny, l, d, id);
    {
        const int k = !(d[0] > 0.0) ? 0 : (!(d[1] > 0.0) ? 1 : (!(d[2] > 0.0) ? 2 : 3));
        x[3] = (k == 3) ? 1.0 : 0.0;
        x[2] = ((k == 2) ? 1.0 : 0.0) - l[5] * x[3];
        x[1] = ((k == 1) ? 1.0 : 0.0) - l[3] * x[2] - l[4] * x[3];
        x[0] = ((k == 0) ? 1.0 : 0.0) - l[0] * x[1] - l[1] * x[2] - l[2] * x[3];
        const double s = rsqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] *= s;
    }
    double rho, res2;
    rayleigh4(M, x, rho, res2);
    for (int it = 0; it < 10; ++it) {
        const bool last = res2 <= 1e-10 * rho * rho;
        const double shift = (res2 > 1e-2 * rho * rho) ? rho - sqrt(res2) : rho;
        ldl4(M, shift, tiny, l, d, id);
        // (L D L^T) z = x
        double z0 = x[0], z1 = fma(-l[0], z0, x[1]);
        double z2 = fma(-l[3], z1, fma(-l[1], z0, x[2]));
        double z3 = fma(-l[5], z2, fma(-l[4], z1, fma(-l[2], z0, x[3])));
        z0 *= id[0]; z1 *= id[1]; z2 *= id[2]; z3 *= id[3];
        z2 = fma(-l[5], z3, z2);
        z1 = fma(-l[4], z3, fma(-l[3], z2, z1));
        z0 = fma(-l[2], z3, fma(-l[1], z2, fma(-l[0], z1, z0)));
        const double s = rsqrt(fma(z3, z3, fma(z2, z2, fma(z1, z1, z0 * z0))));
        if (!(s > 0.0) || !(s < 1e300)) return false;
        x[0] = z0 * s; x[1] = z1 * s; x[2] = z2 * s; x[3] = z3 * s;
        rayleigh4(M, x, rho, res2);
        if (last) { lam = rho; return rho < 0.0; }
    }
    return false;
}

// Hb -= lam * (W v)(W v)^T on the six upper blocks, W = C (x) I2
__device__ __forceinline__ void psd_rank1(double Hb[6][2][2], const double C[3][2], double lam, const double v[4])
{
    double y[3][2];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        y[k][0] = C[k][0] * v[0] + C[k][1] * v[2];
        y[k][1] = C[k][0] * v[1] + C[k][1] * v[3];
    }
    const int bk[6] = {0, 0, 0, 1, 1, 2}, bl[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int b = 0; b < 6; ++b)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) Hb[b][i][j] -= lam * y[bk[b]][i] * y[bl[b]][j];
}

// ---- projection of the element Hessian onto the PSD cone ---------------------------------------
// makePD clamps the negative eigenvalues of the 6x6 matrix in vertex space (IglUtils.hpp:71-90).
// The exact Hessian annihilates the two UV translations, so with the fixed orthonormal basis
// W = C (x) I2 of their complement (C = [q1 q2], q1 = (1,-1,0)/sqrt2, q2 = (1,1,-2)/sqrt6)
//     H = W M W^T,  M = W^T H W (4x4),   PSD(H) = W PSD(M) W^T      (same eigenpairs)
// and PSD(M) = M + sum_{lambda_i<0} |lambda_i| v_i v_i^T.  PSD blocks are left bit-untouched, like
// the reference's early-out.  Returns the number of clamped eigenvalues.
__device__ __forceinline__ int sd_project_psd(double Hb[6][2][2])
{
    const double s2 = 0.70710678118654752440, s6 = 0.40824829046386301637;
    const double C[3][2] = {{s2, s6}, {-s2, s6}, {0.0, -2.0 * s6}};
    double M[4][4];
    // M_ab(i,j) = sum_kl C[k][a] C[l][b] H_kl(i,j) with H_lk = H_kl^T
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = a; b < 2; ++b)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double acc = 0.0;
                    acc += C[0][a] * C[0][b] * Hb[0][i][j];
                    acc += C[0][a] * C[1][b] * Hb[1][i][j] + C[1][a] * C[0][b] * Hb[1][j][i];
                    acc += C[0][a] * C[2][b] * Hb[2][i][j] + C[2][a] * C[0][b] * Hb[2][j][i];
                    acc += C[1][a] * C[1][b] * Hb[3][i][j];
                    acc += C[1][a] * C[2][b] * Hb[4][i][j] + C[2][a] * C[1][b] * Hb[4][j][i];
                    acc += C[2][a] * C[2][b] * Hb[5][i][j];
                    M[2 * a + i][2 * b + j] = acc;
                }
    // symmetrise (M11, M22 are symmetric up to rounding; M21 = M12^T)
    M[1][0] = M[0][1] = 0.5 * (M[0][1] + M[1][0]);
    M[3][2] = M[2][3] = 0.5 * (M[2][3] + M[3][2]);
    M[2][0] = M[0][2]; M[2][1] = M[1][2]; M[3][0] = M[0][3]; M[3][1] = M[1][3];

    if (pd4(M)) return 0;
    // Indefinite.  These element Hessians almost always have exactly ONE negative eigenvalue (91 % of the triangles at
    // the Tutte start, 67 % near convergence, none with two), so the projection is a rank-1 update with the smallest
    // eigenpair, found by a safeguarded Rayleigh-quotient iteration (~4 shifted 4x4 LDL^T solves) instead of a full
    // Jacobi eigen-solve (~6 sweeps x 6 rotations).  M - 2 lam v v^T must then be positive definite (the one
    // negative eigenvalue reflected); if it is not, or the iteration did not converge, the general path below runs.
    {
        double lam, v[4];
        if (min_eigpair4(M, lam, v)) {
            double R[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) R[i][j] = fma(-2.0 * lam * v[i], v[j], M[i][j]);
            if (pd4(R)) {
                psd_rank1(Hb, C, lam, v);
                return 1;
            }
        }
    }
#ifdef OCB_ELEMENT_HOST
    ++host_general_path_count;          // test harness only: how often the general eigen-solve still runs
#endif
    double V[4][4];
    jacobi4(M, V);
    int clamped = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double lam = M[e][e];
        if (lam < 0.0) {
            ++clamped;
            const double v[4] = {V[0][e], V[1][e], V[2][e], V[3][e]};
            psd_rank1(Hb, C, lam, v);
        }
    }
    return clamped;
}

// ---- largest step keeping the signed area positive (smallest positive root of a t^2 + b t + c) ---
__device__ __forceinline__ double sd_step_bound(Vec2 U1, Vec2 U2, Vec2 U3, Vec2 D1, Vec2 D2, Vec2 D3, double cur)
{
    const Vec2 u = U2 - U1, v = U3 - U1, du = D2 - D1, dv = D3 - D1;
    const double a = cross(du, dv);
    const double b = u.x * dv.y - u.y * dv.x + du.x * v.y - du.y * v.x;
    const double c = cross(u, v);
    const double delta = b * b - 4.0 * a * c;
    double bound = cur;
    if (a > 0.0) {
        if (b < 0.0 && delta >= 0.0) bound = 2.0 * c / (-b + sqrt(delta));
    } else if (a < 0.0) {
        if (b < 0.0) bound = 2.0 * c / (-b + sqrt(delta));
        else bound = (-b - sqrt(delta)) / 2.0 / a;
    } else if (b < 0.0) {
        bound = -c / b;
    }
    return bound < cur ? bound : cur;
}

}  // namespace ocb
