// optcuts_b200 — kernel (6): batched evaluation of candidate topology operations.
//
// The reference scores every candidate vertex split / edge merge by building a tiny local mesh (<= ~20
// triangles, 1-2 free vertices), running a nested dense projected-Newton Optimizer on it (relGL2Tol 1e-6,
// <= maxIter iterations) and taking E_init - E_final (TriMesh::computeLocalLDec -> computeLocalEdDec_* ->
// Optimizer(useDense) : TriMesh.cpp:2105-2794, Optimizer.cpp:203-261, 505-673).  Candidates are independent:
// here ONE CTA runs the whole local Newton solve of one stencil out of shared memory (rest features,
// element Hessians + PSD projection, dense LDL^T, step bound, line search), and a block-reduce picks the
// arg-max score (first maximum, like the reference's serial scan TriMesh.cpp:726-731).
// Non-bijective stencils only (no local air mesh: that needs the host's Triangle call).
// Compiled with -fmad=false like the element kernels.
#include "ocb_internal.cuh"
#include "ocb_element.cuh"

namespace ocb {

static constexpr int kStBlock = 128;
static constexpr int kMaxV = 64, kMaxT = 96, kMaxFree = 16;     // local vertices, triangles, free vertices

struct StencilParams {
    int nStencil;
    const int32_t* vertPtr; const int32_t* triPtr;
    const double* Vrest; const double* UV; const int32_t* F; const uint8_t* isFree;
    const double* scoreScale; const double* scoreOffset;
    int maxIter; double relGL2Tol;
    double* Einit; double* Efinal; double* UVout; int32_t* iters; double* score; int32_t* status;
};

// deterministic block sum / min of one double (kStBlock threads)
template <bool IS_MIN>
__device__ __forceinline__ double block_reduce(double v, double* sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, o); v = IS_MIN ? fmin(v, t) : v + t; }
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double r = sm[0];
#pragma unroll
    for (int w = 1; w < kStBlock / 32; ++w) r = IS_MIN ? fmin(r, sm[w]) : r + sm[w];
    return r;
}

__global__ void __launch_bounds__(kStBlock)
stencil_newton_kernel(StencilParams P)
{
    __shared__ double sU[kMaxV][2], sU0[kMaxV][2], sDir[kMaxV][2];
    __shared__ double sRest[8][kMaxT];
    __shared__ int sF[kMaxT][3];
    __shared__ int sFreeIdx[kMaxV];            // local vertex -> free slot or -1
    __shared__ double sHt[kMaxT][21];          // projected element Hessians, upper triangle of the 6x6
    __shared__ double sGt[kMaxT][6];           // element gradients
    __shared__ double sH[2 * kMaxFree][2 * kMaxFree + 1];
    __shared__ double sG[2 * kMaxFree], sP[2 * kMaxFree];
    __shared__ double sRed[kStBlock / 32];
    __shared__ int sFreeList[kMaxFree];

    const int s = blockIdx.x;
    if (s >= P.nStencil) return;
    const int v0 = P.vertPtr[s], nV = P.vertPtr[s + 1] - v0, t0 = P.triPtr[s], nT = P.triPtr[s + 1] - t0;
    const int tid = threadIdx.x;
    if (nV > kMaxV || nT > kMaxT || nV <= 0 || nT <= 0) {
        if (tid == 0) { P.status[s] = -2; P.Einit[s] = 0.0; P.Efinal[s] = 0.0; P.iters[s] = 0; P.score[s] = -INFINITY; }
        return;
    }
    // ---- load
    for (int v = tid; v < nV; v += kStBlock) { sU[v][0] = P.UV[2 * (size_t)(v0 + v)]; sU[v][1] = P.UV[2 * (size_t)(v0 + v) + 1]; }
    for (int t = tid; t < nT; t += kStBlock) for (int k = 0; k < 3; ++k) sF[t][k] = P.F[3 * (size_t)(t0 + t) + k];
    __syncthreads();
    int nFree = 0;
    if (tid == 0) {
        for (int v = 0; v < nV; ++v) { if (P.isFree[v0 + v] && nFree < kMaxFree) { sFreeIdx[v] = nFree; sFreeList[nFree] = v; ++nFree; } else sFreeIdx[v] = -1; }
        sRed[0] = (double)nFree;
    }
    __syncthreads();
    nFree = (int)sRed[0];
    int nFreeAll = 0;
    for (int v = 0; v < nV; ++v) nFreeAll += P.isFree[v0 + v] ? 1 : 0;
    if (nFreeAll > kMaxFree) {
        if (tid == 0) { P.status[s] = -2; P.Einit[s] = 0.0; P.Efinal[s] = 0.0; P.iters[s] = 0; P.score[s] = -INFINITY; }
        return;
    }
    // ---- rest features of the local mesh (TriMesh::computeFeatures, TriMesh.cpp:343-398), weights w = A_t / sum A
    double myArea = 0.0;
    for (int t = tid; t < nT; t += kStBlock) {
        const double* p0 = P.Vrest + 3 * (size_t)(v0 + sF[t][0]); const double* p1 = P.Vrest + 3 * (size_t)(v0 + sF[t][1]);
        const double* p2 = P.Vrest + 3 * (size_t)(v0 + sF[t][2]);
        const double ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
        const double bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
        const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
        const double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz), A2 = area * area;
        const double e0 = ax * ax + ay * ay + az * az, e1 = bx * bx + by * by + bz * bz, d = ax * bx + ay * by + az * bz;
        sRest[0][t] = area; sRest[1][t] = A2; sRest[2][t] = e0; sRest[3][t] = e1; sRest[4][t] = d;
        sRest[5][t] = e0 / 2. / A2; sRest[6][t] = e1 / 2. / A2; sRest[7][t] = d / 2. / A2;
        myArea += area;
    }
    const double surf = block_reduce<false>(myArea, sRed);

    auto energy_at = [&](bool& inverted) {
        double e = 0.0, inv = 0.0;
        for (int t = tid; t < nT; t += kStBlock) {
            const Vec2 U1 = mk(sU[sF[t][0]][0], sU[sF[t][0]][1]), U2 = mk(sU[sF[t][1]][0], sU[sF[t][1]][1]), U3 = mk(sU[sF[t][2]][0], sU[sF[t][2]][1]);
            double db;
            e += sd_energy(U2 - U1, U3 - U1, sRest[1][t], sRest[2][t], sRest[3][t], sRest[4][t], sRest[0][t] / surf, db);
            if (db < 0.0) inv += 1.0;
        }
        const double tot = block_reduce<false>(e, sRed);
        inverted = block_reduce<false>(inv, sRed) > 0.0;
        return tot;
    };

    bool inv0;
    double E = energy_at(inv0);
    const double E0 = E;
    // Optimizer::updateTargetGRes (Optimizer.cpp:675-678) with energyParamSum = 1
    const double targetGRes = (double)(nV - (nV - nFree)) / (double)nV * P.relGL2Tol;
    int it = 0, status = inv0 ? -4 : 0;
    const int n = 2 * nFree;
    for (; it < P.maxIter && status == 0 && nFree > 0; ++it) {
        // ---- element gradients + projected Hessians
        for (int t = tid; t < nT; t += kStBlock) {
            const Vec2 U1 = mk(sU[sF[t][0]][0], sU[sF[t][0]][1]), U2 = mk(sU[sF[t][1]][0], sU[sF[t][1]][1]), U3 = mk(sU[sF[t][2]][0], sU[sF[t][2]][1]);
            const double w = sRest[0][t] / surf;
            Vec2 g[3];
            sd_gradient(U1, U2, U3, sRest[1][t], sRest[2][t], sRest[3][t], sRest[4][t], w, g);
            for (int k = 0; k < 3; ++k) { sGt[t][2 * k] = g[k].x; sGt[t][2 * k + 1] = g[k].y; }
            double Hb[6][2][2];
            sd_hessian(U1, U2, U3, sRest[1][t], sRest[5][t], sRest[6][t], sRest[7][t], w, Hb);
            sd_project_psd(Hb);
            // upper triangle of the 6x6, row-major: (r, c >= r)
            const int bOf[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
            int q = 0;
            for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) {
                const int k = r >> 1, l = c >> 1, i = r & 1, j = c & 1;
                sHt[t][q++] = Hb[bOf[k][l]][i][j];
            }
        }
        __syncthreads();
        // ---- gather into the dense free-DOF system, fixed order (deterministic)
        for (int e = tid; e < n * n; e += kStBlock) {
            const int r = e / n, c = e % n;
            if (c < r) continue;
            const int vr = sFreeList[r >> 1], vc = sFreeList[c >> 1];
            double acc = 0.0;
            for (int t = 0; t < nT; ++t) {
                int kr = -1, kc = -1;
                for (int k = 0; k < 3; ++k) { if (sF[t][k] == vr) kr = k; if (sF[t][k] == vc) kc = k; }
                if (kr < 0 || kc < 0) continue;
                int rr = 2 * kr + (r & 1), cc = 2 * kc + (c & 1);
                if (rr > cc) { const int tmp = rr; rr = cc; cc = tmp; }
                acc += sHt[t][rr * 6 - rr * (rr - 1) / 2 + (cc - rr)];
            }
            sH[r][c] = acc; sH[c][r] = acc;
        }
        for (int r = tid; r < n; r += kStBlock) {
            const int vr = sFreeList[r >> 1];
            double acc = 0.0;
            for (int t = 0; t < nT; ++t) for (int k = 0; k < 3; ++k) if (sF[t][k] == vr) acc += sGt[t][2 * k + (r & 1)];
            sG[r] = acc;
        }
        __syncthreads();
        double sq = 0.0;
        for (int r = 0; r < n; ++r) sq += sG[r] * sG[r];
        if (sq < targetGRes) { status = 1; ++it; break; }     // converged; globalIterNum counts this pass too (Optimizer.cpp:215-221)
        // ---- dense LDL^T solve H p = -g (one warp; the reference uses Eigen::LDLT on the same matrix)
        if (tid < 32) {
            for (int k = 0; k < n; ++k) {
                const double dk = sH[k][k];
                __syncwarp();
                for (int i = k + 1 + tid; i < n; i += 32) sH[i][k] = sH[i][k] / dk;       // L(i,k)
                __syncwarp();
                for (int i = k + 1 + tid; i < n; i += 32) {
                    const double lik = sH[i][k];
                    for (int j = k + 1; j <= i; ++j) sH[i][j] -= lik * dk * sH[j][k];
                }
                __syncwarp();
            }
            if (tid == 0) {
                for (int i = 0; i < n; ++i) { double v = -sG[i]; for (int j = 0; j < i; ++j) v -= sH[i][j] * sP[j]; sP[i] = v; }
                for (int i = 0; i < n; ++i) sP[i] /= sH[i][i];
                for (int i = n - 1; i >= 0; --i) { double v = sP[i]; for (int j = i + 1; j < n; ++j) v -= sH[j][i] * sP[j]; sP[i] = v; }
            }
        }
        __syncthreads();
        for (int v = tid; v < nV; v += kStBlock) {
            const int f = sFreeIdx[v];
            sDir[v][0] = f >= 0 ? sP[2 * f] : 0.0; sDir[v][1] = f >= 0 ? sP[2 * f + 1] : 0.0;
            sU0[v][0] = sU[v][0]; sU0[v][1] = sU[v][1];
        }
        __syncthreads();
        // ---- step bound (SymDirichletEnergy::initStepSize) and line search (Optimizer::lineSearch, :575-652)
        double bound = 1.0;
        for (int t = tid; t < nT; t += kStBlock) {
            const int a = sF[t][0], b = sF[t][1], c = sF[t][2];
            bound = sd_step_bound(mk(sU[a][0], sU[a][1]), mk(sU[b][0], sU[b][1]), mk(sU[c][0], sU[c][1]),
                                  mk(sDir[a][0], sDir[a][1]), mk(sDir[b][0], sDir[b][1]), mk(sDir[c][0], sDir[c][1]), bound);
        }
        double alpha = block_reduce<true>(bound, sRed) * 0.99;
        const double Elast = E;
        bool inverted = false;
        double Etry = 0.0;
        auto step_to = [&](double a) {
            __syncthreads();
            for (int v = tid; v < nV; v += kStBlock) {
                sU[v][0] = __dadd_rn(sU0[v][0], __dmul_rn(a, sDir[v][0]));
                sU[v][1] = __dadd_rn(sU0[v][1], __dmul_rn(a, sDir[v][1]));
            }
            __syncthreads();
            Etry = energy_at(inverted);
        };
        step_to(alpha);
        while (Etry > Elast && alpha > 0.0) { alpha /= 2.0; step_to(alpha); }          // plain decrease test (:597-610)
        while (inverted && alpha > 0.0) { alpha /= 2.0; step_to(alpha); }              // inversion guard (:615-629)
        E = Etry;
        const double eDec = Elast - E;
        if (alpha == 0.0) { status = 1; ++it; break; }
        if ((eDec / Elast < 1.0e-6 * alpha) && (alpha > 1.0e-3)) { status = 1; ++it; break; }     // allowEDecRelTol stop (:635)
    }
    __syncthreads();
    for (int v = tid; v < nV; v += kStBlock) { P.UVout[2 * (size_t)(v0 + v)] = sU[v][0]; P.UVout[2 * (size_t)(v0 + v) + 1] = sU[v][1]; }
    if (tid == 0) {
        P.Einit[s] = E0; P.Efinal[s] = E; P.iters[s] = it; P.status[s] = status < 0 ? status : 0;
        const double sc = P.scoreScale ? P.scoreScale[s] : 1.0, off = P.scoreOffset ? P.scoreOffset[s] : 0.0;
        P.score[s] = status < 0 ? -INFINITY : sc * (E0 - E) + off;
    }
}

// ---------------------------------------------------------------------------------------------
// Bijective stencils: ONE Newton iteration per launch.
//
// With bijectivity on, the nested Optimizer of a candidate re-triangulates its local air mesh after EVERY Newton iteration
// (Optimizer.cpp:252-257 -> Scaffold.cpp:153-199 -> Triangle "qYQ", Steiner points included), and Triangle stays with the
// host.  So the candidates advance in lock step: the host triangulates the air region of every active candidate
// (shim/CudaCandidates.cpp, all host threads), packs mesh + air mesh of all of them into one ragged batch, and this
// kernel runs one iteration of Optimizer::solve's loop body for each (one CTA per candidate, everything in shared memory):
//   air-mesh rest features from the current positions (TriMesh.cpp:355-398 incl. the degeneracy clamp),
//   gradient of mesh term + w_scaf/|Fa| x uniform air term, convergence test ||g||^2 < targetGRes (Optimizer.cpp:209-221),
//   projected element Hessians -> dense free-DOF system -> LDL^T -> search direction (:505-573, dense flavour),
//   step bound over mesh and air triangles x 0.99, line search with both inversion guards, lastEDec without the scaffold's
//   change, the relative-decrease stop (:575-652).
// Free DOFs: the candidate's 1-4 free mesh vertices plus the Steiner points of its air mesh.
static constexpr int kSV = 128, kST = 192, kSFree = 32;        // combined vertices, triangles, free vertices per stencil

struct StepParams {
    int nStencil;
    const int32_t* vertPtr; const int32_t* triPtr; const int32_t* nVm; const int32_t* nTm;
    const double* Vrest; const double* UV; const int32_t* F; const uint8_t* isFree;
    const double* areaThres; const double* targetGRes; double wScaf;
    double* UVout; double* out6; int32_t* result;
};

struct StepSmem {
    double U[kSV][2], U0[kSV][2], Dir[kSV][2];
    double Rest[8][kST];
    double Ht[kST][21];
    double Gt[kST][6];
    double H[2 * kSFree][2 * kSFree + 1];
    double G[2 * kSFree], P[2 * kSFree];
    double Red[kStBlock / 32];
    int F[kST][3];
    int FreeIdx[kSV];
    int FreeList[kSFree];
    int nFree;
};

__global__ void __launch_bounds__(kStBlock)
stencil_step_kernel(StepParams P)
{
    extern __shared__ __align__(16) unsigned char stepRaw[];
    StepSmem& S = *reinterpret_cast<StepSmem*>(stepRaw);
    const int s = blockIdx.x, tid = threadIdx.x;
    if (s >= P.nStencil) return;
    const int v0 = P.vertPtr[s], nV = P.vertPtr[s + 1] - v0, t0 = P.triPtr[s], nT = P.triPtr[s + 1] - t0;
    const int nTm = P.nTm[s], nTa = nT - nTm;
    double* out = P.out6 + 6 * (size_t)s;
    auto fail = [&](int code) { if (tid == 0) { P.result[s] = code; for (int k = 0; k < 6; ++k) out[k] = 0.0; } };
    if (nV > kSV || nT > kST || nV <= 0 || nTm <= 0 || nTa < 0) { fail(-2); return; }
    for (int v = tid; v < nV; v += kStBlock) { S.U[v][0] = P.UV[2 * (size_t)(v0 + v)]; S.U[v][1] = P.UV[2 * (size_t)(v0 + v) + 1]; }
    for (int t = tid; t < nT; t += kStBlock) for (int k = 0; k < 3; ++k) S.F[t][k] = P.F[3 * (size_t)(t0 + t) + k];
    if (tid == 0) {
        int nf = 0, all = 0;
        for (int v = 0; v < nV; ++v) {
            const bool fr = P.isFree[v0 + v] != 0;
            all += fr ? 1 : 0;
            if (fr && nf < kSFree) { S.FreeIdx[v] = nf; S.FreeList[nf] = v; ++nf; } else S.FreeIdx[v] = -1;
        }
        S.nFree = all > kSFree ? -1 : nf;
    }
    __syncthreads();
    const int nFree = S.nFree;
    if (nFree < 0) { fail(-2); return; }
    bool badIndex = false;
    for (int t = tid; t < nT; t += kStBlock) for (int k = 0; k < 3; ++k) if (S.F[t][k] < 0 || S.F[t][k] >= nV) badIndex = true;
    if (__syncthreads_or(badIndex ? 1 : 0)) { fail(-2); return; }

    // ---- rest features: the mesh's from V_rest, the air mesh's from its current positions (z = 0) with the clamp
    double myArea = 0.0;
    const double thres = P.areaThres[s];
    for (int t = tid; t < nT; t += kStBlock) {
        double f8[8];
        const int a = S.F[t][0], b = S.F[t][1], c = S.F[t][2];
        if (t < nTm) {
            const double* p0 = P.Vrest + 3 * (size_t)(v0 + a); const double* p1 = P.Vrest + 3 * (size_t)(v0 + b); const double* p2 = P.Vrest + 3 * (size_t)(v0 + c);
            myArea += sd_rest_features(p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2], p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2], 0.0, f8);
        } else {
            sd_rest_features(S.U[b][0] - S.U[a][0], S.U[b][1] - S.U[a][1], 0.0, S.U[c][0] - S.U[a][0], S.U[c][1] - S.U[a][1], 0.0, thres, f8);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) S.Rest[q][t] = f8[q];
    }
    const double surf = block_reduce<false>(myArea, S.Red);
    const double wS = nTa > 0 ? P.wScaf / nTa : 0.0;

    // energy of mesh (area weights) and air mesh (uniform weights) at the positions in S.U
    auto energy_at = [&](double& Esd, double& Eair, bool& inverted) {
        double em = 0.0, ea = 0.0, inv = 0.0;
        for (int t = tid; t < nT; t += kStBlock) {
            const int a = S.F[t][0], b = S.F[t][1], c = S.F[t][2];
            const Vec2 U1 = mk(S.U[a][0], S.U[a][1]), U2 = mk(S.U[b][0], S.U[b][1]), U3 = mk(S.U[c][0], S.U[c][1]);
            double db;
            const double e = sd_energy(U2 - U1, U3 - U1, S.Rest[1][t], S.Rest[2][t], S.Rest[3][t], S.Rest[4][t], t < nTm ? S.Rest[0][t] / surf : 1.0, db);
            if (t < nTm) em += e; else ea += e;
            if (db < 0.0) inv += 1.0;
        }
        Esd = block_reduce<false>(em, S.Red);
        Eair = block_reduce<false>(ea, S.Red);
        inverted = block_reduce<false>(inv, S.Red) > 0.0;
    };
    double Esd, Eair; bool inv0;
    energy_at(Esd, Eair, inv0);
    if (inv0) { fail(-4); return; }
    const double lastScaf = wS * Eair, Elast = Esd + lastScaf;
    const int n = 2 * nFree;

    // ---- element gradients + projected Hessians (the air mesh's scaled by w_scaf/|Fa| after the projection)
    for (int t = tid; t < nT; t += kStBlock) {
        const int a = S.F[t][0], b = S.F[t][1], c = S.F[t][2];
        const Vec2 U1 = mk(S.U[a][0], S.U[a][1]), U2 = mk(S.U[b][0], S.U[b][1]), U3 = mk(S.U[c][0], S.U[c][1]);
        const double w = t < nTm ? S.Rest[0][t] / surf : 1.0, sc = t < nTm ? 1.0 : wS;
        Vec2 g[3];
        sd_gradient(U1, U2, U3, S.Rest[1][t], S.Rest[2][t], S.Rest[3][t], S.Rest[4][t], w, g);
        for (int k = 0; k < 3; ++k) { S.Gt[t][2 * k] = g[k].x; S.Gt[t][2 * k + 1] = g[k].y; }
        double Hb[6][2][2];
        sd_hessian(U1, U2, U3, S.Rest[1][t], S.Rest[5][t], S.Rest[6][t], S.Rest[7][t], w, Hb);
        sd_project_psd(Hb);
        const int bOf[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        int q = 0;
        for (int r = 0; r < 6; ++r) for (int cc = r; cc < 6; ++cc) S.Ht[t][q++] = sc * Hb[bOf[r >> 1][cc >> 1]][r & 1][cc & 1];
    }
    __syncthreads();
    // gradient of the free DOFs: 1.0 * (mesh sum, triangle order) + w_scaf/|Fa| * (air sum, triangle order)
    for (int r = tid; r < n; r += kStBlock) {
        const int vr = S.FreeList[r >> 1];
        double gm = 0.0, ga = 0.0;
        for (int t = 0; t < nT; ++t) for (int k = 0; k < 3; ++k) if (S.F[t][k] == vr) { if (t < nTm) gm += S.Gt[t][2 * k + (r & 1)]; else ga += S.Gt[t][2 * k + (r & 1)]; }
        S.G[r] = nTa > 0 ? gm + wS * ga : gm;
    }
    __syncthreads();
    double sq = 0.0;
    for (int r = 0; r < n; ++r) sq += S.G[r] * S.G[r];
    if (sq < P.targetGRes[s] || nFree == 0) {            // converged: no step (Optimizer.cpp:215-221)
        for (int v = tid; v < nV; v += kStBlock) { P.UVout[2 * (size_t)(v0 + v)] = S.U[v][0]; P.UVout[2 * (size_t)(v0 + v) + 1] = S.U[v][1]; }
        if (tid == 0) { P.result[s] = 1; out[0] = Esd; out[1] = Elast; out[2] = sq; out[3] = 0.0; out[4] = Elast; out[5] = 0.0; }
        return;
    }
    // ---- dense free-DOF system, contributions in triangle order
    for (int e = tid; e < n * n; e += kStBlock) {
        const int r = e / n, c = e % n;
        if (c < r) continue;
        const int vr = S.FreeList[r >> 1], vc = S.FreeList[c >> 1];
        double acc = 0.0;
        for (int t = 0; t < nT; ++t) {
            int kr = -1, kc = -1;
            for (int k = 0; k < 3; ++k) { if (S.F[t][k] == vr) kr = k; if (S.F[t][k] == vc) kc = k; }
            if (kr < 0 || kc < 0) continue;
            int rr = 2 * kr + (r & 1), cc = 2 * kc + (c & 1);
            if (rr > cc) { const int tmp = rr; rr = cc; cc = tmp; }
            acc += S.Ht[t][rr * 6 - rr * (rr - 1) / 2 + (cc - rr)];
        }
        S.H[r][c] = acc; S.H[c][r] = acc;
    }
    __syncthreads();
    // ---- LDL^T (right-looking, the trailing update spread over the whole CTA) and the two triangular solves (one warp,
    // column sweeps): H p = -g
    for (int k = 0; k < n; ++k) {
        __syncthreads();                                   // the trailing update of step k - 1 is complete
        const double dk = S.H[k][k];
        // a vanished pivot (the projected element Hessians are only semi-definite) is skipped like Eigen::LDLT::solve does:
        // zero column, zero component of the solution
        const bool zeroPivot = !(fabs(dk) > 0.0);
        for (int i = k + 1 + tid; i < n; i += kStBlock) S.H[i][k] = zeroPivot ? 0.0 : S.H[i][k] / dk;              // L(i,k)
        __syncthreads();
        const int m = n - k - 1;
        for (int e = tid; e < m * m; e += kStBlock) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) S.H[i][j] -= S.H[i][k] * dk * S.H[j][k];
        }
    }
    __syncthreads();
    if (tid < 32) {
        for (int i = tid; i < n; i += 32) S.P[i] = -S.G[i];
        __syncwarp();
        for (int j = 0; j < n; ++j) {                      // forward: L y = -g
            const double yj = S.P[j];
            for (int i = j + 1 + tid; i < n; i += 32) S.P[i] -= S.H[i][j] * yj;
            __syncwarp();
        }
        for (int i = tid; i < n; i += 32) S.P[i] = (fabs(S.H[i][i]) > 0.0) ? S.P[i] / S.H[i][i] : 0.0;
        __syncwarp();
        for (int j = n - 1; j >= 0; --j) {                 // backward: L^T p = D^-1 y
            const double pj = S.P[j];
            for (int i = tid; i < j; i += 32) S.P[i] -= S.H[j][i] * pj;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int v = tid; v < nV; v += kStBlock) {
        const int f = S.FreeIdx[v];
        S.Dir[v][0] = f >= 0 ? S.P[2 * f] : 0.0; S.Dir[v][1] = f >= 0 ? S.P[2 * f + 1] : 0.0;
        S.U0[v][0] = S.U[v][0]; S.U0[v][1] = S.U[v][1];
    }
    __syncthreads();
    // ---- step bound over mesh and air triangles, line search
    double bound = 1.0;
    for (int t = tid; t < nT; t += kStBlock) {
        const int a = S.F[t][0], b = S.F[t][1], c = S.F[t][2];
        bound = sd_step_bound(mk(S.U[a][0], S.U[a][1]), mk(S.U[b][0], S.U[b][1]), mk(S.U[c][0], S.U[c][1]),
                              mk(S.Dir[a][0], S.Dir[a][1]), mk(S.Dir[b][0], S.Dir[b][1]), mk(S.Dir[c][0], S.Dir[c][1]), bound);
    }
    double alpha = block_reduce<true>(bound, S.Red) * 0.99;
    bool inverted = false;
    double EsdT = 0.0, EairT = 0.0, Etry = 0.0;
    auto step_to = [&](double a) {
        __syncthreads();
        for (int v = tid; v < nV; v += kStBlock) {
            S.U[v][0] = __dadd_rn(S.U0[v][0], __dmul_rn(a, S.Dir[v][0]));
            S.U[v][1] = __dadd_rn(S.U0[v][1], __dmul_rn(a, S.Dir[v][1]));
        }
        __syncthreads();
        energy_at(EsdT, EairT, inverted);
        Etry = EsdT + wS * EairT;
    };
    step_to(alpha);
    while (Etry > Elast && alpha > 0.0) { alpha /= 2.0; step_to(alpha); }          // plain decrease test (:597-610)
    while (inverted && alpha > 0.0) { alpha /= 2.0; step_to(alpha); }              // inversion guards, mesh and air mesh (:615-629)
    double eDec = Elast - Etry;
    if (nTa > 0) eDec += (-lastScaf + wS * EairT);                                  // the scaffold's own change does not count (:631-634)
    const bool stopped = (alpha == 0.0) || ((eDec / Elast < 1.0e-6 * alpha) && (alpha > 1.0e-3));
    __syncthreads();
    bool bad = false;
    for (int v = tid; v < nV; v += kStBlock) if (!isfinite(S.U[v][0]) || !isfinite(S.U[v][1])) bad = true;
    if (__syncthreads_or(bad ? 1 : 0)) { fail(-5); return; }       // non-finite result: the caller evaluates this stencil itself
    for (int v = tid; v < nV; v += kStBlock) { P.UVout[2 * (size_t)(v0 + v)] = S.U[v][0]; P.UVout[2 * (size_t)(v0 + v) + 1] = S.U[v][1]; }
    if (tid == 0) { P.result[s] = stopped ? 2 : 0; out[0] = EsdT; out[1] = Etry; out[2] = sq; out[3] = alpha; out[4] = Elast; out[5] = eDec; }
}

int launch_stencil_step(ocb_ctx* c, const StencilStepHost& h)
{
    ProfScope prof(c, K_STENCILS);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(stencil_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem));
        if (e != cudaSuccess) return cuda_fail(c, e, "stencil_step_kernel attribute");
        attr = true;
    }
    StepParams P;
    P.nStencil = h.nStencil; P.vertPtr = h.vertPtr; P.triPtr = h.triPtr; P.nVm = h.nVm; P.nTm = h.nTm; P.Vrest = h.Vrest; P.UV = h.UV; P.F = h.F;
    P.isFree = h.isFree; P.areaThres = h.areaThres; P.targetGRes = h.targetGRes; P.wScaf = h.wScaf; P.UVout = h.UVout; P.out6 = h.out6; P.result = h.result;
    stencil_step_kernel<<<h.nStencil, kStBlock, sizeof(StepSmem), c->stream>>>(P);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "stencil_step_kernel");
    return 0;
}

// block-reduce arg-max over the scores: first maximum wins (TriMesh.cpp:726-731)
__global__ void __launch_bounds__(256)
argmax_kernel(int n, const double* __restrict__ score, int* __restrict__ out)
{
    __shared__ double sv[256]; __shared__ int si[256];
    double best = -INFINITY; int bi = -1;
    for (int i = threadIdx.x; i < n; i += 256) { const double v = score[i]; if (v > best) { best = v; bi = i; } }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = sv[threadIdx.x + o]; const int j = si[threadIdx.x + o];
            if (j >= 0 && (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && (si[threadIdx.x] < 0 || j < si[threadIdx.x])))) { sv[threadIdx.x] = v; si[threadIdx.x] = j; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = si[0];
}

int launch_stencils(ocb_ctx* c, const StencilHost& h)
{
    ProfScope prof(c, K_STENCILS);
    StencilParams P;
    P.nStencil = h.nStencil; P.vertPtr = h.vertPtr; P.triPtr = h.triPtr; P.Vrest = h.Vrest; P.UV = h.UV; P.F = h.F; P.isFree = h.isFree;
    P.scoreScale = h.scoreScale; P.scoreOffset = h.scoreOffset; P.maxIter = h.maxIter; P.relGL2Tol = h.relGL2Tol;
    P.Einit = h.Einit; P.Efinal = h.Efinal; P.UVout = h.UVout; P.iters = h.iters; P.score = h.score; P.status = h.status;
    stencil_newton_kernel<<<h.nStencil, kStBlock, 0, c->stream>>>(P);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "stencil_newton_kernel");
    argmax_kernel<<<1, 256, 0, c->stream>>>(h.nStencil, h.score, h.argmax);
    c->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "argmax_kernel");
    return 0;
}

}  // namespace ocb
