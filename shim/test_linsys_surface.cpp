// TEST ONLY: the whole virtual surface of OptCuts::LinSysSolver (LinSysSolver.hpp:22-256) on CudaLinSysSolver, checked against
// Eigen::SimplicialLDLT (what EigenLibSolver wraps, EigenLibSolver.cpp:71-107) on a grid Laplacian with 2x2 blocks:
// set_pattern(vNeighbor, fixedVert) + update_a + factorize + solve, set_pattern(SparseMatrix), multiply, coeffMtr,
// getNumNonzeros.  Prints "linsys surface: ok" and exits 0; tests/test_gpu_dropin.py runs it on the GPU box.
#include "CudaLinSysSolver.hpp"
#include <Eigen/Eigen>
#include <cmath>
#include <cstdio>
#include <set>

using namespace OptCuts;
typedef CudaLinSysSolver<Eigen::VectorXi, Eigen::VectorXd> Solver;

static double relDiff(const Eigen::VectorXd& a, const Eigen::VectorXd& b) { return (a - b).norm() / b.norm(); }

int main(void)
{
    const int m = 24, nV = m * m, n = 2 * nV;
    // SPD block matrix: for every grid edge (u, v) a PSD 4x4 element [[B, -B], [-B, B]], B = [[2, .5], [.5, 1]], plus 1e-2 I
    std::vector<std::set<int>> nb(nV);
    std::vector<Eigen::Triplet<double>> trip;
    std::vector<int> I, J; std::vector<double> S;
    auto add = [&](int r, int c, double v) { trip.emplace_back(r, c, v); if (r <= c) { I.push_back(r); J.push_back(c); S.push_back(v); } };
    const double B[2][2] = {{2.0, 0.5}, {0.5, 1.0}};
    auto edge = [&](int u, int v) {
        nb[u].insert(v); nb[v].insert(u);
        for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
            add(2 * u + a, 2 * u + b, B[a][b]); add(2 * v + a, 2 * v + b, B[a][b]);
            add(2 * u + a, 2 * v + b, -B[a][b]); add(2 * v + a, 2 * u + b, -B[a][b]);
        }
    };
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) { if (i + 1 < m) edge(i * m + j, (i + 1) * m + j); if (j + 1 < m) edge(i * m + j, i * m + j + 1); }
    for (int k = 0; k < n; ++k) add(k, k, 1.0e-2);
    Eigen::SparseMatrix<double> A(n, n);
    A.setFromTriplets(trip.begin(), trip.end());
    Eigen::VectorXd rhs(n);
    for (int k = 0; k < n; ++k) rhs[k] = std::sin(0.37 * k) + 0.1 * std::cos(1.3 * k);
    Eigen::SimplicialLDLT<Eigen::SparseMatrix<double>> ldlt(A);
    const Eigen::VectorXd want = ldlt.solve(rhs);
    std::vector<double> xy(2 * nV);                       // geometry hint: the grid itself
    for (int v = 0; v < nV; ++v) { xy[v] = v / m; xy[nV + v] = v % m; }
    int nHint = nV;
    cudaCoordinateHint(true, nHint, xy);

    int bad = 0;
    {   // route 1: the Optimizer's call sequence (Optimizer.cpp:173-183)
        Solver s;
        s.set_type(1, 2);
        s.set_pattern(nb, std::set<int>());
        Eigen::VectorXi II = Eigen::Map<Eigen::VectorXi>(I.data(), I.size()), JJ = Eigen::Map<Eigen::VectorXi>(J.data(), J.size());
        Eigen::VectorXd SS = Eigen::Map<Eigen::VectorXd>(S.data(), S.size());
        s.update_a(II, JJ, SS);
        s.analyze_pattern();
        if (!s.factorize()) { std::printf("factorize failed\n"); return 1; }
        Eigen::VectorXd x, Ax;
        s.solve(rhs, x);
        const double e1 = relDiff(x, want);
        s.multiply(want, Ax);
        const double e2 = relDiff(Ax, rhs);
        double e3 = 0.0;
        for (int k = 0; k < 50; ++k) { const int r = (37 * k) % n, c = (r + (k % 3)) % n; e3 = std::max(e3, std::abs(s.coeffMtr(r, c) - A.coeff(std::min(r, c), std::max(r, c)))); }
        const int nnzUpper = (int)Eigen::SparseMatrix<double>(A.triangularView<Eigen::Upper>()).nonZeros();
        std::printf("route vNeighbor + update_a: solve %.2e (%d CG iterations), multiply %.2e, coeffMtr %.2e, nnz %d (upper triangle %d)\n",
                    e1, s.lastIterations(), e2, e3, s.getNumNonzeros(), nnzUpper);
        bad += !(e1 < 1e-9) + !(e2 < 1e-12) + !(e3 < 1e-12) + (s.getNumNonzeros() != nnzUpper);
    }
    {   // route 2: EigenLibSolver::set_pattern(const SparseMatrix&) (EigenLibSolver.cpp:46-69): pattern and values in one call
        Solver s;
        s.set_pattern(A);
        if (!s.factorize()) { std::printf("factorize failed\n"); return 1; }
        Eigen::VectorXd x;
        s.solve(rhs, x);
        const double e1 = relDiff(x, want);
        std::printf("route set_pattern(SparseMatrix): solve %.2e (%d CG iterations)\n", e1, s.lastIterations());
        bad += !(e1 < 1e-9);
    }
    std::printf(bad ? "linsys surface: FAILED\n" : "linsys surface: ok\n");
    return bad ? 1 : 0;
}
