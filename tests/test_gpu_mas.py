"""GPU: the multilevel additive Schwarz preconditioned CG against the reference's search direction, and the
iteration counts the CPU prototype (tools/mas_proto.py --lib, same hierarchy) predicts: 174 (state 1) / 107 (state 100)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _newton_system(ctx, state):
    state.upload(ctx)
    ctx.gradient(state.p0, download=False)
    ctx.set_pattern_from_elements()
    ctx.hessian_assemble(state.p0)
    ctx.factorize()


def test_hierarchy_is_active(ctx, state1):
    _newton_system(ctx, state1)
    info = ctx.precond_info()
    assert info["enabled"] and info["levels"] >= 2 and info["nodes"][0] >= state1.nV // 8
    assert 6 * info["nodes"][-1] <= 3072                 # the coarse level, inverted exactly


@pytest.mark.parametrize("which,bound", [("state1", 260), ("state100", 170)])
def test_mas_pcg_matches_reference_direction(ctx, request, which, bound):
    state = request.getfixturevalue(which)
    _newton_system(ctx, state)
    p, info = ctx.solve(None, 1e-12, 0)
    ref = state.r("searchDir")
    assert np.linalg.norm(p - ref) / np.linalg.norm(ref) < 1e-8
    assert info["iters"] < bound, info                       # block-Jacobi needs 2466 / 1076
    # true residual through the stand-alone SpMV
    g, _ = ctx.gradient(state.p0)
    res = ctx.multiply(p) + g
    assert np.linalg.norm(res) / np.linalg.norm(g) < 1e-10


def test_row_order_from_the_host_uv_mirror_equals_the_downloaded_one(state1, monkeypatch):
    """ocb_set_pattern_* orders the rows along a curve through the UVs; it takes them from the host mirror ocb_set_uv
    keeps (no download + sync) when x has not moved on the device since: same order, hence the same iteration counts and
    the same solution (assembly, preconditioner set-up and PCG are all free of atomics: tests/test_gpu_reproducible.py)."""
    import optcuts_b200 as ob
    got = []
    for no_mirror in ("0", "1"):
        monkeypatch.setenv("OCB_NO_UV_MIRROR", no_mirror)
        c = ob.Context(0)
        try:
            _newton_system(c, state1)
            p, info = c.solve(None, 1e-12, 0)
            got.append((p.copy(), info["iters"]))
            # after a device-side step the mirror is stale: the next pattern must come from a download again
            r = c.newton_step(state1.p0, 0.0)
            c.set_pattern_from_elements()
            c.gradient(state1.p0, download=False)
            c.hessian_assemble(state1.p0)
            p2, info2 = c.solve(None, 1e-12, 0)
            got[-1] += (p2.copy(), info2["iters"], r["E_new"])
        finally:
            c.close()
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert abs(got[0][1] - got[1][1]) <= 2 and rel(got[0][0], got[1][0]) < 1e-9
    assert abs(got[0][3] - got[1][3]) <= 2 and rel(got[0][2], got[1][2]) < 1e-9 and abs(got[0][4] - got[1][4]) <= 1e-11 * got[1][4]


def test_solve_with_explicit_rhs_and_fixed_rows(ctx, state1):
    """LinSysSolver::solve(rhs, result) with an arbitrary right-hand side, also on the fixed vertex's rows."""
    _newton_system(ctx, state1)
    rng = np.random.default_rng(3)
    n = ctx.sizes()["nSys"]
    b = rng.standard_normal(n)
    x, info = ctx.solve(b, 1e-12, 0)
    assert np.linalg.norm(ctx.multiply(x) - b) / np.linalg.norm(b) < 1e-10


def test_pattern_without_uv_falls_back_to_block_jacobi(ctx, state1):
    """A bare LinSysSolver (set_pattern + update_a, no mesh on the device) has no coordinates: the hierarchy
    is skipped and the solve still converges."""
    import optcuts_b200 as ob
    from oracle import portapi
    c2 = ob.Context(0)
    try:
        n = 40
        idx = np.arange(n * n).reshape(n, n)
        adj = [set() for _ in range(n * n)]
        for i in range(n):
            for j in range(n):
                for di, dj in ((0, 1), (1, 0), (1, 1)):
                    if i + di < n and j + dj < n:
                        a, b = idx[i, j], idx[i + di, j + dj]
                        adj[a].add(b); adj[b].add(a)
        ptr = np.zeros(n * n + 1, np.int32)
        ptr[1:] = np.cumsum([len(s) for s in adj])
        ind = np.concatenate([sorted(s) for s in adj]).astype(np.int32)
        c2.set_pattern(ptr, ind, [0])
        assert not c2.precond_info()["enabled"]
        # graph Laplacian + identity on every 2x2 block diagonal
        I, J, S = [], [], []
        for a in range(n * n):
            for k in range(2):
                I.append(2 * a + k); J.append(2 * a + k); S.append(len(adj[a]) + 1.0)
            for b in adj[a]:
                if b > a and a != 0:
                    for k in range(2):
                        I.append(2 * a + k); J.append(2 * b + k); S.append(-1.0)
        c2.update_values_triplets(np.array(I, np.int32), np.array(J, np.int32), np.array(S))
        rhs = np.random.default_rng(0).standard_normal(2 * n * n)
        rhs[:2] = 0.0
        x, info = c2.solve(rhs, 1e-12, 0)
        assert np.linalg.norm(c2.multiply(x) - rhs) / np.linalg.norm(rhs) < 1e-10
    finally:
        c2.close()


def test_bare_solver_with_coordinate_hint(ctx, state1):
    """The drop-in route (shim/CudaLinSysSolver): a solver-only context fed with the reference's adjacency and triplets
    gets its geometry through ocb_set_coordinate_hint (mesh UVs only; the air mesh's interior vertices are placed by
    the library) and must build the two-level preconditioner, reproduce the reference direction and converge in the
    prototype's iteration count."""
    import optcuts_b200 as ob
    from test_oracle_golden import _merged_adjacency
    s = state1
    (ptr, idx), nVtot = _merged_adjacency(s)
    # the matrix: assembled by the full context, handed over as triplets of the reference CSR (upper triangle)
    s.upload(ctx)
    ctx.set_pattern(ptr, idx, s.fixed)
    ctx.hessian_assemble(s.p0)
    ia, ja, a = ctx.download_csr()
    I = np.repeat(np.arange(len(ia) - 1, dtype=np.int32), np.diff(ia))
    J = (ja - 1).astype(np.int32)
    g, _ = ctx.gradient(s.p0)
    c2 = ob.Context(0)
    try:
        c2.set_coordinate_hint(s.UV)
        c2.set_pattern(ptr, idx, s.fixed)
        info = c2.precond_info()
        assert info["enabled"] and 6 * info["nodes"][-1] <= 3072
        c2.update_values_triplets(I, J, a)
        p, it = c2.solve(-g, 1e-12, 0)
        ref = s.r("searchDir")
        assert np.linalg.norm(p - ref) / np.linalg.norm(ref) < 1e-8
        assert it["iters"] < 260, it
    finally:
        c2.close()


@pytest.mark.parametrize("dense", [0, 1], ids=["block-tridiagonal", "dense"])
def test_direct_safety_net_matches_pcg(state1, monkeypatch, dense):
    """The direct safety net of the solve (csrc/ocb_direct.cu: block-tridiagonal Cholesky over breadth-first levels, or one dense
    potrf of the whole matrix) on the golden state: the reference's LDL^T direction to 1e-8 (the bar of the PCG comparison),
    true residual at rounding level, and it really ran (cuSOLVER / cuBLAS found)."""
    import optcuts_b200 as ob
    if dense:
        monkeypatch.setenv("OCB_DIRECT_DENSE", "1")
    c = ob.Context(0)
    try:
        _newton_system(c, state1)
        x_cg, info = c.solve(None, 1e-12, 0)
        d0 = c.precond_info()["direct_solves"]
        c.set_option("force_direct", 1)
        x_d, _ = c.solve(None, 1e-12, 0)
        c.set_option("force_direct", 0)
        assert c.precond_info()["direct_solves"] == d0 + 1, "the direct path did not run (cuSOLVER / cuBLAS not found?)"
        g, _ = c.gradient(state1.p0)
        res = np.linalg.norm(c.multiply(x_d) + g) / np.linalg.norm(g)
        ref = state1.r("searchDir")
        err_ref = np.linalg.norm(x_d - ref) / np.linalg.norm(ref)
        err_cg = np.linalg.norm(x_d - x_cg) / np.linalg.norm(x_cg)
        print("direct (%s): true residual %.2e, vs the reference's LDL^T direction %.2e, vs the PCG direction %.2e (%d CG iterations)"
              % ("dense" if dense else "block-tridiagonal", res, err_ref, err_cg, info["iters"]))
        assert res < 1e-10 and err_ref < 1e-8 and err_cg < 1e-8
    finally:
        c.close()
