"""Exercises every kernel of liboptcuts_b200.so once, at ~1k faces (synthetic grid, every entry point) and on the
golden bimba state with its air mesh (one whole Newton iteration), for `compute-sanitizer --tool memcheck|racecheck|
initcheck|synccheck python tools/sanitize_driver.py` (tools/gpu_sanitize.sh keeps the logs under profiles/).
No torch: the library is driven through ctypes only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import optcuts_b200 as ob  # noqa: E402


def grid_mesh(n):
    xs, ys = np.meshgrid(np.arange(n, dtype=float), np.arange(n, dtype=float), indexing="ij")
    P = np.stack([xs.ravel(), ys.ravel()], axis=1)
    P += 0.2 * np.random.default_rng(0).uniform(-1, 1, P.shape)
    V_rest = np.column_stack([P, 0.3 * np.sin(P[:, 0]) * np.cos(0.7 * P[:, 1])])
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    F = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    UV = P @ np.array([[1.3, 0.2], [-0.1, 0.8]]).T + 0.05 * np.sin(P[:, ::-1])
    return V_rest, F, UV


def small(ctx, quick):
    V_rest, F, UV = grid_mesh(24)                       # 1058 faces
    rest8, sc = ctx.rest_features(V_rest, F)
    ctx.set_mesh(UV.shape[0], F, rest8, sc["surfaceArea"], [0])
    ctx.set_uv(UV)
    ctx.energy(0.975); ctx.energy_per_elem(); ctx.gradient(0.975)
    ctx.set_pattern_from_elements()
    ctx.hessian_assemble(0.975); ctx.hessian_blocks(); ctx.hessian_triplets()
    ia, ja, a = ctx.download_csr()
    ctx.factorize()
    p, info = ctx.solve(None, 1e-12, 0)
    ctx.set_option("force_direct", 1)                   # the direct safety net (csrc/ocb_direct.cu: block-tridiagonal Cholesky, cuSOLVER / cuBLAS blocks)
    pd, _ = ctx.solve(None, 1e-12, 0)
    ctx.set_option("force_direct", 0)
    assert np.linalg.norm(pd - p) <= 1e-7 * np.linalg.norm(p), "direct solve differs from PCG"
    ctx.multiply(p)
    al = ctx.step_bound(None, 1.0)
    ctx.save_uv()
    ctx.line_search(0.975, 0.0, 0.99 * al)
    ctx.restore_uv()
    ctx.newton_step(0.975, 0.0)
    ctx.divgrad_scores()
    ctx.get_uv()
    coh = np.array([[0, 1, 1, 0], [5, 6, 6, 5]], np.int32)
    ctx.seam_energy(coh, np.array([1.0, 2.0]), np.array([0, 0], np.int32), 0.0, 1.0, 1.0, True)
    I = np.repeat(np.arange(len(ia) - 1, dtype=np.int32), np.diff(ia)); J = (ja - 1).astype(np.int32)
    ctx.update_values_triplets(I, J, a)
    ctx.solve(np.ones(2 * UV.shape[0]), 1e-10, 0)
    from stencils import one_ring
    batch = [one_ring(F, V_rest, UV, [v]) for v in (100, 200, 300)] + [one_ring(F, V_rest, UV, [150, 151])]
    out = ctx.eval_stencils(batch)
    assert np.all(out["status"] == 0)
    print("small: ok, pcg iters", info["iters"], "launches", ctx.launch_count())


def golden(ctx):
    from conftest import State, GOLDEN
    g = np.load(os.path.join(GOLDEN, "bimba_cfg2_states.npz"))
    s = State({k: g[k] for k in g.files}, "s1_", "r1_", "s2_")
    s.upload(ctx)
    ctx.set_pattern_from_elements()
    r = ctx.newton_step(s.p0, 0.0, 1e-12 if "--tight" in sys.argv else 1e-6)
    print("golden: ok, E_new %.12g, pcg iters %d" % (r["E_new"], r["pcg_iters"]))


if __name__ == "__main__":
    c = ob.Context(0)
    small(c, "--quick" in sys.argv)
    if "--no-golden" not in sys.argv:
        golden(c)
    c.close()
