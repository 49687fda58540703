"""Batch partitioning for the multi-GPU path: a single mesh is one coupled sparse system, so
multi-GPU work = independent meshes per GPU, no data-path collective.  The reference's batch unit is one
process per mesh, run serially (batch.py:11-14); here the meshes of a batch are assigned to the ranks by
longest-processing-time-first on the face count, every rank runs its shard on its own GPU, and the only
cross-process step is gathering the per-mesh result rows.
"""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def benchmark71():
    """(name, faces, vertices) of the reference's 71-mesh benchmark set (metadata only)."""
    rows = []
    for ln in open(os.path.join(_HERE, "data", "benchmark71.txt")):
        if ln.startswith("#") or not ln.strip():
            continue
        name, f, v = ln.split()
        rows.append((name, int(f), int(v)))
    return rows


def lpt_partition(costs, world):
    """Greedy LPT: items sorted by cost descending go to the currently least-loaded rank.
    Returns a list of index lists, one per rank; deterministic (ties by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    shards = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    return shards


def cost_model(faces):
    """Relative cost of one mesh: Newton iterations scale ~ faces^0.5, each costs ~ faces (PCG iterations ~ faces^0.5)."""
    return float(faces) ** 1.5


def run_sharded(items, work, rank, world, gather=None):
    """Runs work(item) for this rank's shard and gathers {index: result} over all ranks.
    `gather` is torch.distributed.all_gather_object-like (None: single process)."""
    shards = lpt_partition([cost_model(it[1]) for it in items], world)
    mine = {i: work(items[i]) for i in shards[rank]}
    if gather is None or world == 1:
        return mine, shards
    parts = [None] * world
    gather(parts, mine)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged, shards


# ---------------------------------------------------------------------------------------------------------------------
# the real batch: the reference's 71 benchmark meshes through the reference's own host program (one process per mesh,
# the per-mesh command line of batch.py:11-14 in headless mode 100) with the GPU plugins dropped in
ARCHIVE = os.path.join(os.path.dirname(_HERE), "tests", "golden", "inputs", "benchmark71.tar.xz")
CUDA_HOST = os.path.join(os.path.dirname(_HERE), "shim", "_build", "OptCuts_cuda_probe")
REF_HOST = os.path.join(os.path.dirname(_HERE), "oracle", "_ref", "OptCuts_probe")
MESH_ARGS = ["0.999", "1", "0", "4.1", "1", "0"]          # BASELINE.json configs[0] / batch.py: lambda_init 0.999, OptCuts, b_d 4.1, bijective


def extract_benchmark(dst):
    """unpacks the 71 OBJ files (5 MB archive, a data fixture) into dst; returns {name: path}"""
    import tarfile
    with tarfile.open(ARCHIVE) as t:
        t.extractall(dst)
    return {f: os.path.join(dst, f) for f in sorted(os.listdir(dst)) if f.endswith(".obj")}


def visible_device(gpu):
    """the gpu-th VISIBLE device as CUDA_VISIBLE_DEVICES names it (a launcher may have restricted / renumbered the devices already)"""
    vis = [d for d in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if d.strip()]
    return vis[gpu] if gpu < len(vis) else str(gpu)


class MpsDaemon:
    """NVIDIA MPS control daemon for ONE GPU, for the life of a batch: the per-mesh host processes attach to the server's GPU
    context instead of creating their own (1.7-4.7 s per process on this pool's boxes, erratic -> 0.2-0.6 s; profiles/
    r2_mps_process_start.txt) and their kernels run side by side instead of time-sliced.  A batch of one short process per mesh
    (batch.py:11-14) is what MPS is for.  Own pipe / log directory per GPU, so every rank of a multi-GPU batch runs its own daemon;
    nothing is started when the binary is missing, OCB_BATCH_MPS=0, or the daemon does not come up -- the batch then runs as before.
    Use as a context manager; child_env() is what the per-mesh processes need."""

    def __init__(self, gpu):
        self.gpu, self.dev, self.dir, self.up = gpu, visible_device(gpu), None, False

    def __enter__(self):
        import shutil
        import subprocess
        import tempfile
        if os.environ.get("OCB_BATCH_MPS", "1") == "0" or not shutil.which("nvidia-cuda-mps-control"):
            return self
        self.dir = tempfile.mkdtemp(prefix="ocb_mps_%d_" % self.gpu)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=self.dev, CUDA_MPS_PIPE_DIRECTORY=os.path.join(self.dir, "pipe"), CUDA_MPS_LOG_DIRECTORY=os.path.join(self.dir, "log"))
        os.makedirs(env["CUDA_MPS_PIPE_DIRECTORY"]); os.makedirs(env["CUDA_MPS_LOG_DIRECTORY"])
        try:
            self.up = subprocess.run(["nvidia-cuda-mps-control", "-d"], env=env, timeout=30, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
        except Exception:   # noqa: BLE001
            self.up = False
        self._env = env
        return self

    def child_env(self):
        if not self.up:
            return {"CUDA_VISIBLE_DEVICES": self.dev}
        # the server sees exactly one device (index 0 of ITS list)
        return {"CUDA_VISIBLE_DEVICES": "0", "CUDA_MPS_PIPE_DIRECTORY": self._env["CUDA_MPS_PIPE_DIRECTORY"], "CUDA_MPS_LOG_DIRECTORY": self._env["CUDA_MPS_LOG_DIRECTORY"]}

    def __exit__(self, *exc):
        import shutil
        import subprocess
        if self.up:
            try:
                subprocess.run(["nvidia-cuda-mps-control"], input="quit\n", text=True, env=self._env, timeout=60, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            except Exception:   # noqa: BLE001
                pass
            self.up = False
        if self.dir:
            shutil.rmtree(self.dir, ignore_errors=True)
        return False


def run_mesh(exe, mesh_path, workdir, max_iters, gpu=None, timeout=1800, extra_env=None):
    """one mesh through a host program (probe build: ORACLE_MAX_ITERS bounds the run).  Returns dict(rc, wall_s, iters)."""
    import subprocess
    import time
    os.makedirs(workdir, exist_ok=True)
    env = dict(os.environ, ORACLE_TRACE=os.path.join(workdir, "trace.txt"))
    if max_iters:
        env["ORACLE_MAX_ITERS"] = str(int(max_iters))
    if gpu is not None:
        env["CUDA_VISIBLE_DEVICES"] = visible_device(gpu)
    env.update(extra_env or {})
    t0 = time.perf_counter()
    try:
        r = subprocess.run([exe, "100", mesh_path] + MESH_ARGS + ["b"], cwd=workdir, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)
        rc = r.returncode
    except subprocess.TimeoutExpired:
        rc = -999
    dt = time.perf_counter() - t0
    n = 0
    try:
        n = sum(1 for _ in open(os.path.join(workdir, "trace.txt")))
    except OSError:
        pass
    return dict(rc=rc, wall_s=dt, iters=n)


def synthetic_disk(faces, seed=0):
    """A synthetic open triangle mesh with ~`faces` triangles: a jittered grid lifted onto a bumpy surface, with
    a distorted but inversion-free initial UV (stands in for the benchmark meshes, which cannot travel)."""
    n = max(3, int(round(np.sqrt(faces / 2.0))) + 1)
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(n, dtype=float), np.arange(n, dtype=float), indexing="ij")
    P = np.stack([xs.ravel(), ys.ravel()], axis=1)
    P += 0.2 * rng.uniform(-1, 1, P.shape)
    V_rest = np.column_stack([P, 0.15 * n * np.sin(3.0 * P[:, 0] / n) * np.cos(2.0 * P[:, 1] / n)])
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    F = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    # radial squeeze towards the centre (Tutte-like area distortion), orientation preserving
    C = P - P.mean(axis=0)
    r = np.linalg.norm(C, axis=1) / (0.75 * n) + 1e-9
    UV = C * (0.3 + 0.7 * r[:, None] ** 2)
    return V_rest, F, UV
