"""Multi-GPU path = independent meshes per rank (no data-path collective).  Host-side logic tested with two
`gloo` processes on CPU, as the driver asks: the LPT partition is identical on every rank, covers every mesh
exactly once, and the gathered result table is complete."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from optcuts_b200 import batch
    items = batch.benchmark71()

    def work(item):      # host-only stand-in for the per-mesh GPU run
        return {"name": item[0], "faces": item[1], "rank": rank}
    merged, shards = batch.run_sharded(items, work, rank, world, gather=dist.all_gather_object)
    # value reduction the bench does: max over ranks of the per-rank time
    t = torch.tensor([float(sum(batch.cost_model(items[i][1]) for i in shards[rank]))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({"merged": merged, "shards": shards, "tmax": float(t[0])}, out)
    dist.destroy_process_group()


def test_two_rank_gloo_batch(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    from optcuts_b200 import batch
    items = batch.benchmark71()
    assert len(items) == 71 and sum(it[1] for it in items) == 547725
    assert sorted(r["merged"].keys()) == list(range(71))
    assert sorted(i for s in r["shards"] for i in s) == list(range(71))
    for i, row in r["merged"].items():
        assert row["name"] == items[i][0] and i in r["shards"][row["rank"]]
    loads = [sum(batch.cost_model(items[i][1]) for i in s) for s in r["shards"]]
    assert max(loads) / (sum(loads) / 2) < 1.05                             # LPT balance
    assert abs(r["tmax"] - max(loads)) < 1e-6 * max(loads)


def test_lpt_partition_properties():
    from optcuts_b200 import batch
    rng = np.random.default_rng(0)
    costs = list(rng.uniform(1, 100, 71))
    for world in (1, 2, 4, 8):
        shards = batch.lpt_partition(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(71))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) <= sum(costs) / world + max(costs)                # LPT bound
    assert batch.lpt_partition([], 4) == [[], [], [], []]                   # empty batch


def test_synthetic_disk_is_valid(port):
    from optcuts_b200 import batch
    for faces in (152, 1000, 20000):
        V_rest, F, UV = batch.synthetic_disk(faces, seed=1)
        assert abs(len(F) - faces) < 0.1 * faces + 40
        r8, sc, rc = port.rest_features(V_rest, F)
        assert rc == 0
        u, v = UV[F[:, 1]] - UV[F[:, 0]], UV[F[:, 2]] - UV[F[:, 0]]
        assert np.all(u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0] > 0)             # inversion-free start


def test_visible_device_and_mps_fallback(monkeypatch):
    """the per-mesh processes of a batch: the gpu-th VISIBLE device under a launcher's CUDA_VISIBLE_DEVICES, and an MpsDaemon that
    was told not to start (OCB_BATCH_MPS=0) hands out plain child environments and leaves nothing behind"""
    from optcuts_b200 import batch
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,5")
    assert batch.visible_device(0) == "3" and batch.visible_device(1) == "5" and batch.visible_device(2) == "2"
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES")
    assert batch.visible_device(1) == "1"
    monkeypatch.setenv("OCB_BATCH_MPS", "0")
    with batch.MpsDaemon(1) as mps:
        assert not mps.up and mps.child_env() == {"CUDA_VISIBLE_DEVICES": "1"} and mps.dir is None
