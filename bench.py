#!/usr/bin/env python
"""bench.py — Newton iterations / s of the OptCuts geometry step on B200 (BASELINE.json's metric).

A "step" is ONE Newton iteration of Optimizer::solve(1) (Optimizer.cpp:203-261, 505-673): gradient ->
convergence test -> element Hessians + PSD projection + scatter into the BSR matrix -> linear solve ->
step bound -> line search, on mesh + air-mesh (scaffold) elements, everything fp64.

Workloads (config.workload):
  bimba10k   (default; BASELINE.json configs[1]) bimba_i_f10000, lambda 0.025, bijectivity on: the
             iteration is replayed from two reference states recorded from the reference run
             (tests/golden: it=1 = Tutte start, it=100 = 37 seam edges), alternating between them.
  bimba_x4 / bimba_x10   (configs[2]) the it=1 state subdivided 4x4 / 10x10 per triangle: 159 984 /
             999 900 faces, no scaffold (the air mesh is the host program's Triangle call).

`value`  = steps / device time with the state resident in HBM (UV restored from a device snapshot);
`e2e`    = the same through the C-ABI with HOST buffers: every step uploads the UVs (pinned host memory) and, with
           bijectivity on, the air mesh + rebuilds the sparsity pattern (as the host program must after every
           scaffold re-triangulation), runs the iteration and downloads the new UVs + the result scalars.
The default line (no --workload) keeps bimba10k as `value` / `e2e` and adds, measured in the same run:
  "workloads"     the same measurement (value, e2e, roofline, kernels, cpu_baseline) for bimba_x4 and bimba_x10,
  "batch71"       BASELINE.json configs[3]: the reference's 71 benchmark meshes (tests/golden/inputs/benchmark71.tar.xz) through
                  the reference's own host program with the GPU plugins (shim/_build/OptCuts_cuda_probe, one process per mesh,
                  the per-mesh command line of batch.py:11-14), LPT-sharded over the ranks, every mesh bounded to
                  --batch-iters Newton iterations (150), --batch-procs processes at a time per GPU (3) attached to an MPS daemon
                  started for the batch: batch wall time (max over ranks), per-rank load, slowest mesh,
  "host_program"  the WHOLE run of configs[1] (geometry + topology to convergence) in the host program: GPU plugins vs the
                  unmodified reference on this box's host cores, with the reference's own info.txt timers of both.
`--impl reference` times the reference's own CPU implementation of the same step
(oracle/_ref/liboptcuts_ref.so = unmodified OptCuts::Optimizer with Eigen SimplicialLDLT, TBB shim on all
host cores) from the same states: wall time of solve(1) minus the reference's own "scaffolding" timer.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden", "bimba_cfg2_states.npz")

# algorithmic HBM bytes per unit of work (restated in DESIGN.md §4)
BYTES_PER_FACE = {"energy": 60.0, "gradient": 68.0, "step_bound": 28.0, "hessian_psd_scatter": 172.0}
PCG_BYTES_PER_FACE_PER_ITER = 228.0          # BSR(2x2) SpMV 144 + block-Jacobi + dots + axpys
# the two-level preconditioner's own compulsory traffic per CG iteration: fp32 group inverses of the
# levels below the coarse one, 48x48 per 8 nodes = 2304*4 B per 64 vertices per level (x 8/7 for the geometric series
# of levels) = 82 B/face, one extra pass over z (write + read, 16 B/face), and the fp32 coarse inverse, 4 * nC^2 bytes
MAS_BYTES_PER_FACE_PER_ITER = 82.0 + 16.0


def pcg_bytes_per_iter(faces, n_coarse):
    return (PCG_BYTES_PER_FACE_PER_ITER + (MAS_BYTES_PER_FACE_PER_ITER if n_coarse else 0.0)) * faces + 4.0 * n_coarse * n_coarse


def load_states(workload):
    g = np.load(GOLDEN)
    p0 = float(g["energyParam0"])

    def st(tag, rt):
        return dict(V_rest=g[tag + "V_rest"], F=g[tag + "F"], UV=g[tag + "V"], fixed=g[tag + "fixedVert"],
                    rest8=g[rt + "rest8"], surfaceArea=float(g[rt + "surfaceArea"]), targetGRes=float(g[rt + "targetGRes"]),
                    air=dict(V=g[tag + "air_V"], F=g[tag + "air_F"], l2g=g[tag + "air_localVI2Global"],
                             nBnd=len(g[tag + "air_bnd"]), rest8=g[rt + "air_rest8"], fixed=g[rt + "air_fixed"]),
                    w_scaf=float(g[rt + "w_scaf"]), cohE=g[tag + "cohE"])
    if workload == "bimba10k":
        return [st("s1_", "r1_"), st("s100_", "r100_")], p0
    n = {"bimba_x4": 4, "bimba_x10": 10}[workload]
    from optcuts_b200 import synth
    s = st("s1_", "r1_")
    Vr, F, UV = synth.subdivide(s["V_rest"], s["F"], s["UV"], n)
    return [dict(V_rest=Vr, F=F, UV=UV, fixed=np.array([0], np.int32), rest8=None, surfaceArea=None,
                 targetGRes=s["targetGRes"], air=None, w_scaf=0.0, cohE=np.zeros((0, 4), np.int32))], p0


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                c = [x.strip() for x in ln.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ our arm
def setup_context(ob, dev, s, p0, stream):
    ctx = ob.Context(dev)
    # every kernel, copy, the L2 flush and the timing events run on ONE explicit stream (a handle of 0 would be the
    # legacy default stream; the library now takes the handle as given)
    ctx.set_stream(stream.cuda_stream)
    if s["rest8"] is None:
        s["rest8"], sc = ctx.rest_features(s["V_rest"], s["F"])
        s["surfaceArea"] = sc["surfaceArea"]
    ctx.set_mesh(s["UV"].shape[0], s["F"], s["rest8"], s["surfaceArea"], s["fixed"])
    ctx.set_uv(s["UV"])
    a = s["air"]
    if a is not None:
        ctx.set_air(a["F"], a["rest8"], a["l2g"], a["nBnd"], a["fixed"], s["w_scaf"] / a["F"].shape[0])
        ctx.set_uv(None, a["V"])
    ctx.set_pattern_from_elements()
    ctx.save_uv()
    return ctx


def measure_workload(args, workload, K, W, torch, ob, local, world, stream, flush, with_cpu_baseline):
    """value / e2e / roofline / kernels of one workload on this rank's GPU (max over ranks inside)"""
    states, p0 = load_states(workload)
    ctxs = [setup_context(ob, local, s, p0, stream) for s in states]
    pcg_tol, pcg_max = args.pcg_tol, args.pcg_max_it

    def step_resident(i):
        c, s = ctxs[i % len(ctxs)], states[i % len(ctxs)]
        c.restore_uv()
        return c.newton_step(p0, s["targetGRes"], pcg_tol, pcg_max)

    # pinned host staging for the e2e leg
    pinned = []
    for s in states:
        nV = s["UV"].shape[0]
        hV = torch.empty((2, nV), dtype=torch.float64).pin_memory()
        hV.numpy()[:] = np.asfortranarray(s["UV"]).T
        hVa = None
        if s["air"] is not None:
            hVa = torch.empty((2, s["air"]["V"].shape[0]), dtype=torch.float64).pin_memory()
            hVa.numpy()[:] = np.asfortranarray(s["air"]["V"]).T
        hOut = torch.empty((2, nV), dtype=torch.float64).pin_memory()
        pinned.append((hV, hVa, hOut))

    import ctypes as C
    _d = C.POINTER(C.c_double)

    def step_e2e(i):
        k = i % len(ctxs)
        c, s = ctxs[k], states[k]
        hV, hVa, hOut = pinned[k]
        L, h = c._L, c._h
        a = s["air"]
        if a is not None:
            # what the host program does every Newton iteration with bijectivity on (Optimizer.cpp:236-239, 522-530):
            # hand over the freshly triangulated air mesh and rebuild the sparsity pattern
            c.set_air(a["F"], a["rest8"], a["l2g"], a["nBnd"], a["fixed"], s["w_scaf"] / a["F"].shape[0])
        c._chk(L.ocb_set_uv(h, C.cast(hV.data_ptr(), _d), C.cast(hVa.data_ptr(), _d) if hVa is not None else None))
        if a is not None:
            c.set_pattern_from_elements()
        r = c.newton_step(p0, s["targetGRes"], pcg_tol, pcg_max)
        c._chk(L.ocb_get_uv(h, C.cast(hOut.data_ptr(), _d), None))
        return r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, K, W, profile=False):
        for i in range(W):
            fn(i)
        for c in ctxs:
            c.profile_enable(profile)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        l0 = sum(c.launch_count() for c in ctxs)
        res = []
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            flush.zero_()                      # L2 flush between timed iterations (outside the event pair, same stream)
            ev[i][0].record()
            res.append(fn(i))
            ev[i][1].record()
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        launches = sum(c.launch_count() for c in ctxs) - l0
        return ms, wall, launches, res

    ms, wall, launches, res = timed(step_resident, K, W, profile=True)
    prof = {}
    for c in ctxs:
        for name, (pms, cnt) in c.profile_get().items():
            a = prof.setdefault(name, [0.0, 0])
            a[0] += pms; a[1] += cnt
        c.profile_enable(False)
    ms_e2e, wall_e2e, _, res_e2e = timed(step_e2e, K, W)

    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    value = world * K / (ms * 1e-3)
    e2e_value = world * K / (ms_e2e * 1e-3)

    # roofline of the dominant kernel class, from the live per-class CUDA-event profile
    faces = np.mean([s["F"].shape[0] + (s["air"]["F"].shape[0] if s["air"] is not None else 0) for s in states])
    total_prof = sum(v[0] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k][0])
    pcg_iters = float(np.mean([r["pcg_iters"] for r in res]))
    pinfo = [c.precond_info() for c in ctxs]
    n_coarse = float(np.mean([6 * p["nodes"][-1] if p["enabled"] else 0 for p in pinfo]))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")

    def kernel_line(name):
        pms, cnt = prof[name]
        if cnt == 0:
            return None
        per_launch_ms = pms / cnt
        if name == "pcg":
            bytes_per_launch = pcg_bytes_per_iter(faces, n_coarse) * pcg_iters
        elif name in BYTES_PER_FACE:
            bytes_per_launch = BYTES_PER_FACE[name] * faces
        else:
            return dict(ms_per_launch=per_launch_ms, launches=cnt, share=pms / total_prof)
        gbs = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        return dict(ms_per_launch=per_launch_ms, launches=cnt, share=pms / total_prof, alg_bytes=bytes_per_launch,
                    achieved_gbs=gbs, frac=gbs / peak)
    kernels = {k: kernel_line(k) for k in prof if prof[k][1] > 0}
    d = kernels[dom]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(workload, {}).get(dom)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d.get("achieved_gbs"), "peak": peak, "peak_source": peak_src,
                "unit": "GB/s", "frac": d.get("frac"), "traffic": traffic,
                "note": ("algorithmic bytes = (%.0f B/face [SpMV + vectors] + %.0f B/face [preconditioner levels] ) x faces + 4 x %d^2 [coarse inverse] "
                         "per CG iteration, x CG iterations per launch%s" % (PCG_BYTES_PER_FACE_PER_ITER, MAS_BYTES_PER_FACE_PER_ITER if n_coarse else 0.0, int(n_coarse),
                                                                              "; the working set of this workload (< 4 MB) stays in shared memory / L2 for the whole solve, so the "
                                                                              "kernel is bound by barrier and instruction latency on 16 SMs, not by HBM (SURVEY 8d: quote it/s here, "
                                                                              "the HBM fraction on the >= 100k-face workloads bimba_x4 / bimba_x10)" if faces < 50000 else ""))
                if dom == "pcg" else "algorithmic bytes per face x faces"}

    def air_bytes(s):
        a = s["air"]
        return 0 if a is None else a["F"].nbytes + a["rest8"].nbytes + a["l2g"].nbytes + 16 * a["V"].shape[0]
    h2d = int(np.mean([16 * s["UV"].shape[0] + air_bytes(s) for s in states]))
    d2h = int(np.mean([16 * s["UV"].shape[0] for s in states])) + 16 * 8
    out = {
        "value": value, "unit": "it/s", "steps": K, "warmup": W, "ms_per_step": ms / K, "wall_ms_per_step_incl_flush": 1e3 * wall / K,
        "data": "reference states of bimba_i_f10000 (recorded from the reference run, tests/golden)" if workload == "bimba10k"
                else "synthetic: bimba Tutte state subdivided",
        "config": {"workload": workload, "faces": int(states[0]["F"].shape[0]), "states": len(states),
                   "air_faces": [int(s["air"]["F"].shape[0]) if s["air"] is not None else 0 for s in states],
                   "pcg_rel_tol": pcg_tol, "pcg_iters_mean": pcg_iters,
                   "preconditioner": ("two-level additive Schwarz: levels %s, exact coarse inverse of %d DOFs" % (pinfo[0]["nodes"], 6 * pinfo[0]["nodes"][-1]))
                                     if pinfo[0]["enabled"] else "block-Jacobi", "l2": "flushed between timed iterations (256 MB memset on the timing stream)",
                   "parallelism": "independent meshes per GPU, no collective"},
        "e2e": {"value": e2e_value, "unit": "it/s", "ms_per_step": ms_e2e / K, "wall_ms_per_step_incl_flush": 1e3 * wall_e2e / K,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels,
        "E_new": [res[i]["E_new"] for i in range(min(len(states), len(res)))],
    }
    for c in ctxs:
        c.close()
    if with_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, workload, budget_s=20.0)
    return out


def run_ours(args):
    import torch
    import optcuts_b200 as ob
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)                # torch's current stream for the rest of the run: flush + events land on it
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    K, W = args.steps, max(args.warmup, 3)
    full = args.workload is None                 # the default line: headline workload + the sub-objects
    headline = args.workload or "bimba10k"
    cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    sampler = ClockSampler(local)
    sampler.start()
    main = measure_workload(args, headline, K, W, torch, ob, local, world, stream, flush, cpu)
    clocks = sampler.stop()
    line = {"metric": "newton_iters_per_s", "value": main["value"], "unit": "it/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64"}
    line.update({k: v for k, v in main.items() if k not in line})
    line["clocks"] = clocks
    if full and not args.quick:
        line["workloads"] = {}
        for wl in ("bimba_x4", "bimba_x10"):
            if world == 1:
                try:
                    line["workloads"][wl] = measure_workload(args, wl, min(K, 6), 3, torch, ob, local, world, stream, flush, cpu)
                except Exception as e:   # noqa: BLE001
                    line["workloads"][wl] = {"unavailable": repr(e)}
            else:
                line["workloads"][wl] = measure_workload(args, wl, min(K, 6), 3, torch, ob, local, world, stream, flush, cpu)
        del flush
        torch.cuda.empty_cache()
        if world == 1:                                   # a failing sub-leg must not take the headline line with it
            try:
                line["batch71"] = batch71_ours(args, rank, world, local, torch)
            except Exception as e:   # noqa: BLE001
                line["batch71"] = {"unavailable": repr(e)}
        else:                                            # (with several ranks every rank has to reach the collectives inside)
            line["batch71"] = batch71_ours(args, rank, world, local, torch)
        if rank == 0 and world == 1:
            line["host_program"] = host_program_run()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- batch + host program
def batch71_ours(args, rank, world, local, torch):
    """the 71 benchmark meshes through the host program with the GPU plugins, one process per mesh, this rank's LPT shard
    on this rank's GPU; wall time = max over ranks (strong scaling: the batch is fixed)"""
    from optcuts_b200 import batch
    import tempfile
    items = batch.benchmark71()
    shards = batch.lpt_partition([batch.cost_model(it[1]) for it in items], world)
    if not (os.path.exists(batch.CUDA_HOST) and os.path.exists(batch.ARCHIVE)):
        return {"unavailable": "shim/_build/OptCuts_cuda_probe or the mesh archive is missing"}
    rows = {}
    with tempfile.TemporaryDirectory() as wd, batch.MpsDaemon(local) as mps:
        paths = batch.extract_benchmark(os.path.join(wd, "in"))
        cenv = mps.child_env()
        warm = batch.run_mesh(batch.CUDA_HOST, paths[items[shards[rank][-1]][0]], os.path.join(wd, "warm"), 1, extra_env=cenv)   # smallest mesh of the shard, 1 iteration
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        # `procs` host-program processes at a time on this rank's GPU (largest mesh first): a process spends most of its life in
        # host code (OBJ parsing, Triangle, mesh edits, the reference's query code), so the GPU is shared (side by side under MPS)
        from concurrent.futures import ThreadPoolExecutor
        procs = max(1, min(args.batch_procs, (os.cpu_count() or 1) // max(1, world)))
        # torchrun exports OMP_NUM_THREADS=1 to its ranks and the per-mesh processes would inherit it: the direct safety net's library
        # calls (cuSOLVER potrf) then run their host part on one thread -- male_2 took 70-93 s of a 2-GPU batch instead of 11 s
        # (profiles/r2_bench_n2_before_omp.json).  Every rank's processes get the rank's share of the cores instead (1 GPU: all cores,
        # which is what an unset OMP_NUM_THREADS means; a third of that per process cost the 1-GPU batch 25 %:
        # profiles/r2_bench_batch71_omp5.json, 126 s against 101 s).
        threads = max(4, (os.cpu_count() or 1) // max(1, world))
        cenv = dict(cenv, OMP_NUM_THREADS=str(threads))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=procs) as ex:
            res = list(ex.map(lambda i: batch.run_mesh(batch.CUDA_HOST, paths[items[i][0]], os.path.join(wd, "m%d" % i), args.batch_iters, extra_env=cenv), shards[rank]))
        rows = dict(zip(shards[rank], res))
        mine = time.perf_counter() - t0
        mps_up = mps.up
    its = sum(r["iters"] for r in rows.values())
    bad = [items[i][0] for i, r in rows.items() if r["rc"] != 0]
    slow = max(rows, key=lambda i: rows[i]["wall_s"])
    per_rank = [mine]
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([mine, float(its), float(len(bad))], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [float(x[0]) for x in allt]
        its = int(sum(float(x[1]) for x in allt))
        nbad = int(sum(float(x[2]) for x in allt))
    else:
        nbad = len(bad)
    wall = max(per_rank)
    return {"meshes": len(items), "newton_iters_per_mesh_cap": args.batch_iters, "newton_iters": int(its), "wall_s": wall, "it_per_s": its / wall,
            "per_rank_s": per_rank, "limiting_rank": int(np.argmax(per_rank)), "failed_meshes": nbad, "failed_on_rank0": bad,
            "slowest_mesh_rank0": {"name": items[slow][0], "faces": items[slow][1], "wall_s": rows[slow]["wall_s"]},
            "one_iteration_process_wall_s": warm["wall_s"], "concurrent_processes_per_gpu": procs, "omp_threads_per_process": threads, "mps": bool(mps_up),
            "scaling": "strong",
            "note": "one host-program process per mesh (reference main + Optimizer hooks + device candidate evaluation), config args %s, every mesh "
                    "bounded to the cap, process start-up inside; mps = the per-mesh processes attach to an NVIDIA MPS daemon started for the batch "
                    "(own CUDA context creation costs 1.7-4.7 s per process on this pool's boxes, 0.2-0.6 s through MPS: profiles/r2_mps_process_start.txt); "
                    "one_iteration_process_wall_s = the fixed cost of one process" % " ".join(batch.MESH_ARGS)}


def host_program_run():
    """whole run of BASELINE.json configs[1] in the host program: GPU plugins vs the unmodified reference, this box"""
    import subprocess
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "host_program_timing.py")], capture_output=True, text=True, errors="replace",
                             timeout=900, env=dict(os.environ, OCB_TIMING_SKIP="cuda0")).stdout
    except Exception as e:   # noqa: BLE001
        return {"unavailable": str(e)}
    res = {"config": "bimba_i_f10000.obj 0.025 1 2 4.1 1 0 (geometry + topology to convergence)", "cores": os.cpu_count()}
    cur = None
    for ln in out.split("\n"):
        ln = ln.strip()
        if ln.startswith("-- "):
            cur = ln.split()[1]
            res[cur] = {"wall_s": float(ln.split("wall")[1].split()[0])}
        elif cur and ln.startswith("iterations:"):
            res[cur]["newton_iters_topo_steps"] = ln.split(":", 1)[1].split()[:2]
        elif cur and ln.startswith("timers:"):
            res[cur]["info_txt_timers"] = ln.split(":", 1)[1].strip()
        elif cur and ln.startswith("final"):
            res[cur]["final_E_SD_E_se"] = ln.split(":", 1)[1].split()
        elif cur and ln.startswith("[ocb candidates]"):
            w = ln.split()
            res[cur]["candidates"] = {"queries": int(w[2]), "local_problems": int(w[4]), "lock_step_rounds": int(w[11]), "total_s": float(w[19]),
                                      "triangle_s": float(w[22]), "device_and_packing_s": float(w[27])}
    if "cuda" in res and "ref" in res:
        res["speedup_whole_run"] = res["ref"]["wall_s"] / res["cuda"]["wall_s"]

        def tm(d, key):      # one entry of the reference's own info.txt timer line ("topo1.29 desc0.47 ... bSplit0.10 iSplit0.08 cMerge0.0003")
            return next((float(t[len(key):]) for t in d.get("info_txt_timers", "").split() if t.startswith(key)), None)
        try:                 # kernel (6): nested local solves of querySplit / queryMerge (timer_step boundarySplit + interiorSplit + cornerMerge)
            ours, ref = (sum(tm(res[k], t) for t in ("bSplit", "iSplit", "cMerge")) for k in ("cuda", "ref"))
            n = res["cuda"].get("candidates", {}).get("local_problems")
            res["speedup_in_process_timers"] = tm(res["ref"], "topo") and (sum(tm(res["ref"], t) for t in ("topo", "desc", "scaf")) / sum(tm(res["cuda"], t) for t in ("topo", "desc", "scaf")))
            if n:
                res["candidate_evaluation"] = {"local_problems": n, "ours_s": ours, "reference_s": ref, "ours_candidates_per_s": n / ours,
                                               "reference_candidates_per_s": n / ref, "ratio": ref / ours,
                                               "note": "same local problems in both arms when the op sequences agree; seconds = the reference's timer_step boundarySplit + interiorSplit + cornerMerge of info.txt in each arm"}
        except Exception:    # noqa: BLE001
            pass
    return res


# ------------------------------------------------------------------------------------- reference arm
def ref_step_times(workload, n_steps, budget_s):
    """Times Optimizer::solve(1) of the UNMODIFIED reference (oracle/_ref) from the workload's states."""
    from oracle import refapi
    states, p0 = load_states(workload)
    kind = "reference"
    if not refapi.available():
        raise RuntimeError("oracle/_ref/liboptcuts_ref.so is missing (built by `make -C oracle ref` where /root/reference exists)")
    out_dir = tempfile.mkdtemp(prefix="ocb_ref_")
    refapi.set_output_folder(out_dir + "/")
    times, t_begin = [], time.perf_counter()
    # the reference chats on stdout per iteration (Optimizer.cpp:212,582,612,642): keep the JSON line clean
    sys.stdout.flush()
    saved_fd = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        return _ref_loop(refapi, states, p0, n_steps, budget_s, times, t_begin, 0), kind
    finally:
        os.dup2(saved_fd, 1)
        os.close(saved_fd); os.close(devnull)


def _ref_loop(refapi, states, p0, n_steps, budget_s, times, t_begin, i):
    while len(times) < n_steps and (time.perf_counter() - t_begin < budget_s or len(times) < 1):
        s = states[i % len(states)]
        i += 1
        m = refapi.RefMesh(s["V_rest"], s["F"], s["UV"], cohE=s["cohE"] if len(s["cohE"]) else None)
        opt = refapi.RefOptimizer(m, p0, scaffolding=s["air"] is not None, mute=False)     # precompute(): untimed
        refapi.timers_reset()
        t0 = time.perf_counter()
        opt.solve(1)
        dt = time.perf_counter() - t0
        t4, _ = refapi.timers()
        times.append(dt - t4["scaffolding"])
        opt.close(); m.close()
    return times


def cpu_baseline(args, workload, budget_s):
    """bounded sample of the reference's CPU implementation of the same step, this box's host cores.  The 1M-face workload
    runs in a child process under a time limit: one reference iteration there is minutes of sparse LDL^T."""
    if workload == "bimba_x10":
        import subprocess
        limit = args.cpu_limit_x10
        t0 = time.perf_counter()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0"],
                               capture_output=True, text=True, errors="replace", timeout=limit)
            ln = [x for x in r.stdout.split("\n") if x.startswith("{")]
            d = json.loads(ln[-1])
            if "cpu_baseline" in d:
                return d["cpu_baseline"]
            return {"unavailable": d.get("unavailable", "no result")}
        except subprocess.TimeoutExpired:
            return {"unavailable": "one reference Newton iteration (precompute + solve(1), Eigen SimplicialLDLT) did not finish within %d s on %d cores"
                                   % (limit, os.cpu_count()), "lower_bound_s_per_step": time.perf_counter() - t0, "cores": os.cpu_count(), "kind": "reference"}
        except Exception as e:   # noqa: BLE001
            return {"unavailable": str(e)}
    try:
        times, kind = ref_step_times(workload, n_steps=64, budget_s=budget_s)
    except Exception as e:   # noqa: BLE001
        return {"unavailable": str(e)}
    return {"value": len(times) / sum(times), "unit": "it/s", "cores": os.cpu_count(), "kind": kind,
            "ms_per_step": 1e3 * sum(times) / len(times),
            "sample": "%d Newton iterations of the unmodified reference (Optimizer::solve(1), Eigen SimplicialLDLT, TBB shim on all cores) "
                      "from the same states; wall time minus the reference's own scaffolding timer" % len(times)}


def reference_line(workload, K, W, budget):
    times, kind = ref_step_times(workload, n_steps=K + W, budget_s=budget)
    t = times[W:] if len(times) > W else times
    v = len(t) / sum(t)
    states, _ = load_states(workload)
    return {"value": v, "unit": "it/s", "steps": len(t), "warmup": W, "ms_per_step": 1e3 * sum(t) / len(t),
            "data": "same states as the GPU arm",
            "config": {"workload": workload, "faces": int(states[0]["F"].shape[0]), "states": len(states)},
            "cpu_baseline": {"value": v, "unit": "it/s", "cores": os.cpu_count(), "kind": kind,
                             "sample": "%d timed Newton iterations (Optimizer::solve(1) minus its scaffolding timer), rank 0 only" % len(t)},
            "e2e": {"value": v, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def batch71_reference(args, world):
    """the reference arm of the batch: `world` concurrent CPU processes (one mesh each at a time), same meshes, same cap"""
    from optcuts_b200 import batch
    from concurrent.futures import ThreadPoolExecutor
    items = batch.benchmark71()
    if not (os.path.exists(batch.REF_HOST) and os.path.exists(batch.ARCHIVE)):
        return {"unavailable": "oracle/_ref/OptCuts_probe or the mesh archive is missing"}
    pick = list(range(len(items)))
    if args.batch_ref_sample and args.batch_ref_sample < len(items):          # bounded sample: every k-th mesh by size
        order = sorted(pick, key=lambda i: items[i][1])
        pick = order[::max(1, len(order) // args.batch_ref_sample)][:args.batch_ref_sample]
    with tempfile.TemporaryDirectory() as wd:
        paths = batch.extract_benchmark(os.path.join(wd, "in"))
        order = sorted(pick, key=lambda i: -items[i][1])
        t0 = time.perf_counter()
        conc = max(1, min(world * args.batch_procs, os.cpu_count() or 1))      # as many host processes at a time as the GPU arm runs
        with ThreadPoolExecutor(max_workers=conc) as ex:
            rows = list(ex.map(lambda i: batch.run_mesh(batch.REF_HOST, paths[items[i][0]], os.path.join(wd, "m%d" % i), args.batch_iters), order))
        wall = time.perf_counter() - t0
    its = sum(r["iters"] for r in rows)
    return {"meshes": len(pick), "of": len(items), "newton_iters_per_mesh_cap": args.batch_iters, "newton_iters": int(its), "wall_s": wall,
            "it_per_s": its / wall, "concurrent_processes": conc, "cores": os.cpu_count(), "failed_meshes": sum(1 for r in rows if r["rc"] != 0),
            "note": "unmodified reference (oracle/_ref/OptCuts_probe), %d meshes at a time, largest first" % conc}


def run_reference(args):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 1)
    headline = args.workload or "bimba10k"
    try:
        main = reference_line(headline, K, W, 150.0)
    except Exception as e:   # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": str(e)}))
        return
    line = {"impl": "reference", "metric": "newton_iters_per_s", "value": main["value"], "unit": "it/s", "n_gpus": world, "steps": main["steps"],
            "warmup": W, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64"}
    line.update({k: v for k, v in main.items() if k not in line})
    if args.workload is None and not args.quick:
        line["workloads"] = {}
        try:
            line["workloads"]["bimba_x4"] = reference_line("bimba_x4", 2, 0, 30.0)
        except Exception as e:   # noqa: BLE001
            line["workloads"]["bimba_x4"] = {"unavailable": str(e)}
        line["batch71"] = batch71_reference(args, world)
    print(json.dumps(line))


def run_batch(args):
    """--workload batch71: the batch alone (both arms)"""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if rank != 0:
            return
        b = batch71_reference(args, world)
        v = b.get("it_per_s")
        print(json.dumps({"impl": "reference", "metric": "newton_iters_per_s", "value": v, "unit": "it/s", "n_gpus": world, "steps": 1, "warmup": 0,
                          "ms_per_step": 1e3 * b.get("wall_s", 0.0), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "the reference's 71 benchmark meshes", "config": {"workload": "batch71"}, "batch71": b,
                          "cpu_baseline": {"value": v, "unit": "it/s", "cores": os.cpu_count(), "kind": "reference", "sample": b.get("note")},
                          "e2e": {"value": v, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local); sampler.start()
    b = batch71_ours(args, rank, world, local, torch)
    clocks = sampler.stop()
    if rank == 0:
        v = b.get("it_per_s")
        print(json.dumps({"metric": "newton_iters_per_s", "value": v, "unit": "it/s", "n_gpus": world, "steps": 1, "warmup": 1,
                          "ms_per_step": 1e3 * b.get("wall_s", 0.0), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "the reference's 71 benchmark meshes (tests/golden/inputs/benchmark71.tar.xz)",
                          "config": {"workload": "batch71", "parallelism": "LPT shard of independent meshes per GPU, no collective"},
                          "e2e": {"value": v, "unit": "it/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}, "batch71": b, "clocks": clocks}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["bimba10k", "bimba_x4", "bimba_x10", "batch71"],
                    help="default: bimba10k as the headline + bimba_x4 / bimba_x10 / batch71 / host_program sub-objects")
    ap.add_argument("--quick", action="store_true", help="headline workload only (no sub-objects)")
    ap.add_argument("--batch-iters", type=int, default=150, help="batch71: cap of Newton iterations per mesh (both arms)")
    ap.add_argument("--batch-procs", type=int, default=3, help="batch71: host-program processes at a time per GPU (the reference arm runs gpus x this many CPU processes)")
    ap.add_argument("--batch-ref-sample", type=int, default=24, help="reference arm of batch71: number of meshes sampled across the size range (0 = all 71)")
    ap.add_argument("--cpu-limit-x10", type=int, default=45, help="time limit (s) of the reference's iteration at 1M faces")
    ap.add_argument("--pcg-tol", type=float, default=1e-12)
    ap.add_argument("--pcg-max-it", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload == "batch71":
        run_batch(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
