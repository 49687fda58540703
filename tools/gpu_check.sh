#!/bin/bash
# round-1f check 2: GPU tests, bench lines, PCG phase cycles, host-side timing of the e2e path
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
OCB_PCG_DEBUG=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "ocb pcg" | tail -3 > gpurun_out/pcg_phase_cycles.txt
for w in bimba_x4 bimba_x10; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2>/dev/null
  OCB_PCG_DEBUG=1 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "ocb pcg" | tail -3 >> gpurun_out/pcg_phase_cycles.txt
done
python tools/gpu_host_timing.py > gpurun_out/host_timing.txt 2>&1
cat gpurun_out/pcg_phase_cycles.txt
grep -v "set_mesh" gpurun_out/host_timing.txt | head -30
python - <<PY
import json
for f in ("bench","bench_bimba_x4","bench_bimba_x10"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "iters", d["config"].get("pcg_iters_mean"))
        for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_launch"]*1000,1),"us", "frac", round(v.get("frac",0),3))
    except Exception as e: print(f, "ERR", e)
PY
