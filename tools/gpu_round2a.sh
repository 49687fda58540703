#!/bin/bash
# round-2 check A: GPU tests (x2 for reproducibility), bench line, compute-sanitizer logs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2a_gpu.txt
for i in 1 2; do
  timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2a_pytest_$i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_$i.log; tail -5 gpurun_out/r2a_pytest_$i.log
done
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -2 gpurun_out/r2a_bench.err
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -4 gpurun_out/r2a_sanitizer_$tool.log
done
for w in bimba_x4 bimba_x10; do
  timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_$w.json 2>gpurun_out/r2a_bench_$w.err
done
python - <<PY
import json
for f in ("r2a_bench","r2a_bench_bimba_x4","r2a_bench_bimba_x10"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "iters", d["config"].get("pcg_iters_mean"))
        for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_launch"]*1000,1),"us", "frac", round(v.get("frac",0) or 0,3))
    except Exception as e: print(f, "ERR", e)
PY
