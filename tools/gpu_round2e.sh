#!/bin/bash
# round-2 check E: GPU tests (crash backtrace armed), bench lines with AoS gather records
mkdir -p gpurun_out
timeout 3200 python -m pytest tests -q -m gpu -x --durations=10 -s > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -50 gpurun_out/r2e_pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -2 gpurun_out/r2e_bench.err
for w in bimba_x4 bimba_x10; do
  timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench_$w.json 2>gpurun_out/r2e_bench_$w.err
done
python - <<PY
import json
for f in ("r2e_bench","r2e_bench_bimba_x4","r2e_bench_bimba_x10"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "iters", d["config"].get("pcg_iters_mean"))
        for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_launch"]*1000,1),"us", "frac", round(v.get("frac",0) or 0,3))
    except Exception as e: print(f, "ERR", e)
PY
