#!/bin/bash
# round-2 check C: PCG stopping-norm diagnostic, GPU tests (incl. sweeps + whole runs), bench lines with the new kernels
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag_pcg_norm.py > gpurun_out/r2c_pcg_norm.txt 2>&1; cat gpurun_out/r2c_pcg_norm.txt
timeout 3000 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -25 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -2 gpurun_out/r2c_bench.err
for w in bimba_x4 bimba_x10; do
  timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_$w.json 2>gpurun_out/r2c_bench_$w.err
done
python - <<PY
import json
for f in ("r2c_bench","r2c_bench_bimba_x4","r2c_bench_bimba_x10"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), "iters", d["config"].get("pcg_iters_mean"))
        for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_launch"]*1000,1),"us", "frac", round(v.get("frac",0) or 0,3))
    except Exception as e: print(f, "ERR", e)
PY
python tools/host_program_timing.py > gpurun_out/r2c_host_program.txt 2>&1; tail -12 gpurun_out/r2c_host_program.txt
