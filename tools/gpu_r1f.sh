#!/bin/bash
# round-1f check: GPU parity tests, bench lines (10k / x4 / x10) and an ncu --set full capture of the element kernels at 1M faces
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
for w in bimba_x4 bimba_x10; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2>/dev/null
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hessian_kernel|energy_kernel|gradient_kernel|step_bound|sqnorm' -s 0 -c 10 -f -o gpurun_out/prof_elem_x10 python bench.py --workload bimba_x10 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import json
for f in ("bench","bench_bimba_x4","bench_bimba_x10"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"))
        for k,v in d["kernels"].items(): print("   ",k, round(v["ms_per_launch"]*1000,1),"us", "frac", round(v.get("frac",0),3))
    except Exception as e: print(f, "ERR", e)
PY
