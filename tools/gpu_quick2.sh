#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/gpu_host_timing.py 2>&1 | tee gpurun_out/host_timing.txt
for w in bimba10k bimba_x4 bimba_x10; do
OCB_PCG_DEBUG=1 python bench.py --workload $w --steps 4 --warmup 3 --pcg-max-it 60000 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
tail -1 gpurun_out/bench_$w.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", "it/s", round(d["value"],3), "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "pcg iters", d["config"]["pcg_iters_mean"], "E", d["E_new"])
for k,v in d["kernels"].items():
    print("   ", k, {a:(round(b,5) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms_per_launch","launches","share")})
PY
done
