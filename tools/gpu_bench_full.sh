#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 700 gpurun_out/bench_ref.json; echo
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench.json"))
print({k:d[k] for k in ("value","ms_per_step","steps","gpu_launches")}, d["e2e"], d["roofline"], d.get("cpu_baseline"), d["config"], d["clocks"])
PY
OCB_PCG_DEBUG=1 python bench.py --workload bimba_x4 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bimba_x4.json 2> gpurun_out/bench_bimba_x4.err; tail -1 gpurun_out/bench_bimba_x4.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_bimba_x4.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"], d["roofline"])
PY
