#!/bin/bash
# round evidence: tests, smoke, both bench arms, launch list, ncu --set full of the dominant kernel (traffic), scaling workloads
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2>/dev/null
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
OCB_PCG_DEBUG=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "ocb pcg" | tail -2 > gpurun_out/pcg_phase_cycles.txt
for w in bimba_x4 bimba_x10; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2>/dev/null
  OCB_PCG_DEBUG=1 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "ocb pcg" | tail -1 >> gpurun_out/pcg_phase_cycles.txt
done
python bench.py --impl reference --workload bimba_x4 --steps 2 --warmup 1 > gpurun_out/bench_ref_x4.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bimba10k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel' -s 3 -c 1 -f -o gpurun_out/prof_pcg10k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel' -s 3 -c 1 -f -o gpurun_out/prof_pcg_x10 python bench.py --workload bimba_x10 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hessian_kernel|energy_kernel|gradient_kernel|step_bound|mas_dense_invert' -s 0 -c 8 -f -o gpurun_out/prof_elem_x10 python bench.py --workload bimba_x10 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python bench.py --workload batch71 --steps 1 > gpurun_out/bench_batch71.json 2>/dev/null
(for ny in 500 1000 2000 4000; do timeout 120 tools/micro/elem_bench 1000 $ny; done) > gpurun_out/elem_bench.txt 2>&1
python tools/gpu_host_timing.py > gpurun_out/host_timing.txt 2>&1
(OCB_MAS_DEBUG=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep 'ocb mas' | tail -1; OCB_MAS_DEBUG=1 python bench.py --workload bimba_x4 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep 'ocb mas' | tail -1; timeout 60 tools/micro/tile_invert_bench) > gpurun_out/mas_dense_phases.txt 2>&1
ls -la gpurun_out | tail -20
python - <<PY
import json
for f in ("bench","bench_ref","bench_bimba_x4","bench_bimba_x10","bench_ref_x4","bench_batch71"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["value"],3), d["unit"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
