#!/bin/bash
# one gpurun call: parity tests, smoke, bench (both arms), ncu launch list + full captures
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --workload bimba_x4 --steps 3 --warmup 3 --pcg-max-it 30000 --no-cpu-baseline > gpurun_out/bench_x4.json 2> gpurun_out/bench_x4.err; tail -c 2500 gpurun_out/bench_x4.json
python bench.py --impl reference --workload bimba_x4 --steps 2 --warmup 1 > gpurun_out/bench_ref_x4.json 2> gpurun_out/bench_ref_x4.err; tail -c 600 gpurun_out/bench_ref_x4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel|hessian_kernel|energy_kernel|gradient_kernel|step_bound' -s 0 -c 12 -o gpurun_out/prof_x10 python bench.py --workload bimba_x10 --steps 1 --warmup 3 --pcg-max-it 100 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
