#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command + full captures of the kernels
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bimba10k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel|hessian_kernel|energy_kernel|gradient_kernel|step_bound' -s 0 -c 14 -o gpurun_out/prof_x10 python bench.py --workload bimba_x10 --steps 1 --warmup 3 --pcg-max-it 100 --no-cpu-baseline > gpurun_out/ncu_full_x10.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel' -s 2 -c 2 -o gpurun_out/prof_pcg10k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_10k.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python bench.py --workload bimba_x4 --steps 3 --warmup 3 --pcg-max-it 60000 > gpurun_out/bench_x4.json 2>/dev/null
python bench.py --workload bimba_x10 --steps 2 --warmup 3 --pcg-max-it 80000 > gpurun_out/bench_x10.json 2>/dev/null
python bench.py --workload batch71 --steps 1 > gpurun_out/bench_batch71.json 2>/dev/null
python bench.py --workload batch71 --impl reference > gpurun_out/bench_batch71_ref.json 2>/dev/null
ls -la gpurun_out
