#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel' -s 3 -c 1 -f -o gpurun_out/prof_pcg10k python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_10k.log 2>&1
tail -3 gpurun_out/ncu_full_10k.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel' -s 3 -c 1 -f -o gpurun_out/prof_pcg_x4 python bench.py --workload bimba_x4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_x4.log 2>&1
tail -3 gpurun_out/ncu_full_x4.log
ls -la gpurun_out/*.ncu-rep
