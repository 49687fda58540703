#!/bin/bash
# round-2 check D: GPU tests incl. device candidate evaluation + whole runs; whole-run timing
mkdir -p gpurun_out
timeout 3200 python -m pytest tests -q -m gpu -x --durations=10 -s > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log; tail -40 gpurun_out/r2d_pytest.log | cut -c1-400
OCB_HOST_TIMING=1 python tools/host_program_timing.py > gpurun_out/r2d_host_program.txt 2>&1; tail -14 gpurun_out/r2d_host_program.txt
