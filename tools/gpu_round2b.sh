#!/bin/bash
# round-2 check B: GPU tests incl. whole runs of the host program with the Optimizer hooks + teacher-forced sweeps;
# whole-run timing of the drop-in host program against the unmodified reference (same box, same inputs)
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -30 gpurun_out/r2b_pytest.log
python tools/host_program_timing.py > gpurun_out/r2b_host_program.txt 2>&1
cat gpurun_out/r2b_host_program.txt
