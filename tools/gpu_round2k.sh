#!/bin/bash
# which change moved the free run of configs[1] off the reference's op sequence?  (library variants under tools/_variants/)
mkdir -p gpurun_out
V="cur:-:,cur_nohalo:-:OCB_PCG_NO_HALO=1,head:tools/_variants/head:,head_scaled:tools/_variants/head:OCB_SCALE_SYSTEM=1,cur_nocluster:-:OCB_PCG_NO_CLUSTER=1"
python tools/gpu_diag_run.py bimba_cfg2 "$V" > gpurun_out/r2k_diag.txt 2>&1
python tools/gpu_diag_run.py torus_cfg1 "cur:-:,head:tools/_variants/head:" >> gpurun_out/r2k_diag.txt 2>&1
cat gpurun_out/r2k_diag.txt | cut -c1-900
python -m pytest tests -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -8 gpurun_out/r2k_pytest.log
