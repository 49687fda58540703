#!/bin/bash
mkdir -p gpurun_out /tmp/o1
cp tests/golden/inputs/* /tmp/o1/
cd /tmp/o1; OCB_CANDIDATES_REPORT=1 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda 100 /tmp/o1/bimba_i_f10000.obj 0.999 1 0 4.1 1 0 t > /tmp/o1/out.txt 2> /tmp/o1/err.txt
grep "ocb " /tmp/o1/err.txt | cut -c1-400; cat /tmp/o1/output/*/info.txt | head -4
cd $GRAFT_REPO_ROOT; python tools/gpu_diag_run.py bimba_cfg1 "cur:-:" 2>&1 | cut -c1-300
python tools/gpu_diag_run.py bimba_cfg2 "cur:-:" 2>&1 | cut -c1-300
python tools/host_program_timing.py 2>&1 | cut -c1-500 | tail -16
