#!/bin/bash
# process start-up cost of the host program with and without another process holding the GPU open (what nvidia-persistenced does)
mkdir -p gpurun_out /tmp/q; cp tests/golden/inputs/* /tmp/q/
run() { cd /tmp/q; rm -rf output; s=$(date +%s.%N); OCB_HOST_TIMING=1 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda 100 /tmp/q/bimba_i_f10000.obj 0.025 1 2 4.1 1 0 t > /dev/null 2> /tmp/q/err.txt; e=$(date +%s.%N); python3 -c "print('$1: process wall %.2f s' % ($e - $s))"; grep -E "process age|cuda context" /tmp/q/err.txt | tr '\n' ';'; echo; sed -n 3p /tmp/q/output/*/info.txt | cut -c1-60; }
nvidia-smi --query-gpu=persistence_mode --format=csv
run "alone (1st)"; run "alone (2nd)"
python -c "
import torch, time, sys
x = torch.zeros(1, device='cuda'); torch.cuda.synchronize(); print('keeper up', flush=True); time.sleep(60)" &
sleep 12
run "with a keeper process (1st)"; run "with a keeper process (2nd)"
kill %1
