#!/bin/bash
# can an MPS daemon take the CUDA context creation out of the per-mesh processes?
mkdir -p gpurun_out /tmp/q /tmp/mps /tmp/mpslog; cp tests/golden/inputs/* /tmp/q/
run() { cd /tmp/q; rm -rf output; s=$(date +%s.%N); OCB_HOST_TIMING=1 timeout 120 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda 100 /tmp/q/bimba_i_f10000.obj 0.025 1 2 4.1 1 0 t > /dev/null 2> /tmp/q/err.txt; rc=$?; e=$(date +%s.%N); python3 -c "print('$1: rc $rc process wall %.2f s' % ($e - $s))"; grep -E "cuda context|age at the report" /tmp/q/err.txt | tr '\n' ';'; echo; sed -n 2,4p /tmp/q/output/*/info.txt | cut -c1-70 | tr '\n' '|'; echo; }
which nvidia-cuda-mps-control nvidia-cuda-mps-server; nvidia-smi --query-gpu=compute_mode --format=csv
run "no MPS"
export CUDA_MPS_PIPE_DIRECTORY=/tmp/mps CUDA_MPS_LOG_DIRECTORY=/tmp/mpslog
timeout 20 nvidia-cuda-mps-control -d; echo "mps-control -d rc=$?"
sleep 1
run "MPS (1st client starts the server)"; run "MPS (2nd)"; run "MPS (3rd)"
( run "MPS concurrent A" & run "MPS concurrent B" & run "MPS concurrent C" & wait )
echo quit | timeout 20 nvidia-cuda-mps-control; echo "mps quit rc=$?"
sleep 1; ps aux | grep -c "[n]vidia-cuda-mps"
tail -5 /tmp/mpslog/control.log 2>/dev/null
unset CUDA_MPS_PIPE_DIRECTORY CUDA_MPS_LOG_DIRECTORY
run "no MPS again"
