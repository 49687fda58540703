#!/bin/bash
# where the host program's time goes: configs[0] (bimba, lambda 0.999) and configs[1] with the library's host timers + candidate report
mkdir -p gpurun_out /tmp/o1 /tmp/o2
cp tests/golden/inputs/* /tmp/o1/; cp tests/golden/inputs/* /tmp/o2/
cd /tmp/o2; s=$(date +%s.%N); OCB_HOST_TIMING=1 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda 100 /tmp/o2/bimba_i_f10000.obj 0.025 1 2 4.1 1 0 t > /tmp/o2/out.txt 2> /tmp/o2/err.txt; e=$(date +%s.%N)
python3 -c "print('configs[1] process wall %.2f s' % ($e - $s))" > $GRAFT_REPO_ROOT/gpurun_out/r2o_cfg2.txt; grep "ocb " /tmp/o2/err.txt >> $GRAFT_REPO_ROOT/gpurun_out/r2o_cfg2.txt; cat /tmp/o2/output/*/info.txt | head -4 >> $GRAFT_REPO_ROOT/gpurun_out/r2o_cfg2.txt
cd /tmp/o1; s=$(date +%s.%N); OCB_HOST_TIMING=1 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda 100 /tmp/o1/bimba_i_f10000.obj 0.999 1 0 4.1 1 0 t > /tmp/o1/out.txt 2> /tmp/o1/err.txt; e=$(date +%s.%N)
python3 -c "print('configs[0] process wall %.2f s' % ($e - $s))" > $GRAFT_REPO_ROOT/gpurun_out/r2o_cfg1.txt; grep "ocb " /tmp/o1/err.txt >> $GRAFT_REPO_ROOT/gpurun_out/r2o_cfg1.txt; cat /tmp/o1/output/*/info.txt | head -4 >> $GRAFT_REPO_ROOT/gpurun_out/r2o_cfg1.txt
cd $GRAFT_REPO_ROOT; cat gpurun_out/r2o_cfg2.txt gpurun_out/r2o_cfg1.txt | cut -c1-200
python -m pytest tests/test_gpu_dropin.py -q -m gpu 2>&1 | tail -3
