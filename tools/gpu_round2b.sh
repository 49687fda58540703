#!/bin/bash
# round-2 check B: GPU tests incl. whole runs of the host program with the Optimizer hooks + teacher-forced sweeps;
# whole-run timing of the drop-in host program against the unmodified reference (same box, same inputs)
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -30 gpurun_out/r2b_pytest.log
W=/tmp/hostrun; rm -rf $W; mkdir -p $W/cuda $W/ref $W/cuda0
cp tests/golden/inputs/* $W/cuda/; cp tests/golden/inputs/* $W/ref/; cp tests/golden/inputs/* $W/cuda0/
ARGS="0.025 1 2 4.1 1 0 t"
( cd $W/cuda && /usr/bin/time -v $OLDPWD/shim/_build/OptCuts_cuda 100 $W/cuda/bimba_i_f10000.obj $ARGS > out.txt 2> err.txt )
( cd $W/cuda0 && OCB_DEVICE_NEWTON=0 /usr/bin/time -v $OLDPWD/shim/_build/OptCuts_cuda 100 $W/cuda0/bimba_i_f10000.obj $ARGS > out.txt 2> err.txt )
( cd $W/ref && /usr/bin/time -v $OLDPWD/oracle/_ref/OptCuts_bin 100 $W/ref/bimba_i_f10000.obj $ARGS > out.txt 2> err.txt )
{
  echo "== whole run, bimba_i_f10000 configs[1] ($ARGS), $(nproc) host cores: info.txt line 2 (iterations) and 3 (timers)"
  for d in cuda cuda0 ref; do
    echo "-- $d  (cuda = Optimizer hooks, device-resident Newton; cuda0 = OCB_DEVICE_NEWTON=0, call-by-call plugins; ref = unmodified reference)"
    sed -n 2,4p $W/$d/output/*/info.txt
    grep "Elapsed (wall clock)" $W/$d/err.txt
  done
} > gpurun_out/r2b_host_program.txt 2>&1
cat gpurun_out/r2b_host_program.txt
