#!/bin/bash
# round-2 check H: the 9 formerly failing benchmark meshes, GPU tests, batch leg alone
mkdir -p gpurun_out
W=/tmp/failing; rm -rf $W; mkdir -p $W/in; tar -xJf tests/golden/inputs/benchmark71.tar.xz -C $W/in
for m in male_2_f20000 cat_noUV dragon_i_f10000 hand_yaron_f10000 horse_f10000 armadillo_i_f10000 santa_i_f10000 triceratops_param_closed hand_1_i_f5000; do
  mkdir -p $W/$m $W/ref_$m
  ( cd $W/$m && ORACLE_MAX_ITERS=40 ORACLE_TRACE=$W/$m/trace.txt timeout 300 $OLDPWD/shim/_build/OptCuts_cuda_probe 100 $W/in/$m.obj 0.999 1 0 4.1 1 0 b > out.txt 2> err.txt; echo "== $m rc=$? iterations $(wc -l < trace.txt 2>/dev/null) last: $(tail -1 trace.txt | cut -d' ' -f1-3,12-14)"; tail -2 err.txt | cut -c1-200 )
  ( cd $W/ref_$m && ORACLE_MAX_ITERS=40 ORACLE_TRACE=$W/ref_$m/trace.txt timeout 300 $OLDPWD/oracle/_ref/OptCuts_probe 100 $W/in/$m.obj 0.999 1 0 4.1 1 0 b > out.txt 2> err.txt; echo "   ref rc=$? iterations $(wc -l < trace.txt 2>/dev/null) last: $(tail -1 trace.txt | cut -d' ' -f1-3,12-14)" )
done 2>&1 | tee gpurun_out/r2h_failing.txt
timeout 3200 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -12 gpurun_out/r2h_pytest.log | cut -c1-300
timeout 900 python bench.py --workload batch71 > gpurun_out/r2h_batch.json 2> gpurun_out/r2h_batch.err; cut -c1-1500 gpurun_out/r2h_batch.json
