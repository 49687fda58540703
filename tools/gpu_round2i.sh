#!/bin/bash
# round-2 check I: symmetric system scaling on/off: CG iterations + step time at the three sizes; the 9 hard benchmark meshes; tests
mkdir -p gpurun_out
W=/tmp/failing; rm -rf $W; mkdir -p $W/in; tar -xJf tests/golden/inputs/benchmark71.tar.xz -C $W/in
for m in male_2_f20000 cat_noUV dragon_i_f10000 hand_yaron_f10000 horse_f10000 armadillo_i_f10000 santa_i_f10000 triceratops_param_closed hand_1_i_f5000; do
  mkdir -p $W/$m
  ( cd $W/$m && s=$(date +%s.%N) && ORACLE_MAX_ITERS=40 ORACLE_TRACE=$W/$m/trace.txt timeout 300 $OLDPWD/shim/_build/OptCuts_cuda_probe 100 $W/in/$m.obj 0.999 1 0 4.1 1 0 b > out.txt 2> err.txt; rc=$?; e=$(date +%s.%N); echo "== $m rc=$rc wall $(python3 -c "print(round($e-$s,1))") s iterations $(wc -l < trace.txt 2>/dev/null) last: $(tail -1 trace.txt | cut -d' ' -f1,12-13)"; echo "   truncated: $(grep -c truncated out.txt) cap: $(grep -c "iteration cap" out.txt) lifted: $(grep -c lifted out.txt)" )
done 2>&1 | tee gpurun_out/r2i_hard_meshes.txt
timeout 3200 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -12 gpurun_out/r2i_pytest.log | cut -c1-300
