#!/bin/bash
# round-2 check G: why 9 of the 71 benchmark meshes fail; process fixed cost; GPU tests; bench quick line
mkdir -p gpurun_out
for i in 1 2; do ./tools/micro/init_bench $PWD/optcuts_b200/lib/liboptcuts_b200.so; done 2>&1 | tee gpurun_out/r2g_init.txt
W=/tmp/failing; rm -rf $W; mkdir -p $W/in; tar -xJf tests/golden/inputs/benchmark71.tar.xz -C $W/in
for m in male_2_f20000 cat_noUV dragon_i_f10000 hand_yaron_f10000 horse_f10000 armadillo_i_f10000 santa_i_f10000 triceratops_param_closed hand_1_i_f5000; do
  mkdir -p $W/$m; ( cd $W/$m && ORACLE_MAX_ITERS=15 ORACLE_TRACE=$W/$m/trace.txt timeout 300 $OLDPWD/shim/_build/OptCuts_cuda_probe 100 $W/in/$m.obj 0.999 1 0 4.1 1 0 b > out.txt 2> err.txt; echo "== $m rc=$? iterations $(wc -l < trace.txt 2>/dev/null)"; tail -4 err.txt | cut -c1-300; tail -2 out.txt | cut -c1-200 )
done 2>&1 | tee gpurun_out/r2g_failing.txt
timeout 3200 python -m pytest tests -q -m gpu -x --durations=5 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -12 gpurun_out/r2g_pytest.log | cut -c1-300
timeout 600 python bench.py --quick --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2g_bench.json")); print("10k", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k:round(v["ms_per_launch"]*1000,1) for k,v in d["kernels"].items()})
PY
