#!/bin/bash
# round-2 check F: GPU tests; the default bench line (headline + x4 + x10 + batch71 + host program), the reference arm;
# fixed overhead of one host-program process; (optional) the reference's iteration at 1M faces on this box
mkdir -p gpurun_out
timeout 3200 python -m pytest tests -q -m gpu -x --durations=10 -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log; tail -40 gpurun_out/r2f_pytest.log | cut -c1-300
t0=$(date +%s); timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r2f_bench.err
t0=$(date +%s); timeout 1500 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r2f_bench_ref.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print("10k", round(d["value"],1), "it/s e2e", round(d["e2e"]["value"],1), "cpu", d.get("cpu_baseline"))
for k,v in d.get("workloads",{}).items(): print(k, round(v["value"],2), "e2e", round(v["e2e"]["value"],2), "cpu", v.get("cpu_baseline"))
print("batch71", {k:v for k,v in d.get("batch71",{}).items() if k!="note"})
print("host_program", d.get("host_program"))
r=json.load(open("gpurun_out/r2f_bench_ref.json"))
print("ref 10k", r["value"], "x4", r.get("workloads",{}).get("bimba_x4",{}).get("value"), "batch", {k:v for k,v in r.get("batch71",{}).items() if k!="note"})
PY
# fixed cost of one host-program process: start-up + 1 Newton iteration + exit
W=/tmp/fixed; rm -rf $W; mkdir -p $W; cp tests/golden/inputs/bimba_i_f10000.obj $W/
for exe in shim/_build/OptCuts_cuda_probe oracle/_ref/OptCuts_probe; do
  ( cd $W && s=$(date +%s.%N) && OCB_HOST_TIMING=1 ORACLE_MAX_ITERS=1 $OLDPWD/$exe 100 $W/bimba_i_f10000.obj 0.025 1 2 4.1 1 0 t > out.txt 2> err.txt; e=$(date +%s.%N); python3 -c "print('$exe one-iteration process wall', round($e-$s,3), 's')"; grep "set_mesh\|ocb candidates" err.txt | head -3 )
done 2>&1 | tee gpurun_out/r2f_fixed_overhead.txt
if [ "$RUN_REF_X10" = "1" ]; then
  timeout 1500 python bench.py --impl reference --workload bimba_x10 --steps 1 --warmup 0 > gpurun_out/r2f_ref_bimba_x10.json 2>/dev/null; cat gpurun_out/r2f_ref_bimba_x10.json | cut -c1-300
fi
