#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/gpu_host_timing.py 2>&1 | tee gpurun_out/host_timing.txt
OCB_PCG_DEBUG=1 python bench.py --workload bimba10k --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bimba10k.json 2> gpurun_out/bench_bimba10k.err
tail -1 gpurun_out/bench_bimba10k.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_bimba10k.json"))
print("10k", "it/s", round(d["value"],3), "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "pcg iters", d["config"]["pcg_iters_mean"])
for k,v in d["kernels"].items():
    print("   ", k, {a:(round(b,5) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms_per_launch","launches","share")})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_10k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_10k.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    a=agg.setdefault(r[ki][:60],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in agg.items(): print("%-62s n=%3d total %10.1f us  avg %8.2f us"%(k,n,t/1e3,t/1e3/n))
PY
