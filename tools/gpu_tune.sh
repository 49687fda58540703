#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -k "pcg or newton or linsys or spd" 2>&1 | tail -3
for w in bimba10k bimba_x4; do
for rows in 64 128 256 400 550; do
OCB_PCG_ROWS_PER_CTA=$rows python bench.py --workload $w --steps 4 --warmup 3 --pcg-max-it 60000 --no-cpu-baseline > gpurun_out/t.json 2> gpurun_out/t.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/t.json"))
    k=d["kernels"]["pcg"]
    print("$w rows/cta $rows", "ms/step", round(d["ms_per_step"],3), "pcg ms", round(k["ms_per_launch"],3), "iters", d["config"]["pcg_iters_mean"], "us/iter", round(1e3*k["ms_per_launch"]/d["config"]["pcg_iters_mean"],2), "E", d["E_new"][0])
except Exception as e:
    print("$w $rows failed", e, open("gpurun_out/t.err").read()[-300:])
PY
done
done
OCB_PCG_NO_SMEM=1 python bench.py --workload bimba_x10 --steps 2 --warmup 3 --pcg-max-it 60000 --no-cpu-baseline | python -c "
import json,sys
d=json.load(sys.stdin); k=d['kernels']['pcg']; print('x10 ms/step', d['ms_per_step'], 'us/iter', 1e3*k['ms_per_launch']/d['config']['pcg_iters_mean'])"
