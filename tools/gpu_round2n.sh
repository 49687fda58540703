#!/bin/bash
# the default bench lines of both arms with their wall clock (the driver's run), then the 71 meshes to convergence on the GPU arm
mkdir -p gpurun_out
t0=$(date +%s.%N); python bench.py --impl reference > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err; t1=$(date +%s.%N)
python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; t2=$(date +%s.%N)
echo "bench.py --impl reference: $(echo "$t1 - $t0" | bc) s ; bench.py: $(echo "$t2 - $t1" | bc) s" | tee gpurun_out/r2n_bench_wall.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
print(json.dumps(d.get('host_program'))[:2500])
print(json.dumps(d.get('batch71'))[:700])
r=json.loads(open('gpurun_out/r2n_bench_ref.json').read().strip().split('\n')[-1])
print({k:r.get(k) for k in ('value','ms_per_step')}, json.dumps(r.get('batch71'))[:500])
PY
python tools/batch_to_convergence.py cuda gpurun_out/r2n_b71_cuda.json 3 600 2>&1 | tail -3
