#!/bin/bash
mkdir -p gpurun_out
s=$(date +%s); python bench.py --workload batch71 > gpurun_out/r2t_batch.json 2> gpurun_out/r2t_batch.err; e=$(date +%s); echo "ours: $((e-s)) s"
s=$(date +%s); python bench.py --workload batch71 --impl reference > gpurun_out/r2t_batch_ref.json 2> gpurun_out/r2t_batch_ref.err; e=$(date +%s); echo "reference: $((e-s)) s"
python - <<'PY'
import json
for f in ('gpurun_out/r2t_batch.json','gpurun_out/r2t_batch_ref.json'):
    d=json.loads(open(f).read().strip().split('\n')[-1]); b=d['batch71']
    print(f, d['value'], {k:b.get(k) for k in ('meshes','newton_iters','wall_s','it_per_s','mps','concurrent_processes_per_gpu','concurrent_processes','failed_meshes','one_iteration_process_wall_s','slowest_mesh_rank0')})
PY
ps aux | grep "[n]vidia-cuda-mps" | head -3; tail -3 gpurun_out/r2t_batch.err
