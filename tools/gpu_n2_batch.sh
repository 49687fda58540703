#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s); timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload batch71 > gpurun_out/f3n2_batch_omp.json 2> gpurun_out/f3n2_batch.err; echo "rc=$? wall $(( $(date +%s) - t0 )) s"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/f3n2_batch_omp.json") if l.startswith("{")][-1]); b=d["batch71"]
print({k:b.get(k) for k in ("newton_iters","wall_s","it_per_s","per_rank_s","omp_threads_per_process","mps","slowest_mesh_rank0","one_iteration_process_wall_s")})
PY
nproc
