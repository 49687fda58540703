#!/bin/bash
# equilibrated preconditioner on by default: tests, whole runs against the traces, the bounded batch
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log; tail -6 gpurun_out/r2w_pytest.log
for n in bimba_cfg2 torus_cfg1 bimba_cfg1 face_rsp_cfg1; do python tools/gpu_diag_run.py $n "cur:-:" 2>&1 | cut -c1-330; done | tee gpurun_out/r2w_diag.txt
python bench.py --workload batch71 > gpurun_out/r2w_batch.json 2> gpurun_out/r2w_batch.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_batch.json').read().strip().split('\n')[-1]); b=d['batch71']
print(d['value'], {k:b.get(k) for k in ('meshes','newton_iters','wall_s','it_per_s','mps','failed_meshes','slowest_mesh_rank0')})
PY
