#!/bin/bash
# tests + whole runs of the remaining traced configs + the default bench lines of both arms (with their wall clock)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -5 gpurun_out/r2m_pytest.log
python tools/gpu_diag_run.py lucy6k_cfg1 "cur:-:" > gpurun_out/r2m_diag.txt 2>&1
python tools/gpu_diag_run.py bimba_cfg1 "cur:-:" >> gpurun_out/r2m_diag.txt 2>&1
python tools/gpu_diag_run.py face_rsp_cfg1 "cur:-:" >> gpurun_out/r2m_diag.txt 2>&1
cut -c1-400 gpurun_out/r2m_diag.txt
/usr/bin/time -v python bench.py --impl reference > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2m_bench_ref.err
/usr/bin/time -v python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().split('\n')[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')})
print(json.dumps(d.get('host_program'))[:1800])
print(json.dumps(d.get('batch71'))[:600])
PY
