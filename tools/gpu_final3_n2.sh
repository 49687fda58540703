#!/bin/bash
# round-2 check J: the default bench line under torchrun on 2 GPUs (both arms), as the driver launches it
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee gpurun_out/f3n2_gpus.txt
t0=$(date +%s); timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/f3n2_bench_n2.json 2> gpurun_out/f3n2_bench_n2.err; echo "bench n2 rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/f3n2_bench_n2.err | cut -c1-300
t0=$(date +%s); timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/f3n2_bench_ref_n2.json 2> gpurun_out/f3n2_bench_ref_n2.err; echo "ref n2 rc=$? wall $(( $(date +%s) - t0 )) s"
python - <<PY
import json
for f in ("f3n2_bench_n2","f3n2_bench_ref_n2"):
    try:
        ln=[l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]; d=json.loads(ln)
        print(f, "value", d["value"], "e2e", d.get("e2e",{}).get("value"), "n_gpus", d["n_gpus"])
        for k,v in d.get("workloads",{}).items(): print("  ", k, v.get("value"), v.get("e2e",{}).get("value"))
        print("   batch71", {k:v for k,v in d.get("batch71",{}).items() if k not in ("note",)})
    except Exception as e: print(f, "ERR", e)
PY
