#!/bin/bash
python -m pytest tests -q -m gpu -k "pcg or newton or linsys or spd or large" 2>&1 | tail -3
for env in "OCB_PCG_NO_CLUSTER=0" "OCB_PCG_NO_SMEM=1"; do
echo "== $env"
env $env OCB_PCG_DEBUG=1 python bench.py --workload bimba10k --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -1
done
echo "== x4"; OCB_PCG_DEBUG=1 python bench.py --workload bimba_x4 --steps 1 --warmup 3 --pcg-max-it 3000 --no-cpu-baseline 2>&1 >/dev/null | tail -1
echo "== x4 nosmem"; OCB_PCG_NO_SMEM=1 OCB_PCG_DEBUG=1 python bench.py --workload bimba_x4 --steps 1 --warmup 3 --pcg-max-it 3000 --no-cpu-baseline 2>&1 >/dev/null | tail -1
echo "== x10"; OCB_PCG_DEBUG=1 python bench.py --workload bimba_x10 --steps 1 --warmup 3 --pcg-max-it 1000 --no-cpu-baseline 2>&1 >/dev/null | tail -1
