python -m pytest tests/test_gpu_mas.py -q -m gpu -x 2>&1 | tail -1
OCB_MAS_DEBUG=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep "ocb mas" | tail -1
for w in bimba10k bimba_x4; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['value'],2), 'it/s', round(d['ms_per_step'],3), 'ms; mas_setup us', round(d['kernels']['mas_setup']['ms_per_launch']*1000,1), 'iters', d['config']['pcg_iters_mean'])"; done
OCB_MAS_DEBUG=1 python bench.py --workload bimba_x4 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "ocb mas" | tail -1
