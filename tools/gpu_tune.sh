#!/bin/bash
python -m pytest tests -q -m gpu -k "pcg or newton or linsys or dropin" 2>&1 | tail -2
OCB_PCG_DEBUG=1 python bench.py --workload bimba10k --steps 4 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | tail -2
