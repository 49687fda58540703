#!/bin/bash
# host-side timers of the C-ABI inside the reference host program (drop-in binary), 30 Newton iterations
mkdir -p gpurun_out /tmp/dropin && cd /tmp/dropin
python - <<PY
import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from objfixture import write_obj
g = np.load("/root/repo/tests/golden/bimba_cfg2_states.npz")
write_obj("/tmp/dropin/bimba_s1.obj", g["s1_V_rest"], g["s1_F"], g["s1_V"])
PY
OCB_HOST_TIMING=1 ORACLE_TRACE=/tmp/dropin/trace.txt ORACLE_MAX_ITERS=30 /root/repo/shim/_build/OptCuts_cuda_probe 100 /tmp/dropin/bimba_s1.obj 0.025 1 2 4.1 1 0 t > /tmp/dropin/out.txt 2> /tmp/dropin/err.txt
grep "ocb host" /tmp/dropin/err.txt | sort -k6 -n -r | head -30
tail -3 /tmp/dropin/out.txt
