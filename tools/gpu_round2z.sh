#!/bin/bash
# dense-Cholesky safety net: the hard meshes again (150 iterations, both arms), then the tests
mkdir -p gpurun_out /tmp/x; cd /tmp/x; tar xf $GRAFT_REPO_ROOT/tests/golden/inputs/benchmark71.tar.xz 2>/dev/null
OUT=$GRAFT_REPO_ROOT/gpurun_out/r2z_slow_meshes.txt; : > $OUT
for m in cat_noUV torusOnPlane male_2_f20000; do
  f=$(find /tmp/x -name "$m.obj" | head -1)
  for arm in cuda ref; do
    exe=$GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda_probe; [ $arm = ref ] && exe=$GRAFT_REPO_ROOT/oracle/_ref/OptCuts_probe
    mkdir -p /tmp/x/w_${m}_$arm; cd /tmp/x/w_${m}_$arm; rm -rf output trace.txt
    s=$(date +%s.%N); ORACLE_MAX_ITERS=150 ORACLE_TRACE=trace.txt OCB_HOST_TIMING=1 OCB_PCG_DEBUG=1 timeout 300 $exe 100 $f 0.999 1 0 4.1 1 0 b > /dev/null 2> err.txt; e=$(date +%s.%N)
    python3 -c "
t=open('trace.txt').read().strip().split('\n')
g=lambda i: t[i].split()[11] if i < len(t) else '-'
print('== $m $arm: process wall %.2f s, %d iterations, %s; E after 10 / 30 / last: %s %s %s' % ($e - $s, len(t), t[-1].split()[2], g(9), g(29), g(len(t)-1)))" >> $OUT
    if [ $arm = cuda ]; then echo "   direct solves: $(grep -c "CG gave up" err.txt), block-Jacobi retries: $(grep -c 'repeating with block-Jacobi' err.txt)" >> $OUT; grep -E "ocb host\] (newton_step|solve) " err.txt | sed 's/^/   /' >> $OUT; grep "ocb direct" err.txt | head -2 | cut -c1-160 | sed 's/^/   /' >> $OUT; fi
  done
done
cut -c1-250 $OUT
cd $GRAFT_REPO_ROOT; python -m pytest tests -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log; tail -4 gpurun_out/r2z_pytest.log
