#!/bin/bash
# where do the slow meshes of the bounded batch spend their 150 iterations?
mkdir -p gpurun_out /tmp/x; cd /tmp/x; tar xf $GRAFT_REPO_ROOT/tests/golden/inputs/benchmark71.tar.xz 2>/dev/null; ls | head -3
OUT=$GRAFT_REPO_ROOT/gpurun_out/r2x_slow_meshes.txt; : > $OUT
for m in torusOnPlane cat_noUV vase_lion_o_f20000 bimba_i_f10000; do
  f=$(find /tmp/x -name "$m.obj" | head -1)
  for arm in cuda ref; do
    exe=$GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda_probe; [ $arm = ref ] && exe=$GRAFT_REPO_ROOT/oracle/_ref/OptCuts_probe
    mkdir -p /tmp/x/w_$m_$arm; cd /tmp/x/w_$m_$arm; rm -rf output
    s=$(date +%s.%N); ORACLE_MAX_ITERS=150 ORACLE_TRACE=/tmp/x/w_$m_$arm/trace.txt OCB_HOST_TIMING=1 OCB_PCG_DEBUG=1 $exe 100 $f 0.999 1 0 4.1 1 0 b > /dev/null 2> err.txt; e=$(date +%s.%N)
    python3 -c "print('== $m $arm: process wall %.2f s, %d iterations, topology steps %s' % ($e - $s, sum(1 for _ in open('trace.txt')), open('trace.txt').read().strip().split('\n')[-1].split()[2]))" >> $OUT
    if [ $arm = cuda ]; then echo "   rejections of the two-level preconditioner: $(grep -c rejected err.txt)" >> $OUT; grep -E "ocb host\] (newton_step|solve|set_mesh|set_pattern_from|factorize|hessian)|ocb candidates" err.txt | sed 's/^/   /' >> $OUT; grep "iters" err.txt | awk '{for(i=1;i<=NF;i++) if($i=="iters") s+=$(i+1); n++} END {print "   CG iterations total", s, "in", n, "solves"}' >> $OUT; fi
  done
done
cut -c1-260 $OUT
