#!/bin/bash
mkdir -p gpurun_out /tmp/x; cd /tmp/x; tar xf $GRAFT_REPO_ROOT/tests/golden/inputs/benchmark71.tar.xz 2>/dev/null
OUT=$GRAFT_REPO_ROOT/gpurun_out/r2y_cat.txt; : > $OUT
f=$(find /tmp/x -name "cat_noUV.obj" | head -1)
run() { tag=$1; shift; mkdir -p /tmp/x/c; cd /tmp/x/c; rm -rf output trace.txt
  s=$(date +%s.%N); env "$@" ORACLE_MAX_ITERS=30 ORACLE_TRACE=/tmp/x/c/trace.txt OCB_PCG_DEBUG=1 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda_probe 100 $f 0.999 1 0 4.1 1 0 b > /dev/null 2> err.txt; e=$(date +%s.%N)
  python3 -c "print('== $tag: wall %.2f s, E after 30 iterations %s' % ($e - $s, open('trace.txt').read().strip().split('\n')[-1].split()[11]))" >> $OUT
  echo "   rejections $(grep -c rejected err.txt)" >> $OUT; grep rejected err.txt | head -4 | cut -c1-200 >> $OUT
  grep "iters" err.txt | awk '{for(i=1;i<=NF;i++) if($i=="iters") s+=$(i+1); n++} END {print "   CG iterations total", s, "in", n, "solves"}' >> $OUT; }
run default X=1
run "system scaling" OCB_SCALE_SYSTEM=1
run "no equilibration" OCB_MAS_EQUILIBRATE=0
run "no MAS (block-Jacobi)" OCB_NO_MAS=1
run "coarse max 60" OCB_MAS_COARSE_MAX=60
cd /tmp/x/c; ORACLE_MAX_ITERS=30 ORACLE_TRACE=/tmp/x/c/trace_ref.txt $GRAFT_REPO_ROOT/oracle/_ref/OptCuts_probe 100 $f 0.999 1 0 4.1 1 0 b > /dev/null 2>&1; echo "== reference: E after 30 iterations $(tail -1 trace_ref.txt | cut -d' ' -f12)" >> $OUT
awk '{print $1, $12}' /tmp/x/c/trace_ref.txt | head -12 >> $OUT
cat $OUT
