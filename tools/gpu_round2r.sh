#!/bin/bash
# after the pooled parallel_for: whole runs of configs[0] / configs[1] in the host program (timers + candidate report), reference configs[0], tests
mkdir -p gpurun_out /tmp/o1 /tmp/o2 /tmp/o3
for d in o1 o2 o3; do cp tests/golden/inputs/* /tmp/$d/; done
OUT=$GRAFT_REPO_ROOT/gpurun_out/r2r_host_program.txt; : > $OUT
one() { # tag dir exe args...
  tag=$1; dir=$2; exe=$3; shift 3
  cd $dir; rm -rf output; s=$(date +%s.%N); OCB_CANDIDATES_REPORT=1 $exe 100 $dir/bimba_i_f10000.obj "$@" t > /dev/null 2> $dir/err.txt; e=$(date +%s.%N)
  python3 -c "print('== $tag: process wall %.2f s' % ($e - $s))" >> $OUT; grep "ocb " $dir/err.txt >> $OUT; sed -n 2,4p $dir/output/*/info.txt >> $OUT
}
one "configs[1] GPU host program" /tmp/o2 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda 0.025 1 2 4.1 1 0
one "configs[0] GPU host program" /tmp/o1 $GRAFT_REPO_ROOT/shim/_build/OptCuts_cuda 0.999 1 0 4.1 1 0
one "configs[0] reference" /tmp/o3 $GRAFT_REPO_ROOT/oracle/_ref/OptCuts_bin 0.999 1 0 4.1 1 0
one "configs[1] reference" /tmp/o3 $GRAFT_REPO_ROOT/oracle/_ref/OptCuts_bin 0.025 1 2 4.1 1 0
cd $GRAFT_REPO_ROOT; cut -c1-420 $OUT

