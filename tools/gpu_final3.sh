#!/bin/bash
# round-2 FINAL evidence (after the robust-solve work): tests, smoke, both bench arms (default lines), launch list, ncu --set full of the dominant kernels (traffic),
# PCG / dense-inverse phase cycles, host timing, sanitizer logs.  tools/make_profiles2.py turns gpurun_out/ into profiles/r2_*.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/f3_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f3_pytest_gpu.log; tail -3 gpurun_out/f3_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f3_smoke.log 2>&1; tail -1 gpurun_out/f3_smoke.log
t0=$(date +%s); python bench.py --impl reference > gpurun_out/f3_bench_ref.json 2>/dev/null; t1=$(date +%s)
python bench.py > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err; t2=$(date +%s)
echo "wall clock of the default bench runs: python bench.py --impl reference $((t1-t0)) s, python bench.py $((t2-t1)) s" | tee gpurun_out/f3_bench_wall.txt
for w in bimba10k bimba_x4 bimba_x10; do
  OCB_PCG_DEBUG=1 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "ocb pcg" | tail -1
done > gpurun_out/f3_pcg_phase_cycles.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f3_launches_bimba10k.csv python bench.py --workload bimba10k --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel' -s 3 -c 1 -f -o gpurun_out/f3_prof_pcg10k python bench.py --workload bimba10k --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pcg_kernel' -s 3 -c 1 -f -o gpurun_out/f3_prof_pcg_x10 python bench.py --workload bimba_x10 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hessian_elem_kernel|hessian_rows_kernel|energy_kernel|grad_gather_kernel|step_bound|mas_dense_invert|mas_galerkin|mas_coarsen' -s 0 -c 12 -f -o gpurun_out/f3_prof_elem_x10 python bench.py --workload bimba_x10 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/gpu_host_timing.py > gpurun_out/f3_host_timing.txt 2>&1
(OCB_MAS_DEBUG=1 python bench.py --workload bimba10k --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep 'ocb mas' | tail -1; OCB_MAS_DEBUG=1 python bench.py --workload bimba_x4 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep 'ocb mas' | tail -1) > gpurun_out/f3_mas_dense_phases.txt 2>&1
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py > gpurun_out/f3_sanitizer_$tool.log 2>&1; echo "== $tool rc=$?"; tail -2 gpurun_out/f3_sanitizer_$tool.log
done
OCB_HOST_TIMING=1 python tools/host_program_timing.py > gpurun_out/f3_host_program.txt 2>&1; tail -12 gpurun_out/f3_host_program.txt
ls -la gpurun_out | grep f3_
python tools/batch_to_convergence.py ref gpurun_out/f3_b71_ref_sample.json 3 900 -8 2>&1 | tail -2
bash tools/gpu_round2r.sh > /dev/null 2>&1; cp gpurun_out/r2r_host_program.txt gpurun_out/f3_host_program_cfg0_cfg1.txt; cut -c1-300 gpurun_out/f3_host_program_cfg0_cfg1.txt | grep -v "ocb host"
