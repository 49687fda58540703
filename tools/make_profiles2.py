"""Turn the raw captures a GPU evidence run left in gpurun_out/ into the committed summaries under profiles/:
    python tools/make_profiles2.py r2
  * bench lines (copied), PCG phase cycles, host timing, element micro-benchmark (copied)
  * <tag>_launches_bimba10k.csv + a per-kernel summary (launches, total, average, share) of the ncu launch list
  * <tag>_ncu_full_summary.txt: DRAM bytes, duration, occupancy, registers, L2 hit rate, fp64 pipe of every kernel in
    the ncu --set full captures, and profiles/traffic.json (DRAM bytes per launch, read by bench.py)
  * <tag>_ncu_<capture>_lines.txt: warp-stall samples per source line (tools/ncu_lines.py)
ncu here only READS the reports (no GPU needed)."""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
PFX = sys.argv[2] if len(sys.argv) > 2 else "f3_"

for f in ("bench", "bench_ref", "bench_wall"):
    for ext in (".json", ".txt"):
        src = os.path.join(OUT, PFX + f + ext)
        if os.path.exists(src) and os.path.getsize(src):
            shutil.copy(src, os.path.join(PROF, "%s_%s%s" % (tag, f, ext)))
for f, dst in (("pcg_phase_cycles.txt", "pcg_phase_cycles.txt"), ("host_timing.txt", "host_timing_bimba10k.txt"), ("mas_dense_phases.txt", "mas_dense_phases.txt"),
               ("host_program.txt", "host_program.txt"), ("smoke.log", "smoke.log"), ("sanitizer_memcheck.log", "sanitizer_memcheck.log"),
               ("sanitizer_racecheck.log", "sanitizer_racecheck.log"), ("sanitizer_initcheck.log", "sanitizer_initcheck.log"), ("sanitizer_synccheck.log", "sanitizer_synccheck.log")):
    src = os.path.join(OUT, PFX + f)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, "%s_%s" % (tag, dst)))

# ---- launch list
src = os.path.join(OUT, PFX + "launches_bimba10k.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(PROF, tag + "_launches_bimba10k.csv"))
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    h = rows[0]; ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        a = agg.setdefault(r[ik], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PROF, tag + "_launches_bimba10k_summary.txt"), "w") as f:
        f.write("kernel | launches | total us | avg us | share   (ncu --metrics gpu__time_duration.sum --clock-control none; python bench.py --steps 2 --warmup 3)\n")
        for k, (n, t) in agg.items():
            f.write("%-52s | %3d | %10.1f | %8.2f | %.3f\n" % (k[:50], n, t, t / n, t / tot))

# ---- full captures
METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
CLASS = (("pcg_kernel", "pcg"), ("grad_gather_kernel", "gradient"), ("hessian_elem_kernel", "hessian_psd_scatter"), ("hessian_rows_kernel", "hessian_rows"),
         ("mas_dense_invert", "mas_dense_invert"), ("step_bound", "step_bound"), ("energy_kernel", "energy"))
traffic = {}
lines = ["workload | kernel | dram read B | dram write B | duration us | warps active % | regs | L2 hit % | fp64 pipe % | compute-memory throughput %"]
for rep, wl in (("prof_pcg10k", "bimba10k"), ("prof_pcg_x10", "bimba_x10"), ("prof_elem_x10", "bimba_x10")):
    path = os.path.join(OUT, PFX + rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    col = {m: h.index(m) for m in METRICS if m in h}
    per = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        def val(m, scale=None):
            if m not in col or r[col[m]] in ("", "n/a"): return float("nan")
            v = float(r[col[m]].replace(",", ""))
            u = units[col[m]]
            return v * (UNIT.get(u, 1.0) if scale == "b" else TIME.get(u, 1.0) if scale == "t" else 1.0)
        rd, wr, du = val(METRICS[0], "b"), val(METRICS[1], "b"), val(METRICS[2], "t")
        lines.append("%s | %s | %.0f | %.0f | %.2f | %.1f | %.0f | %.1f | %.1f | %.1f" % (wl, name[:44], rd, wr, du, val(METRICS[3]), val(METRICS[4]), val(METRICS[5]), val(METRICS[6]), val(METRICS[7])))
        for key, cls in CLASS:
            if key in name:
                per[cls].append(rd + wr)
    for cls, v in per.items():
        traffic.setdefault(wl, {})[cls] = sum(v) / len(v)
    src_csv = os.path.join("/tmp", rep + "_src.csv")
    with open(src_csv, "w") as f:
        subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-count", "1"] +
                       (["--kernel-name", "regex:hessian_elem_kernel"] if rep == "prof_elem_x10" else []), stdout=f, stderr=subprocess.DEVNULL)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), src_csv, "40"], capture_output=True, text=True).stdout
    if "total samples" in out:
        open(os.path.join(PROF, "%s_ncu_%s_lines.txt" % (tag, rep.replace("prof_", "").replace("elem_x10", "hessian_x10"))), "w").write(out)
open(os.path.join(PROF, tag + "_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
if traffic:
    json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
print("\n".join(lines))
print(json.dumps(traffic, indent=1))
