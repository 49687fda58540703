"""Aggregate the warp-stall samples of an exported `ncu --page source --csv --print-source cuda,sass` file per source line."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
agg = collections.OrderedDict()
stall_cols = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples"); stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or r[2] != "-": continue      # keep the per-line summary rows (Address == '-')
    try: n = int(r[si])
    except: continue
    key = (cur_file, int(r[0]))
    st = {h: int(r[i]) for i, h in stall_cols if r[i] not in ("", "-") and int(r[i]) > 0}
    a = agg.setdefault(key, [0, r[1].strip(), collections.Counter()])
    a[0] += n; a[2].update(st)
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for (f, ln), (n, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% %-12s:%4d  %-90s %s" % (100.0 * n / tot, f, ln, src[:90], dict(st.most_common(3))))
