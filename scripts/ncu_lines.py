"""Per-source-line instruction and stall-sample totals of an .ncu-rep (needs -lineinfo and --import-source on).
usage: python scripts/ncu_lines.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; tot_i = tot_s = 0
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 8: continue
    if r[0] == "":   # sass row
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    try:
        ni = int(r[7]); ns = int(r[6])
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0]); a[0] += ni; a[1] += ns
    tot_i += ni; tot_s += ns
print("total warp-instr %d samples %d" % (tot_i, tot_s))
for (f, ln, src), (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-14s %4d  instr %5.1f%%  samples %5.1f%%  %s" % (f, ln, 100.0 * ni / tot_i, 100.0 * ns / max(tot_s, 1), src))
if len(sys.argv) > 3:   # region totals for one file: name:lo-hi,...
    for spec in sys.argv[3].split(","):
        f, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        ni = sum(v[0] for k, v in agg.items() if k[0] == f and lo <= k[1] <= hi)
        ns = sum(v[1] for k, v in agg.items() if k[0] == f and lo <= k[1] <= hi)
        print("%-28s instr %5.1f%% (%d)  samples %5.1f%%" % (spec, 100.0 * ni / tot_i, ni, 100.0 * ns / max(tot_s, 1)))
