"""Opcode totals per source region of an .ncu-rep.  usage: ncu_regions.py rep file:lo-hi[,file:lo-hi...] [warps]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; specs = sys.argv[2].split(","); warps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; cur_line = None
data = collections.defaultdict(collections.Counter)   # (file,line) -> opcode -> count
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if len(r) < 8: continue
    if r[0] != "":
        try: cur_line = int(r[0])
        except ValueError: cur_line = None
        continue
    if cur_line is None: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3])
    if not m: continue
    try: ni = int(r[7])
    except ValueError: continue
    data[(cur_file, cur_line)][m.group(2)] += ni
for spec in specs:
    f, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
    tot = collections.Counter()
    for (ff, ln), c in data.items():
        if ff == f and lo <= ln <= hi: tot.update(c)
    s = sum(tot.values())
    print("%-26s total %9.1f  " % (spec, s / warps) + " ".join("%s %.0f" % (k, v / warps) for k, v in tot.most_common(9)))
