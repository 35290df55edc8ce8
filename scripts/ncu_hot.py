"""Top SASS instructions by stall samples, with the dominant stall reasons.
usage: python scripts/ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for i, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    ns = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:3]
    data.append((ns, i, r[ix['Source']].strip(), st, int(r[ix['Instructions Executed']])))
tot = sum(d[0] for d in data)
# aggregate stall type totals by opcode for long_sb
for ns, i, s, st, ne in sorted(data, reverse=True)[:N]:
    print("%6d %5.2f%% #%5d x%-9d %-70s %s" % (ns, 100.0 * ns / tot, i, ne, s[:70], " ".join("%s:%d" % (b, a) for a, b in st if a)))
