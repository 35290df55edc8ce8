#!/usr/bin/env python
"""Per-source-line attribution of an .ncu-rep (built with -lineinfo, captured with --import-source on):
joins the SASS page (executed warp-instructions and stall samples per instruction address) with the line table of the
cubin (nvdisasm -g) and prints, per source line, warp-instructions, FP64-pipe issue slots (DFMA/DMUL/DADD/DSETP = 1,
DMMA.8x8x4 = 8: the FP64 tensor path shares the pipe and moves 8 FMA per lane) and stall samples.

usage: python scripts/ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTRING [launch_index] [top_n]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, pattern = sys.argv[1], sys.argv[2], sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 45

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the CSV holds one table per profiled launch, each introduced by a "Kernel Name" style preamble line
tables, cur, last = [], None, -1
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":          # preamble of the next profiled launch
        cur = None
        continue
    if row and row[0] == "Address":
        hdr = row
        continue
    if row and re.match(r"^0x[0-9a-f]+$|^[0-9a-f]+$", row[0]):
        addr = int(row[0], 16)
        if cur is None or addr <= last:          # addresses restart: next profiled launch
            cur = {"hdr": hdr, "rows": []}
            tables.append(cur)
        last = addr
        cur["rows"].append(row)
tab = tables[launch]
ix = {h: i for i, h in enumerate(tab["hdr"])}

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
# every .text section whose (mangled) name contains the pattern; a template has many instantiations, so the one whose
# instruction count equals the profiled kernel's SASS table is taken (pass the full mangled name to be sure)
n_sass = len(tab["rows"])
sections = []
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    cur_map, where = None, None
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            cur_map = {} if pattern in ln else None
            if cur_map is not None:
                sections.append((ln.strip(), cur_map))
            where = None
            continue
        if cur_map is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            where = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and where:
            cur_map.setdefault(int(m.group(1), 16), where)
if not sections:
    sys.exit("no .text section matches %r" % pattern)
sections.sort(key=lambda sm: abs(len(sm[1]) - n_sass))
name, line_of = sections[0]
if len(sections) > 1:
    print("# %d sections match %r; using %s (%d instructions with line info, SASS table %d)" % (len(sections), pattern, name, len(line_of), n_sass), file=sys.stderr)

per = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
base = None
for r in tab["rows"]:
    addr = int(r[ix["Address"]], 16)
    if base is None:
        base = addr
    sass = r[ix["Source"]]
    n = int(float(r[ix["Instructions Executed"]] or 0))
    s = int(float(r[ix["# Samples"]] or 0))
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    slots = n * (8 if op.startswith("DMMA") else 1 if re.match(r"D(FMA|MUL|ADD|SETP|MNMX)", op) else 0)
    key = line_of.get(addr - base, ("?", 0))
    for i, v in enumerate((n, slots, s)):
        per[key][i] += v
        tot[i] += v
src_cache = {}


def text(key):
    f, l = key
    if f not in src_cache:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "transport_b200", "csrc", f))
        src_cache[f] = open(cand[0]).read().splitlines() if cand else []
    return src_cache[f][l - 1].strip()[:90] if 0 < l <= len(src_cache[f]) else ""


print("total warp-instr %d  fp64-pipe slots %d  samples %d" % tuple(tot))
for key, (n, sl, s) in sorted(per.items(), key=lambda kv: -kv[1][1] - kv[1][0] * 0.2)[:top]:
    print("%-14s %5d  instr %5.1f%%  fp64slots %5.1f%%  samples %5.1f%%  %s"
          % (key[0], key[1], 100.0 * n / tot[0], 100.0 * sl / max(tot[1], 1), 100.0 * s / max(tot[2], 1), text(key)))
