"""Summarise an .ncu-rep: key metrics + per-opcode instruction/stall breakdown.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.csv] [launch_index]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr, units, vals = rows[0], rows[1], rows[2 + launch]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg']
out = []
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')):
        out.append((h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, samp = collections.Counter(), collections.Counter()
ti = ts = 0
last_addr = -1
seen = -1                                  # index of the profiled launch whose table is being read
for r in rows:
    if r and r[0] == "Kernel Name":        # preamble of the next profiled launch
        seen += 1
        last_addr = -1
        continue
    if seen != launch or len(r) < len(hdr):
        continue
    try:                                   # (older reports repeat a launch's table: first copy only)
        addr = int(r[ix['Address']], 16)
    except ValueError:
        continue
    if addr <= last_addr:
        break
    last_addr = addr
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
    op = m.group(2).split('.')[0] if m else '?'
    try:
        ni, ns = int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
    except ValueError:
        continue   # repeated header of a further kernel in the report
    ops[op] += ni
    samp[op] += ns
    ti += ni
    ts += ns
for h, u, v in out:
    print("%-80s %-14s %s" % (h, u, v))
print("\nopcode      warp-instr      %   stall-samples   %")
for op, c in ops.most_common(22):
    print("%-10s %12d %6.1f %10d %6.1f" % (op, c, 100.0 * c / ti, samp[op], 100.0 * samp[op] / max(ts, 1)))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerows(out)
        w.writerow([])
        w.writerow(["opcode", "warp_instructions", "stall_samples"])
        for op, c in ops.most_common(30):
            w.writerow([op, c, samp[op]])
