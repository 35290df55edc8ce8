#!/bin/bash
# usage: scripts/bench_variants.sh "1 4 5"   -> one line per variant: value, ms, frac, parity
for v in $1; do
  python bench.py --steps 5 --warmup 3 --variant $v --no-cpu > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err || { echo "v$v FAILED"; tail -5 gpurun_out/bench_v$v.err; continue; }
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v$v.json"))
print("variant $v: %.4g pts/s  kernel %.3f ms  frac %.3f  e2e %.3g  parity %.2e" % (d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["parity_max_rel_err_vs_reference_golden"]))
PY
done
