import json, os, sys, subprocess
for herm in (1, 0):
    for two in (0, 1):
        env = dict(os.environ)
        if two: env["TX_BENCH_TWO_DECIMATIONS"] = "1"
        env["TX_BENCH_HERM"] = str(herm)
        out = subprocess.run([sys.executable, "bench.py", "--config", "cfg4", "--no-cpu"], capture_output=True, text=True, env=env).stdout
        d = json.loads(out.strip().splitlines()[-1]); r = d["roofline"]
        print("herm_shortcut", herm, "two_decimations", two, round(d["ms_per_step"], 2), round(d["value"]), round(r["frac"], 3), {k: round(v["ms"], 2) for k, v in r["kernels"].items()}, d.get("parity_max_rel_err"), d.get("leads_max_err_vs_extended_precision"), flush=True)
