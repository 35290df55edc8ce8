#!/usr/bin/env python
"""Finds the energies of BASELINE config 3 (n=8, M=10 002, 1e5 energies) with the worst device unitarity residual and
writes them to gpurun_out/cfg3_worst_idx.json, for oracle/make_golden_hp.py (the extended-precision fixture)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_b200 import configs  # noqa: E402
from transport_b200.sweep import transmission_sweep  # noqa: E402

h, tnn, _ = configs.disordered_blocks(N=10_000, n=8, W=0.05, seed=1234)
Es = np.linspace(-1.9, 1.9, 100_000).astype(complex)
resid, tot = transmission_sweep(h, tnn, Es, want_rt=False, want_resid=True, want_total=True)
worst = np.abs(resid[0]).max(axis=-1)
order = np.argsort(-worst)
out = {"worst_idx": [int(i) for i in order[:64]], "worst_resid": [float(worst[i]) for i in order[:64]],
       "median_resid": float(np.median(worst)), "p99_resid": float(np.quantile(worst, 0.99)),
       "p999_resid": float(np.quantile(worst, 0.999)), "count_above_1e-10": int((worst > 1e-10).sum()),
       "count_above_1e-9": int((worst > 1e-9).sum()), "count_above_1e-8": int((worst > 1e-8).sum())}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "cfg3_worst_idx.json"), "w") as f:
    json.dump(out, f)
print(json.dumps(out))
