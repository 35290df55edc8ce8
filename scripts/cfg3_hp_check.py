#!/usr/bin/env python
"""Device error of BASELINE config 3 at full length against the extended-precision fixture, energy by energy."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_b200 import configs  # noqa: E402
from transport_b200.sweep import transmission_sweep  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "hp_cfg3_M10002.npz"))
h, tnn, _ = configs.disordered_blocks(N=10_000, n=8, W=0.05, seed=1234)
if "--full-grid" in sys.argv:      # the whole 1e5-energy axis in one sweep (a point's chunk boundaries depend on its warp mates)
    Es = np.linspace(-1.9, 1.9, 100_000).astype(complex)
    R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
    R, T, res = R[:, g["idx"]], T[:, g["idx"]], res[:, g["idx"]]
else:
    R, T, res = transmission_sweep(h, tnn, g["Es"].astype(complex), want_resid=True)


def rel(a, b):
    den = np.maximum(np.abs(b), 1e-13 * np.abs(b).max(-1, keepdims=True))
    return (np.abs(a - b) / den).max(axis=(1, 2))


dev = np.maximum(rel(T[0], g["T_hp"]), rel(R[0], g["R_hp"]))
kap, f64 = g["kappa"], g["f64_err"]
ratio = dev / (kap * 2.2e-16)
order = np.argsort(-ratio)[:8]
print(json.dumps({"dev_max": float(dev.max()), "dev_median": float(np.median(dev)), "f64_max": float(f64.max()),
                  "dev_over_kappa_eps_max": float(ratio.max()), "dev_over_kappa_eps_median": float(np.median(ratio)),
                  "f64_over_kappa_eps_max": float((f64 / (kap * 2.2e-16)).max()),
                  "unitarity_max": float(np.abs(res).max()),
                  "n_above_1e-10": int((dev > 1e-10).sum()), "n": len(dev),
                  "worst": [(int(g["idx"][i]), float(dev[i]), float(kap[i]), float(ratio[i]), float(f64[i])) for i in order]}))
