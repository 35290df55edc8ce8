#!/usr/bin/env python
"""SURVEY §8f rank 2 at the reference's own scale: wavefunction output psi (first block column of G) for the gate
sweeps of run_wfm_gate.py (n = 8, M = 505, 299 energies; ~20 s per point in the reference).
    python scripts/bench_psi.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rgf_oracle                      # checker only
from transport_b200 import configs, green, leads

M, nE = 505, 299
h, tnn, tnnn = configs.cicc_blocks(1, 1.0, -0.2, 0.0, 0.0, 0.0, M - 5, 2)
M = h.shape[0]
Es = (np.logspace(-4, 0, nE) - 2.0).astype(complex)
SL, SR = leads.self_energies_batched(h, tnn, Es, "g_closed")
green.green_column0(h, tnn, Es[:2], SL[:2], SR[:2])         # warm-up
t0 = time.time()
G = green.green_column0(h, tnn, Es, SL, SR)
dt = time.time() - t0
idx = [0, nE // 2, nE - 1]
G00, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es[idx], SL[idx], SR[idx])
err = max(np.abs(G[idx, 0] - G00).max() / np.abs(G00).max(), np.abs(G[idx, -1] - GN0).max() / np.abs(GN0).max())
print(json.dumps({"config": "psi / first block column, n=8, M=%d, %d energies (run_wfm_gate.py scale)" % (M, nE),
                  "host_call_s": dt, "points_per_s": nE / dt, "output_MB": G.nbytes / 1e6,
                  "max_rel_err_G00_GN0_vs_oracle": float(err)}))
