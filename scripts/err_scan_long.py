"""Unitarity residual of the N = 10^4 chain (cfg3) for several chunk lengths / growth bounds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_b200 import _lib, configs
from transport_b200.sweep import transmission_sweep
lib = _lib.load()
h, tnn, _ = configs.disordered_blocks(N=10_000, n=8, W=0.05, seed=1234)
Es = np.linspace(-1.9, 1.9, 100_000).astype(complex)[::250]
ref = None
for kmax, gexp in [(1, 4), (8, 4), (16, 4), (32, 2), (32, 3), (32, 4), (32, 5), (64, 4)]:
    _lib.check(lib.tx_set_option(1, kmax)); _lib.check(lib.tx_set_option(2, gexp))
    R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
    r = np.abs(res[0]).max(axis=-1)
    if ref is None: ref = T.copy()
    worst = np.argsort(r)[-3:]
    dT = np.abs(T - ref).max(axis=(0, 2, 3))
    print("kmax %2d gexp %d  max|R+T-1| %.2e (median %.1e) at E=%s   max|T-T_plain| %.2e (median %.1e)" % (
        kmax, gexp, r.max(), np.median(r), np.round(Es.real[worst], 3), dT.max(), np.median(dT)))
