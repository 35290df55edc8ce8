"""Error of the telescoped sweep against the plain numpy recursion for a range of chunk lengths / growth bounds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import rt_err
from oracle import rgf_oracle as ro
from transport_b200 import _lib, configs
from transport_b200.sweep import transmission_sweep
lib = _lib.load()
h, tnn, _ = configs.disordered_blocks(N=1000, n=8, W=0.05, seed=7)
Es = np.linspace(-1.9, 1.9, 40).astype(complex)
oR, oT = ro.transmission_closed(h, tnn, Es[::8], np.eye(8))
for kmax, gexp in [(1, 4), (8, 4), (16, 4), (24, 4), (32, 2), (32, 3), (32, 4), (32, 5), (32, 6), (48, 4), (64, 4)]:
    _lib.check(lib.tx_set_option(1, kmax)); _lib.check(lib.tx_set_option(2, gexp))
    R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
    print("kmax %3d gexp %2d  T %.2e R %.2e unit %.2e" % (kmax, gexp, rt_err(T[0, ::8], oT), rt_err(R[0, ::8], oR), np.abs(res).max()))
