"""Quick GPU diagnostics: device inverse and sweep errors against the oracle for several block sizes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rgf_oracle  # noqa: E402
from transport_b200 import _lib  # noqa: E402
from transport_b200.sweep import transmission_sweep  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
for n in (1, 2, 3, 4, 6, 8, 12, 16):
    A = rng.standard_normal((37, n, n)) + 1j * rng.standard_normal((37, n, n))
    out = np.empty_like(A)
    rc = lib.tx_debug_invert(_lib.ptr(A), _lib.ptr(out), len(A), n)
    err = np.abs(out @ A - np.eye(n)).max() if rc == 0 else np.nan
    print("inverse n=%2d rc=%d max|inv*A-I|=%.2e nan=%d" % (n, rc, err, np.isnan(out).sum()))
for n in (1, 2, 3, 4, 8, 16):
    for general in (False, True):
        M = 6
        Ah = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
        h = 0.3 * (Ah + np.conj(np.swapaxes(Ah, 1, 2))) / 2
        h[0] = 0
        h[-1] = 0
        tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
        if general:
            tnn[1:-1] += 0.2 * (rng.standard_normal((M - 3, n, n)) + 1j * rng.standard_normal((M - 3, n, n)))
        Es = np.linspace(-1.7, 1.6, 11).astype(complex)
        try:
            R, T = transmission_sweep(h, tnn, Es)
        except Exception as exc:
            print("sweep n=%2d general=%d raised %r" % (n, general, exc))
            continue
        oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, np.eye(n))
        print("sweep n=%2d general=%d  max|dT|=%.2e max|dR|=%.2e nanT=%d/%d" %
              (n, general, np.nanmax(np.abs(T[0] - oT)) if not np.isnan(T).all() else np.nan,
               np.nanmax(np.abs(R[0] - oR)) if not np.isnan(R).all() else np.nan, np.isnan(T).sum(), T.size))
