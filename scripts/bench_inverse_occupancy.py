"""Blocked CTA inverse (block_ops.cuh) per block edge and number of co-resident CTAs: 148 / 296 / 592 blocks = one CTA per
SM, two, two waves of two.

    python scripts/bench_inverse_occupancy.py
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transport_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.tx_debug_last_invert_ms.restype = ctypes.c_double
rng = np.random.default_rng(0)
for n in (32, 64, 56, 72, 128):
    for count in (148, 296, 592):
        A = (rng.standard_normal((count, n, n)) + 1j * rng.standard_normal((count, n, n))).astype(np.complex128)
        out = np.empty_like(A)
        _lib.check(lib.tx_debug_invert_staged(_lib.ptr(A), _lib.ptr(out), count, n))
        err = float(np.max(np.abs(out[0] @ A[0] - np.eye(n))))
        print("n", n, "blocks", count, "ms", round(lib.tx_debug_last_invert_ms(), 4), "err", "%.1e" % err, flush=True)
