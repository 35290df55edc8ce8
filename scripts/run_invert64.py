#!/usr/bin/env python
"""Runs the blocked CTA inverse on 592 random 64 x 64 blocks (two per SM resident) — the ncu target for the inverse."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_b200 import _lib  # noqa: E402

lib = _lib.load()
rng = np.random.default_rng(0)
n, count = 64, 592
A = (rng.standard_normal((count, n, n)) + 1j * rng.standard_normal((count, n, n))).astype(np.complex128)
out = np.empty_like(A)
_lib.check(lib.tx_debug_invert_staged(_lib.ptr(A), _lib.ptr(out), count, n))
print("ms", lib.tx_debug_last_invert_ms(), "err", float(np.max(np.abs(out[0] @ A[0] - np.eye(n)))))
