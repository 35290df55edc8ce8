#!/usr/bin/env python
"""A/B of the n = 64 block product: fragments straight from L2 vs the B operand staged in shared memory by TMA
(tx_debug_gemm_ms / tx_debug_gemm_tma_ms: `count` CTAs, 2 per SM, `reps` products each on their own global-memory blocks)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_b200 import _lib  # noqa: E402

lib = _lib.load()
out = {}
for name, fn in (("l2_direct", lib.tx_debug_gemm_ms), ("tma_staged", lib.tx_debug_gemm_tma_ms)):
    for count in (296, 592):
        ms = ctypes.c_double()
        _lib.check(fn(64, count, 200, ctypes.byref(ms)))
        _lib.check(fn(64, count, 200, ctypes.byref(ms)))
        us = ms.value * 1e3 / 200 / (count / 296)
        out["%s_%dctas" % (name, count)] = {"us_per_product_per_wave": us, "tflops": 8 * 64 ** 3 * 296 / (us * 1e-6) / 1e12}
print(json.dumps(out))
