#!/usr/bin/env python
"""Device time of the CTA-cooperative staged inverse (block_ops.cuh) per block size: tx_debug_invert_staged on a
batch that fills the machine several times.  python scripts/bench_inverse.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_b200 import _lib
from transport_b200._lib import check, ptr
lib = _lib.load()
rng = np.random.default_rng(0)
for n in (24, 40, 56, 64):
    count = 296 * 8
    A = np.ascontiguousarray(rng.normal(size=(count, n, n)) + 1j * rng.normal(size=(count, n, n)))
    out = np.empty_like(A)
    check(lib.tx_debug_invert_staged(ptr(A), ptr(out), count, n))
    ms = lib.tx_debug_last_invert_ms()
    err = max(np.abs(out[i] @ A[i] - np.eye(n)).max() for i in range(0, count, 97))
    print(json.dumps({"n": n, "blocks": count, "ms": ms, "us_per_block_per_sm_slot": ms * 1e3 / (count / 296.0),
                      "gflops_8n3": 8 * n ** 3 * count / (ms * 1e-3) / 1e9, "max_resid": err}))

import ctypes
for n in (24, 40, 64):
    ms = ctypes.c_double()
    reps = 20
    check(lib.tx_debug_gemm_ms(n, 296, reps, ctypes.byref(ms)))
    us = ms.value * 1e3 / reps
    print(json.dumps({"gemm_n": n, "ctas": 296, "us_per_product_2_ctas_per_sm": us,
                      "tflops": 8 * n ** 3 * 296 / (us * 1e-6) / 1e12}))
