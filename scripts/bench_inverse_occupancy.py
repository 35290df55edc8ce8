import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from transport_b200 import _lib
lib = _lib.load()
lib.tx_debug_last_invert_ms.restype = __import__("ctypes").c_double
rng = np.random.default_rng(0)
for n in (64, 56, 72, 128):
    for count in (148, 296, 592):
        A = (rng.standard_normal((count, n, n)) + 1j * rng.standard_normal((count, n, n))).astype(np.complex128)
        out = np.empty_like(A)
        _lib.check(lib.tx_debug_invert_staged(_lib.ptr(A), _lib.ptr(out), count, n))
        print("n", n, "blocks", count, "ms", round(lib.tx_debug_last_invert_ms(), 4), flush=True)
