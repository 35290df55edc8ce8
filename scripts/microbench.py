"""Prints the FP64 / exchange-primitive rates of the current GPU (csrc/microbench.cu)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transport_b200 import _lib  # noqa: E402

lib = _lib.load()
out = {}
v, ms = ctypes.c_double(), ctypes.c_double()
_lib.check(lib.tx_measure_fp64_peak(ctypes.byref(v), ctypes.byref(ms)))
out["dfma_tflops"] = v.value
_lib.check(lib.tx_measure_dmma_peak(ctypes.byref(v), ctypes.byref(ms)))
out["dmma_m8n8k4_tflops"] = v.value
for which, name in ((0, "shfl_b32_Gwarpinstr_s"), (1, "lds128_groupbcast_Gwarpinstr_s"),
                    (2, "lds128_groupbcast_2way_Gwarpinstr_s"), (3, "dfma_lds_8to1_mix_tflops"), (4, "dfma_plus_dmma_equal_flops_mix_tflops")):
    _lib.check(lib.tx_microbench(which, ctypes.byref(v)))
    out[name] = v.value
for w in (10, 11, 12, 13):
    _lib.check(lib.tx_microbench(w, ctypes.byref(v)))
    out["dfma_tflops_%d_warps_per_scheduler" % (w - 9)] = v.value
for wps in (1, 2, 3):
    for ch in (1, 2, 4, 8, 16):
        _lib.check(lib.tx_microbench(100 * wps + ch, ctypes.byref(v)))
        out["dmma_tflops_%dwps_%dchains" % (wps, ch)] = round(v.value, 2)
print(json.dumps(out))
