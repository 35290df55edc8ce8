"""ctypes binding of libtransport_b200.so (include/transport_b200.h).

The library is the product: there is no Python/numpy fallback for any compute entry point.
`load()` raises if the shared object has not been built (python -m transport_b200.build), and
every compute call raises `TransportError` when no CUDA device is present.
"""
import ctypes
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libtransport_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "transport_b200.h")

TX_OK, TX_ERR_ARG, TX_ERR_CUDA, TX_ERR_SINGULAR, TX_ERR_NOCONV, TX_ERR_UNSUPPORTED = range(6)
TX_LEAD_CLOSED, TX_LEAD_TABLE, TX_LEAD_SANCHO, TX_LEAD_SEPARABLE = 0, 1, 2, 3
TX_HOP_AUTO, TX_HOP_SCALAR, TX_HOP_GENERAL = 0, 1, 2
TX_GF_CLOSED, TX_GF_RICEMELE, TX_GF_FIXEDPOINT, TX_GF_SANCHO, TX_GF_SEPARABLE = 0, 1, 2, 3, 4


class TransportError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("transport_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None

c_void_p, c_int, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
_P = c_void_p

_SIGNATURES = {
    "tx_version": (ctypes.c_char_p, []),
    "tx_last_error": (ctypes.c_char_p, []),
    "tx_device_count": (c_int, []),
    "tx_set_device": (c_int, [c_int]),
    "tx_device_synchronize": (c_int, []),
    "tx_set_variant": (c_int, [c_int]),
    "tx_set_option": (c_int, [c_int, c_int]),
    "tx_sweep_launch_count": (c_int, [c_int, c_int, c_int]),
    "tx_launch_counter": (ctypes.c_longlong, []),
    "tx_transmission_sweep": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int,
                                      c_int, _P, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P, _P]),
    "tx_transmission_sweep_large_dev": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int,
                                                _P, _P, c_int, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tx_sweep_workspace_elems": (ctypes.c_longlong, [c_int, ctypes.c_longlong]),
    "tx_transmission_sweep_dev": (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int,
                                          c_int, _P, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P, _P, _P]),
    "tx_transmission_sweep_spin": (c_int, [_P, _P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int,
                                           c_int, _P, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P, _P,
                                           c_int, c_int, _P]),
    "tx_transmission_sweep_spin_dev": (c_int, [_P, _P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_int,
                                               c_int, _P, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P,
                                               c_int, c_int, _P, _P, _P]),
    "tx_context_create": (c_int, [c_int, _P]),
    "tx_context_destroy": (c_int, [_P]),
    "tx_context_set_stream": (c_int, [_P, _P, c_int]),
    "tx_context_set_option": (c_int, [_P, c_int, c_int]),
    "tx_context_set_lead_params": (c_int, [_P, c_double, c_double, c_int]),
    "tx_context_last_lead_iterations": (c_int, [_P, _P]),
    "tx_context_set_gather": (c_int, [_P, _P, c_int, ctypes.c_longlong]),
    "tx_peer_alloc": (c_int, [ctypes.c_longlong, _P]),
    "tx_peer_free": (c_int, [_P]),
    "tx_peer_export": (c_int, [_P, _P]),
    "tx_peer_open": (c_int, [_P, _P]),
    "tx_peer_close": (c_int, [_P]),
    "tx_peer_barrier_dev": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "tx_surface_gf": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_double, c_double, c_int, c_int, c_int,
                              _P, _P, _P, _P, _P]),
    "tx_surface_gf_pair": (c_int, [_P, _P, c_int, _P, c_int, c_double, c_double, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "tx_surface_gf_pair_dev": (c_int, [_P, _P, c_int, _P, c_int, c_double, c_double, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tx_surface_gf_dev": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_double, c_double, c_int, c_int, c_int,
                                  _P, _P, _P, _P, _P, _P, _P]),
    "tx_g_ith": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_double, c_int, _P, _P]),
    "tx_green_column0": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P, _P, _P]),
    "tx_green_full": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P, _P, _P]),
    "tx_hmat": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "tx_landauer_current": (c_int, [_P, _P, c_int, _P, _P, c_int, c_double, _P, _P]),
    "tx_landauer_current_dev": (c_int, [_P, _P, c_int, _P, _P, c_int, c_double, _P, _P, _P]),
    "tx_hsysmat": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, c_int, c_int, _P]),
    "tx_tridiag_eigh_below": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P]),
    "tx_tridiag_count_below_dev": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    "tx_tridiag_eigh_below_dev": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P]),
    "tx_bardeen_wells": (c_int, [_P] * 10 + [c_int, c_int, c_int, c_int, c_double, c_int] + [_P] * 8),
    "tx_bardeen_workspace_bytes": (ctypes.c_longlong, [c_int, c_int, c_int, c_int]),
    "tx_bardeen_wells_dev": (c_int, [_P] * 10 + [c_int, c_int, c_int, c_int, c_double, c_int] + [_P] * 8),
    "tx_bardeen_region_sweep": (c_int, [_P] * 8 + [c_int] + [_P] * 6 + [c_int] * 7 + [_P] * 3 + [c_double, c_int, c_int, c_int, _P]
                                + [_P] * 6 + [_P, c_int, c_double, c_double, _P]),
    "tx_real_gemm": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int]),
    "tx_bardeen_current": (c_int, [_P, _P, c_int, c_int, c_int, _P, c_int, c_double, c_double, _P]),
    "tx_bardeen_current_dev": (c_int, [_P, _P, c_int, c_int, c_int, _P, c_int, c_double, c_double, _P, _P]),
    "tx_bardeen_ts": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, _P, c_int, c_int, _P, _P]),
    "tx_debug_invert": (c_int, [_P, _P, c_int, c_int]),
    "tx_debug_invert_staged": (c_int, [_P, _P, c_int, c_int]),
    "tx_debug_last_invert_ms": (c_double, []),
    "tx_debug_gemm_ms": (c_int, [c_int, c_int, c_int, _P]),
    "tx_debug_gemm_tma_ms": (c_int, [c_int, c_int, c_int, _P]),
    "tx_measure_fp64_peak": (c_int, [_P, _P]),
    "tx_measure_dmma_peak": (c_int, [_P, _P]),
    "tx_microbench": (c_int, [c_int, _P]),
}


def declared_symbols():
    """Every function name include/transport_b200.h declares (used by the CPU export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tx_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load the shared library (once).  Raises if it is missing: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libtransport_b200.so is not built (run `python -m transport_b200.build`); "
                          "transport_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != TX_OK:
        raise TransportError(code, load().tx_last_error().decode())


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_void_p)


def as_c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    return int(load().tx_device_count())


def require_device():
    if device_count() < 1:
        raise TransportError(TX_ERR_CUDA, "no CUDA device available (transport_b200 has no CPU fallback)")
