"""Batched entry point of the T(E) hot path: every reference call `wfm.kernel(...)` is a batch of one
of `transmission_sweep`.  Host-side validation mirrors the reference's checks; all
arithmetic runs in libtransport_b200.so on the GPU.
"""
import numpy as np

from . import _lib
import threading

from ._lib import (TX_HOP_AUTO, TX_LEAD_CLOSED, TX_LEAD_SANCHO, TX_LEAD_SEPARABLE, TX_LEAD_TABLE, TransportError, as_c128,
                   as_f64, check, ptr)

_lead_lock = threading.Lock()   # keeps "set lead parameters, sweep" atomic on the shared default context


def _norm_blocks(h, tnn):
    if not isinstance(h, np.ndarray):
        raise TypeError
    if not isinstance(tnn, np.ndarray):
        raise TypeError
    h = as_c128(h)
    tnn = as_c128(tnn)
    if h.ndim == 3:
        h = h[None]
    if tnn.ndim == 3:
        tnn = tnn[None]
    if h.ndim != 4 or tnn.ndim != 4:
        raise ValueError("h must be [M,n,n] or [P,M,n,n]; tnn [M-1,n,n] or [P,M-1,n,n]")
    if h.shape[-1] != h.shape[-2] or tnn.shape[-2:] != h.shape[-2:]:
        raise ValueError("blocks must be square and of equal size")
    if not tnn.shape[1] + 1 == h.shape[1]:          # wfm/__init__.py:179
        raise ValueError
    return h, tnn


def check_closed_leads(h):
    """`g_closed` refuses non-diagonal lead blocks (wfm/__init__.py:327)."""
    for blk in (h[..., 0, :, :], h[..., -1, :, :]):
        off = blk - np.einsum("...ii->...i", blk)[..., None] * np.eye(blk.shape[-1])
        if np.any(off):
            raise Exception("Not diagonal:\n" + str(blk))


class Context:
    """A `tx_context`: private device workspace, scratch, stream and tuning options (include/transport_b200.h).
    Calls through one Context are serialised; distinct Contexts are independent, so host threads that want
    concurrent sweeps make one each.  `stream` is a raw cudaStream_t handle (int) or None for a context-owned
    non-blocking stream."""

    def __init__(self, device=0, stream=None):
        import ctypes
        lib = _lib.load()
        self._lib = lib
        self._h = ctypes.c_void_p()
        check(lib.tx_context_create(int(device), ctypes.byref(self._h)))
        if stream is not None:
            check(lib.tx_context_set_stream(self._h, ctypes.c_void_p(int(stream)), 1))

    def set_option(self, option, value):
        check(self._lib.tx_context_set_option(self._h, int(option), int(value)))

    def close(self):
        if self._h:
            self._lib.tx_context_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def transmission_sweep(h, tnn, Es, converger="g_closed", sources=None, h1=None, params=None,
                       sigL=None, sigR=None, want_resid=False, out=None, want_total=False,
                       want_caroli=False, want_rt=True, want_spin=False, n_elec_up=None, spin_out=None,
                       total_out=None, context=None):
    """Reflection / transmission probabilities for every (Hamiltonian, energy, source).

    Args
      h      [M,n,n] or [P,M,n,n]   on-site blocks (block 0 / M-1 = first cell of left / right lead)
      tnn    [M-1,n,n] or [P,M-1,n,n] upper hopping blocks H_{j,j+1}
      Es     [nE] complex energies
      converger  "g_closed" (fused closed-form leads); "table" with sigL/sigR [nE,n,n] or [P,nE,n,n];
             ("sancho", eta[, tol[, max_iter]]) Sancho-Rubio decimation of multi-channel leads at E + i*eta, or
             "separable" / ("separable", eta) for leads whose hopping block is t*I -- both build the self-energy
             tables on the device ahead of the sweep (they never travel to the host)
      sources None -> all n one-hot sources; else [n_src, n] real amplitude vectors (Ajsigma)
      h1, params  optional affine family: Hamiltonian p = h + params[p]*h1  (h must then be [M,n,n])
      out    optional (R, T) float64 arrays to fill (e.g. views of pinned memory)
      want_rt  False: the per-channel arrays are neither allocated, written nor copied (compact sweeps)
      want_spin  2 or 4 (True = 4): spin-resolved sums fused into the device epilogue, [P, nE, n_src, want_spin] =
             (T no-flip, T flip[, R no-flip, R flip]) with channels [:n_elec_up] = electron spin up (default n//2;
             run_wfm_gate.py:404-405,473-477); `spin_out` / `total_out` optionally supply the arrays to fill (e.g. pinned memory)
      context  a `Context` (own workspace / stream) or None for the per-device default context
    Returns
      R, T  float64 [P, nE, n_src, n] (if want_rt)  (+ resid [P, nE, n_src] = sum R + sum T - 1 if want_resid,
      + total [P, nE] = T summed over sources and channels if want_total,
      + caroli [P, nE] = Tr[Gamma_R G Gamma_L G^dag] with matrix Gamma if want_caroli,
      + spin [P, nE, n_src, 2|4] if want_spin)
    """
    lib = _lib.load()
    h, tnn = _norm_blocks(h, tnn)
    Es = as_c128(np.atleast_1d(Es))
    n, M = h.shape[-1], h.shape[1]
    n_h, n_t = h.shape[0], tnn.shape[0]
    if h1 is not None:
        if n_h != 1:
            raise ValueError("affine sweeps take a single base Hamiltonian")
        h1 = as_c128(h1)
        params = as_f64(params)
        if h1.shape != h.shape[1:]:
            raise ValueError("h1 must have the shape of h")
        n_ham = len(params)
    else:
        n_ham = max(n_h, n_t)
    if n_h not in (1, n_ham) or n_t not in (1, n_ham):
        raise ValueError("h and tnn batch sizes disagree")

    if converger == "g_closed":
        check_closed_leads(h)
        if h1 is not None:
            check_closed_leads(h1[None])
        lead_mode, n_sig = TX_LEAD_CLOSED, 1
        sigL = sigR = None
    elif converger == "table":
        sigL, sigR = as_c128(sigL), as_c128(sigR)
        if sigL.ndim == 3:
            sigL, sigR = sigL[None], sigR[None]
        if sigL.shape != sigR.shape or sigL.shape[1:] != (len(Es), n, n) or sigL.shape[0] not in (1, n_ham):
            raise ValueError("sigL/sigR must be [nE,n,n] or [P,nE,n,n]")
        lead_mode, n_sig = TX_LEAD_TABLE, sigL.shape[0]
    elif converger == "separable" or (isinstance(converger, tuple) and converger and converger[0] in ("sancho", "separable")):
        kind = converger if isinstance(converger, str) else converger[0]
        args = () if isinstance(converger, str) else tuple(converger[1:])
        lead_params = (float(args[0]) if len(args) > 0 else 0.0, float(args[1]) if len(args) > 1 else 1e-14,
                       int(args[2]) if len(args) > 2 else 200)
        lead_mode, n_sig = (TX_LEAD_SANCHO if kind == "sancho" else TX_LEAD_SEPARABLE), 1
        sigL = sigR = None
    else:
        raise Exception("converger = " + str(converger) + " not supported")

    if sources is None:
        src, n_src = None, n
    else:
        src = as_f64(np.atleast_2d(sources))
        if src.shape[1] != n:
            raise ValueError("sources must be [n_src, n]")
        n_src = src.shape[0]

    shape = (n_ham, len(Es), n_src, n)
    if not want_rt:
        if out is not None:
            raise ValueError("out given with want_rt=False")
        if not (want_resid or want_total or want_caroli or want_spin):
            raise ValueError("no output requested")
        R = T = None
    elif out is None:
        R = np.empty(shape, dtype=np.float64)
        T = np.empty(shape, dtype=np.float64)
    else:
        R, T = out
        if R.shape != shape or T.shape != shape or R.dtype != np.float64 or T.dtype != np.float64:
            raise ValueError("out arrays must be float64 of shape %s" % (shape,))
    resid = np.empty(shape[:3], dtype=np.float64) if want_resid else None
    total = None
    if want_total:
        total = np.empty(shape[:2], dtype=np.float64) if total_out is None else total_out
        if total.shape != shape[:2] or total.dtype != np.float64:
            raise ValueError("total_out must be float64 of shape %s" % (shape[:2],))
    caroli = np.empty(shape[:2], dtype=np.float64) if want_caroli else None
    spin, spin_n, n_up = None, 0, 0
    if want_spin:
        spin_n = 4 if want_spin is True else int(want_spin)
        if spin_n not in (2, 4):
            raise ValueError("want_spin must be 2 or 4")
        n_up = n // 2 if n_elec_up is None else int(n_elec_up)
        if spin_out is None:
            spin = np.empty(shape[:3] + (spin_n,), dtype=np.float64)
        else:
            spin = spin_out
            if spin.shape != shape[:3] + (spin_n,) or spin.dtype != np.float64:
                raise ValueError("spin_out must be float64 of shape %s" % (shape[:3] + (spin_n,),))

    cxh = context._h if context is not None else None
    if lead_mode in (TX_LEAD_SANCHO, TX_LEAD_SEPARABLE):
        _lead_lock.acquire()
        check(lib.tx_context_set_lead_params(cxh, *lead_params))
    try:
        rc = _call_sweep(lib, cxh, h, n_h, h1, params, tnn, n_t, n, M, n_ham, Es, lead_mode, sigL, sigR, n_sig, src, n_src,
                         R, T, resid, total, caroli, n_up, spin_n, spin)
    finally:
        if lead_mode in (TX_LEAD_SANCHO, TX_LEAD_SEPARABLE):
            _lead_lock.release()
    if rc == _lib.TX_ERR_NOCONV:
        raise Exception(lib.tx_last_error().decode())
    return _finish(rc, want_rt, want_resid, want_total, want_caroli, want_spin, R, T, resid, total, caroli, spin)


def _call_sweep(lib, cxh, h, n_h, h1, params, tnn, n_t, n, M, n_ham, Es, lead_mode, sigL, sigR, n_sig, src, n_src,
                R, T, resid, total, caroli, n_up, spin_n, spin):
    return lib.tx_transmission_sweep_spin(cxh,
                                        ptr(h), n_h, ptr(h1), ptr(params), ptr(tnn), n_t, n, M, n_ham,
                                        ptr(Es), len(Es), lead_mode, ptr(sigL), ptr(sigR), n_sig,
                                        TX_HOP_AUTO, ptr(src), None, n_src, ptr(R), ptr(T), ptr(resid), ptr(total),
                                        ptr(caroli), n_up, spin_n, ptr(spin))


def _finish(rc, want_rt, want_resid, want_total, want_caroli, want_spin, R, T, resid, total, caroli, spin):
    if rc == _lib.TX_ERR_SINGULAR:
        # numpy raises LinAlgError for an exactly singular E - H' (wfm/__init__.py:162)
        raise np.linalg.LinAlgError("Singular matrix")
    check(rc)
    ret = (R, T) if want_rt else ()
    if want_resid:
        ret += (resid,)
    if want_total:
        ret += (total,)
    if want_caroli:
        ret += (caroli,)
    if want_spin:
        ret += (spin,)
    return ret[0] if len(ret) == 1 else ret
