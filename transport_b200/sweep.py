"""Batched entry point of the T(E) hot path: every reference call `wfm.kernel(...)` is a batch of one
of `transmission_sweep`.  Host-side validation mirrors the reference's checks; all
arithmetic runs in libtransport_b200.so on the GPU.
"""
import numpy as np

from . import _lib
from ._lib import (TX_HOP_AUTO, TX_LEAD_CLOSED, TX_LEAD_TABLE, TransportError, as_c128, as_f64, check, ptr)


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


def transmission_sweep(h, tnn, Es, converger="g_closed", sources=None, h1=None, params=None,
                       sigL=None, sigR=None, want_resid=False, out=None, want_total=False,
                       want_caroli=False):
    """Reflection / transmission probabilities for every (Hamiltonian, energy, source).

    Args
      h      [M,n,n] or [P,M,n,n]   on-site blocks (block 0 / M-1 = first cell of left / right lead)
      tnn    [M-1,n,n] or [P,M-1,n,n] upper hopping blocks H_{j,j+1}
      Es     [nE] complex energies
      converger  "g_closed" (fused closed-form leads) or "table" with sigL/sigR [nE,n,n] or [P,nE,n,n]
      sources None -> all n one-hot sources; else [n_src, n] real amplitude vectors (Ajsigma)
      h1, params  optional affine family: Hamiltonian p = h + params[p]*h1  (h must then be [M,n,n])
      out    optional (R, T) float64 arrays to fill (e.g. views of pinned memory)
    Returns
      R, T  float64 [P, nE, n_src, n]  (+ resid [P, nE, n_src] = sum R + sum T - 1 if want_resid,
      + total [P, nE] = T summed over sources and channels if want_total,
      + caroli [P, nE] = Tr[Gamma_R G Gamma_L G^dag] with matrix Gamma if want_caroli)
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
    if out is None:
        R = np.empty(shape, dtype=np.float64)
        T = np.empty(shape, dtype=np.float64)
    else:
        R, T = out
        if R.shape != shape or T.shape != shape or R.dtype != np.float64 or T.dtype != np.float64:
            raise ValueError("out arrays must be float64 of shape %s" % (shape,))
    resid = np.empty(shape[:3], dtype=np.float64) if want_resid else None
    total = np.empty(shape[:2], dtype=np.float64) if want_total else None
    caroli = np.empty(shape[:2], dtype=np.float64) if want_caroli else None

    rc = lib.tx_transmission_sweep(ptr(h), n_h, ptr(h1), ptr(params), ptr(tnn), n_t, n, M, n_ham,
                                   ptr(Es), len(Es), lead_mode, ptr(sigL), ptr(sigR), n_sig,
                                   TX_HOP_AUTO, ptr(src), None, n_src, ptr(R), ptr(T), ptr(resid), ptr(total), ptr(caroli))
    if rc == _lib.TX_ERR_SINGULAR:
        # numpy raises LinAlgError for an exactly singular E - H' (wfm/__init__.py:162)
        raise np.linalg.LinAlgError("Singular matrix")
    check(rc)
    ret = (R, T)
    if want_resid:
        ret += (resid,)
    if want_total:
        ret += (total,)
    if want_caroli:
        ret += (caroli,)
    return ret


def spin_flip_sums(T, n_elec_up=None):
    """Spin-resolved sums used by `run_wfm_gate.py:404-405,473-477`: channels [:n/2] carry electron
    spin up, [n/2:] spin down.  Returns (T_noflip, T_flip) [..., n_src]: transmitted weight with the
    electron spin equal / opposite to the source channel's."""
    n = T.shape[-1]
    half = n // 2 if n_elec_up is None else n_elec_up
    up = T[..., :half].sum(-1)
    dn = T[..., half:].sum(-1)
    src_up = np.arange(T.shape[-2]) < half
    return np.where(src_up, up, dn), np.where(src_up, dn, up)
