"""Lead surface Green's functions and self-energies (device kernels K1, csrc/surface_gf.cu).

Batched-over-energy counterparts of `wfm.g_closed`, `g_RiceMele`, `g_iter`, `g_ith` and
`SelfEnergies` (transport/wfm/__init__.py:270-551); `transport_b200.wfm` wraps them with the
reference's scalar signatures.
"""
import numpy as np

from . import _lib
from ._lib import as_c128, check, ptr

#: Sancho-Rubio defaults: |alpha|,|beta| below this ends the decimation
SANCHO_TOL = 1e-14
SANCHO_MAX_ITER = 200


def surface_gf(diag, offdiag, Es, inoutsign, mode, imE=0.0, conv_tol=0.0, conv_rep=5, min_iter=100,
               max_iter=100000, full=False, want_sigma=False):
    """g[nE,n,n] for every energy.  Returns (g,) or (g, conv, n_iter, converged) if `full`;
    with want_sigma the first element is (g, sigma)."""
    lib = _lib.load()
    diag, offdiag = as_c128(diag), as_c128(offdiag)
    Es = as_c128(np.atleast_1d(Es))
    n, nE = diag.shape[0], len(Es)
    if diag.shape != (n, n) or offdiag.shape != (n, n):
        raise ValueError
    g = np.empty((nE, n, n), dtype=np.complex128)
    sig = np.empty((nE, n, n), dtype=np.complex128) if want_sigma else None
    nit = np.zeros(nE, dtype=np.int32)
    conv = np.zeros(nE, dtype=np.float64)
    ok = np.ones(nE, dtype=np.int32)
    check(lib.tx_surface_gf(ptr(diag), ptr(offdiag), n, ptr(Es), nE, int(inoutsign), int(mode), float(imE),
                            float(conv_tol), int(conv_rep), int(min_iter), int(max_iter),
                            ptr(g), ptr(sig), ptr(nit), ptr(conv), ptr(ok)))
    first = (g, sig) if want_sigma else g
    if full:
        return first, conv, nit, ok.astype(bool)
    return (first,)


def surface_gf_pair(diag, offdiag, Es, imE=0.0, conv_tol=SANCHO_TOL, max_iter=SANCHO_MAX_ITER, want_g=False):
    """Both semi-infinite continuations of one periodic lead (diag, offdiag) from a single Sancho-Rubio decimation
    (tx_surface_gf_pair): returns (Sigma_L, Sigma_R, conv, n_iter, converged), each Sigma [nE,n,n]; with want_g
    (g_L, g_R, Sigma_L, Sigma_R, conv, n_iter, converged)."""
    lib = _lib.load()
    diag, offdiag = as_c128(diag), as_c128(offdiag)
    Es = as_c128(np.atleast_1d(Es))
    n, nE = diag.shape[0], len(Es)
    if diag.shape != (n, n) or offdiag.shape != (n, n):
        raise ValueError
    sigL = np.empty((nE, n, n), dtype=np.complex128)
    sigR = np.empty_like(sigL)
    gL = np.empty_like(sigL) if want_g else None
    gR = np.empty_like(sigL) if want_g else None
    nit = np.zeros(nE, dtype=np.int32)
    conv = np.zeros(nE, dtype=np.float64)
    ok = np.ones(nE, dtype=np.int32)
    check(lib.tx_surface_gf_pair(ptr(diag), ptr(offdiag), n, ptr(Es), nE, float(imE), float(conv_tol), int(max_iter),
                                 ptr(gL), ptr(sigL), ptr(gR), ptr(sigR), ptr(nit), ptr(conv), ptr(ok)))
    if want_g:
        return gL, gR, sigL, sigR, conv, nit, ok.astype(bool)
    return sigL, sigR, conv, nit, ok.astype(bool)


def g_ith(diag, offdiag, Es, inoutsign, imE, ith, g_prev):
    """One fixed-point step for every energy; g_prev [nE,n,n] (ignored for ith == 0)."""
    lib = _lib.load()
    diag, offdiag = as_c128(diag), as_c128(offdiag)
    Es = as_c128(np.atleast_1d(Es))
    n, nE = diag.shape[0], len(Es)
    if inoutsign not in (1, -1):
        raise ValueError
    gp = None if g_prev is None else as_c128(g_prev)
    out = np.empty((nE, n, n), dtype=np.complex128)
    check(lib.tx_g_ith(ptr(diag), ptr(offdiag), n, ptr(Es), nE, int(inoutsign), float(imE), int(ith), ptr(gp), ptr(out)))
    return out


def _mode_of(converger):
    if isinstance(converger, str):
        if converger == "g_closed":
            return _lib.TX_GF_CLOSED, {}
        if converger == "g_RiceMele":
            return _lib.TX_GF_RICEMELE, {}
        if converger == "sancho":
            return _lib.TX_GF_SANCHO, dict(imE=0.0, conv_tol=SANCHO_TOL, max_iter=SANCHO_MAX_ITER)
        if converger == "separable":
            return _lib.TX_GF_SEPARABLE, {}
    if isinstance(converger, tuple):
        if len(converger) == 2 and converger[0] == "separable":
            return _lib.TX_GF_SEPARABLE, dict(imE=float(converger[1]))
        if len(converger) >= 2 and converger[0] == "sancho":
            eta = float(converger[1])
            tol = float(converger[2]) if len(converger) > 2 else SANCHO_TOL
            return _lib.TX_GF_SANCHO, dict(imE=eta, conv_tol=tol, max_iter=SANCHO_MAX_ITER)
        # the reference's iterative converger: (imE, conv_tol[, conv_rep, min_iter, max_iter])
        names = ("imE", "conv_tol", "conv_rep", "min_iter", "max_iter")
        return _lib.TX_GF_FIXEDPOINT, {k: v for k, v in zip(names, converger)}
    raise Exception("converger = " + str(converger) + " not supported")


def self_energies_batched(h, tnn, Es, converger):
    """(Sigma_L, Sigma_R), each [nE,n,n]: left lead continues (h[0], tnn[0]), right lead (h[-1], tnn[-1])
    (transport/wfm/__init__.py:283-302).  Host-side argument checks are the reference's."""
    mode, kw = _mode_of(converger)
    h0, t0, hN, tN = np.asarray(h[0]), np.asarray(tnn[0]), np.asarray(h[-1]), np.asarray(tnn[-1])
    if mode == _lib.TX_GF_CLOSED:
        for d in (h0, hN):
            if np.any(d - np.diagflat(np.diagonal(d))):          # :327
                raise Exception("Not diagonal:\n" + str(d))
    elif mode == _lib.TX_GF_RICEMELE:
        for d, o in ((h0, t0), (hN, tN)):
            if len(d) % 2 != 0:
                raise ValueError("diag is not Rice-Mele type")
            ns = len(d) // 2
            chk = 1 * o
            chk[ns:, :ns] = 0 * o[ns:, :ns]
            if np.any(chk):
                raise ValueError("offdiag is not Rice-Mele type")
            u0 = (np.diagonal(d)[:ns] + np.diagonal(d)[ns:]) / 2
            for val in u0:
                assert abs(val) < 1e-10                          # :385
    elif mode == _lib.TX_GF_FIXEDPOINT:
        if len(h0) > 2:                                          # :502
            raise NotImplementedError("Iterative diatomic and spin combined not supported")
    if mode == _lib.TX_GF_SANCHO and np.array_equal(h0, hN) and np.array_equal(t0, tN):
        # both leads the same material: one decimation carries both surfaces (tx_surface_gf_pair)
        sigL, sigR, conv, nit, ok = surface_gf_pair(h0, t0, Es, **kw)
        if not np.all(ok):
            bad = int(np.argmin(ok))
            raise Exception("Sancho-Rubio decimation did not converge at E index %d (residual %.1e)" % (bad, conv[bad]))
        return sigL, sigR
    out = []
    for (d, o, side) in ((h0, t0, -1), (hN, tN, 1)):
        (g_sig, conv, nit, ok) = surface_gf(d, o, Es, side, mode, full=True, want_sigma=True, **kw)
        if not np.all(ok):
            bad = int(np.argmin(ok))
            if mode == _lib.TX_GF_FIXEDPOINT:
                raise Exception("g({:.6f}+{:.6f}j) convergence was {:.0e}, failed to meet {:.0e} tolerance after {:.0e} iterations"
                                .format(np.real(np.atleast_1d(Es)[bad]), kw.get("imE", 0.0), conv[bad],
                                        kw.get("conv_tol", 0.0), kw.get("max_iter", 100000)))
            raise Exception("Sancho-Rubio decimation did not converge at E index %d (residual %.1e)" % (bad, conv[bad]))
        out.append(g_sig[1])
    return out[0], out[1]
