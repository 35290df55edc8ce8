"""Drop-in mirror of `transport.wfm` (reference: transport/wfm/__init__.py) for the T(E) hot path.

Same names, argument order, return shapes/dtypes and exception types as the reference's gen-3 API
(SURVEY.md §4, §8b); every function dispatches through the C ABI (include/transport_b200.h) into
sm_100a CUDA.  There is no numpy fallback: without the built library or without a GPU the calls
raise.  Host code here only validates arguments and shapes the way the reference does.
"""
import numpy as np

from . import _lib, leads, superblocks
from .sweep import transmission_sweep


def _is_str(converger, name):
    return isinstance(converger, str) and converger == name


def _lead_tables(h, tnn, Es, converger):
    """Sigma_L, Sigma_R [nE,n,n] for converger in {"g_RiceMele", (imE, conv_tol)} (device kernels)."""
    return leads.self_energies_batched(h, tnn, Es, converger)


def kernel(h, tnn, tnnn, tl, E, converger, Ajsigma,
           is_psi_jsigma, is_Rhat, all_debug=True, verbose=0):
    """Reflection and transmission probabilities per channel for one energy and one source vector.

    Mirrors `wfm.kernel` (transport/wfm/__init__.py:29-137).  Returns (Rs, Ts), two float64 arrays of
    length n_loc_dof; or the wavefunction psi[M, n] (complex) if `is_psi_jsigma`.
    """
    if not isinstance(h, np.ndarray): raise TypeError
    if not isinstance(tnn, np.ndarray): raise TypeError
    if not isinstance(tnnn, np.ndarray): raise TypeError
    for sigma in range(len(Ajsigma)):                      # :67-70 incident chemical potential is 0
        if Ajsigma[sigma] != 0:
            assert h[0, sigma, sigma] == 0
    assert isinstance(Ajsigma, np.ndarray)                 # :73
    assert len(Ajsigma) == np.shape(h[0])[0]               # :74
    if not len(tnn) + 1 == len(h): raise ValueError        # Hmat's checks, :179-180
    if not len(tnnn) + 2 == len(h): raise ValueError
    has_nnn = bool(np.any(tnnn))        # block-pentadiagonal H (:209-213): sites are paired into super-blocks
    n = np.shape(h[0])[0]
    Es = np.array([E], dtype=complex)

    need_sigma = all_debug or is_psi_jsigma or not _is_str(converger, "g_closed")
    if need_sigma:
        SigL, SigR = SelfEnergies(h, tnn, tnnn, E, converger, verbose=verbose)
    if all_debug:                                          # :82-86
        for Sig in (SigL, SigR):
            assert not np.any(Sig - np.diagflat(np.diagonal(Sig)))
        assert np.any(np.imag(np.diagonal(SigL)))

    if is_psi_jsigma:                                      # :98-99
        from . import green
        vL = -2 * np.imag(np.diagonal(SigL))
        if has_nnn:
            col = superblocks.green_column0_with_tnnn(h, tnn, tnnn, Es, SigL[None], SigR[None])[0]
        else:
            col = green.green_column0(h, tnn, Es, SigL[None], SigR[None])[0]      # [M, n, n]
        return 1j * np.dot(col, Ajsigma * vL)
    if is_Rhat:                                            # :103-105: formula is commented out upstream
        raise NameError("name 'Rhat_matrix' is not defined")

    src = np.asarray(Ajsigma, dtype=float)[None, :]
    if has_nnn:
        tables = (SigL[None], SigR[None]) if need_sigma else None
        R, T = superblocks.sweep_with_tnnn(h, tnn, tnnn, Es, src, converger, tables)
        R, T = R[None], T[None]
    elif _is_str(converger, "g_closed"):
        R, T = transmission_sweep(h, tnn, Es, "g_closed", sources=src)
    else:
        R, T = transmission_sweep(h, tnn, Es, "table", sources=src, sigL=SigL[None], sigR=SigR[None])
    Rs, Ts = R[0, 0, 0].copy(), T[0, 0, 0].copy()
    if all_debug:
        assert abs(np.sum(Rs) + np.sum(Ts) - 1) < 1e-10    # :135
    return Rs, Ts


def Green(h, tnn, tnnn, tl, E, converger, verbose=0) -> np.ndarray:
    """Full Green's function G[M, M, n, n] = inv(E - H') (transport/wfm/__init__.py:139-166),
    computed block-recursively on the device instead of by a dense inverse."""
    from . import green
    if not len(tnn) + 1 == len(h): raise ValueError
    if not len(tnnn) + 2 == len(h): raise ValueError
    SigL, SigR = SelfEnergies(h, tnn, tnnn, E, converger, verbose=verbose)
    Es = np.array([E], dtype=complex)
    if np.any(tnnn):
        return superblocks.green_full_with_tnnn(h, tnn, tnnn, Es, SigL[None], SigR[None])[0]
    return green.green_full(h, tnn, Es, SigL[None], SigR[None])[0]


def Hmat(h, tnn, tnnn) -> np.ndarray:
    """Dense block-pentadiagonal H[(M n), (M n)] (transport/wfm/__init__.py:168-215), assembled by a
    device scatter kernel."""
    from . import green
    if not len(tnn) + 1 == len(h): raise ValueError
    if not len(tnnn) + 2 == len(h): raise ValueError
    return green.hmat(h, tnn, tnnn)


def Hprime(h, tnn, tnnn, tlmag, E, converger, verbose=0) -> np.ndarray:
    """H' = H + Sigma_L (+) Sigma_R (transport/wfm/__init__.py:217-268)."""
    n = np.shape(h[0])[0]
    Hp = Hmat(h, tnn, tnnn)
    SigL, SigR = SelfEnergies(h, tnn, tnnn, E, converger, verbose=verbose)
    Hp[0:n, 0:n] += SigL
    Hp[-n:, -n:] += SigR
    for s in range(n):                                     # :249-251
        if abs(np.imag(SigL[s, s])) > 1e-10 and abs(np.imag(SigR[s, s])) > 1e-10:
            assert np.sign(np.imag(SigL[s, s])) == np.sign(np.imag(SigR[s, s]))
    return Hp


def SelfEnergies(h, tnn, tnnn, E, converger, verbose=0) -> tuple:
    """(Sigma_L, Sigma_R), each [n, n] complex (transport/wfm/__init__.py:270-304)."""
    if not (_is_str(converger, "g_closed") or _is_str(converger, "g_RiceMele") or isinstance(converger, tuple)):
        raise Exception("converger = " + str(converger) + " not supported")
    SigL, SigR = leads.self_energies_batched(np.asarray(h), np.asarray(tnn), np.array([E], dtype=complex), converger)
    return SigL[0], SigR[0]


def g_closed(diag, offdiag, E, inoutsign) -> np.ndarray:
    """Closed-form surface GF of a monatomic lead (transport/wfm/__init__.py:306-340)."""
    if np.shape(diag) != np.shape(offdiag) or np.shape(diag) == (): raise ValueError
    if inoutsign not in [1, -1]: raise ValueError
    if np.any(diag - np.diagflat(np.diagonal(diag))): raise Exception("Not diagonal:\n" + str(diag))
    return leads.surface_gf(diag, offdiag, np.array([E], dtype=complex), inoutsign, _lib.TX_GF_CLOSED)[0][0]


def g_RiceMele(diag, offdiag, E, inoutsign) -> np.ndarray:
    """Closed-form surface GF of a Rice-Mele lead (transport/wfm/__init__.py:352-414)."""
    if np.shape(diag) != np.shape(offdiag) or np.shape(diag) == (): raise ValueError
    if inoutsign not in [1, -1]: raise ValueError
    if len(diag) % 2 != 0: raise ValueError("diag is not Rice-Mele type")
    n_spin = len(diag) // 2
    offdiag_check = 1 * offdiag
    offdiag_check[n_spin:, :n_spin] = 0 * offdiag[n_spin:, :n_spin]
    if np.any(offdiag_check): raise ValueError("offdiag is not Rice-Mele type")
    u0 = (np.diagonal(diag)[:n_spin] + np.diagonal(diag)[n_spin:]) / 2
    for u0_ss in u0: assert abs(u0_ss) < 1e-10             # :385
    return leads.surface_gf(diag, offdiag, np.array([E], dtype=complex), inoutsign, _lib.TX_GF_RICEMELE)[0][0]


def g_iter(diag, offdiag, E, inoutsign, imE, conv_tol,
           conv_rep=5, min_iter=int(1e2), max_iter=int(1e5), full=False) -> np.ndarray:
    """Fixed-point surface GF with the reference's DOS stopping rule (transport/wfm/__init__.py:479-530)."""
    if np.shape(diag) != np.shape(offdiag): raise ValueError
    if inoutsign not in [1, -1]: raise ValueError
    if len(diag) > 2: raise NotImplementedError("Iterative diatomic and spin combined not supported")
    g, conv, nit, ok = leads.surface_gf(diag, offdiag, np.array([E], dtype=complex), inoutsign,
                                        _lib.TX_GF_FIXEDPOINT, imE=imE, conv_tol=conv_tol, conv_rep=conv_rep,
                                        min_iter=min_iter, max_iter=max_iter, full=True)
    if not ok[0]:
        raise Exception("g({:.6f}+{:.6f}j) convergence was {:.0e}, failed to meet {:.0e} tolerance after {:.0e} iterations"
                        .format(np.real(E), imE, conv[0], conv_tol, max_iter))
    if full: return g[0], conv[0], int(nit[0])
    return g[0]


def g_ith(diag, offdiag, E, inoutsign, imE, ith, g_prev) -> np.ndarray:
    """One fixed-point step of the surface GF (transport/wfm/__init__.py:532-551)."""
    if not isinstance(ith, int): raise TypeError
    return leads.g_ith(diag, offdiag, np.array([E], dtype=complex), inoutsign, imE, ith,
                       None if ith == 0 else np.asarray(g_prev)[None])[0]


def Rhat_matrix(h, tnn, tnnn, tl, E, converger="g_closed") -> np.ndarray:
    """Reflection operator  Rhat = 2i tl G_00 sin(ka_L) - 1  (n x n complex), the formula that is commented
    out in the reference (transport/wfm/__init__.py:103) and whose absence makes `kernel(..., is_Rhat=True)`
    raise NameError there (:105).  Extension for the gate-fidelity drivers (run_wfm_gate.py:465,603);
    `kernel` itself keeps the reference's behaviour.  ka_L follows `Hprime`'s convention (:253):
    ka_L = arccos((E - diag h[0]) / (-2 tl)), one value per channel (applied to the source column)."""
    from . import green
    if not len(tnn) + 1 == len(h): raise ValueError
    if not len(tnnn) + 2 == len(h): raise ValueError
    n = np.shape(h[0])[0]
    SigL, SigR = SelfEnergies(h, tnn, tnnn, E, converger)
    if np.any(tnnn):
        G00 = superblocks.green_column0_with_tnnn(h, tnn, tnnn, np.array([E], dtype=complex), SigL[None], SigR[None])[0, 0]
    else:
        G00 = green.green_column0(h, tnn, np.array([E], dtype=complex), SigL[None], SigR[None])[0, 0]
    ka_L = np.arccos((E - np.diagonal(h[0])) / (-2 * tl))
    return 2 * tl * complex(0, 1) * G00 * np.sin(ka_L)[None, :] - np.eye(n)


# ---- Rice-Mele band helpers: scalar host formulas, not on the device path --------------------
def _rm_uvw(diag, offdiag):
    u0 = np.sum(np.diagonal(diag)) / len(diag)
    assert len(diag) == 2
    return u0, (diag[0, 0] - diag[1, 1]) / 2, diag[0, -1], offdiag[-1, 0]


def dispersion_RiceMele(diag, offdiag, ks) -> np.ndarray:
    """E_pm(k), shape (2, len(ks)) (transport/wfm/__init__.py:416-429)."""
    u0, u, v, w = _rm_uvw(diag, offdiag)
    band = np.sqrt(u * u + v * v + w * w + 2 * v * w * np.cos(ks))
    return np.array([u0 - band, u0 + band])


def inverted_RiceMele(diag, offdiag, Es) -> np.ndarray:
    """k(E) (transport/wfm/__init__.py:431-444)."""
    u0, u, v, w = _rm_uvw(diag, offdiag)
    return np.arccos(1 / (2 * v * w) * ((Es - u0) ** 2 - u ** 2 - v ** 2 - w ** 2))


def bandedges_RiceMele(diag, offdiag) -> np.ndarray:
    """Four band edges (transport/wfm/__init__.py:446-459)."""
    u0, u, v, w = _rm_uvw(diag, offdiag)
    B = np.array([np.sqrt(u * u + (w + v) * (w + v)), np.sqrt(u * u + (w - v) * (w - v))])
    return np.array([u0 - np.max(B), u0 - np.min(B), u0 + np.min(B), u0 + np.max(B)])


def string_RiceMele(diag, offdiag, energies=True, tex=True) -> str:
    """Parameter label (transport/wfm/__init__.py:461-477)."""
    u0, u, v, w = _rm_uvw(diag, offdiag)
    edges = bandedges_RiceMele(diag, offdiag)
    out = "$u = {:.2f}, v ={:.2f}, w ={:.2f}$".format(u, v, w)
    if energies:
        out += ", $E_{gap}=" + "{:.2f}".format(edges[2] - edges[1]) + ", E_{band}=$" + "{:.2f}$".format(edges[1] - edges[0])
    if not tex:
        out = out.replace("$", "")
    return out
