"""
TEST INFRASTRUCTURE — not product code.  Nothing under transport_b200/ may import this.

CPU restatement (numpy) of the reference's wave-function-matching T(E) path,
`/root/reference/transport/wfm/__init__.py` plus the one helper it takes from
`/root/reference/transport/tdfci/utils.py`.  Every function cites the reference
lines it follows.  The restatement keeps the reference's *algorithm*: per point it
assembles the dense (M n) x (M n) matrix H' = H + Sigma, inverts E - H' densely with
LAPACK and reads two blocks of the inverse.  Two flavours are provided:

  * ``*_loops``  functions keep the reference's Python loop nests (this is where the
    reference spends 60-90 % of its time), so that timing them gives a
    faithful "port" CPU baseline;
  * the plain functions are the same arithmetic with the index shuffles done by
    numpy reshapes, used by the parity tests where the loops would be too slow.

Pinning: tests/golden/*.npz were produced by the *unmodified* reference imported in the
build container (oracle/make_golden.py, oracle/ref_loader.py); tests/test_oracle.py
checks this file against every one of them (bit-for-bit up to LAPACK round-off, 1e-13
relative).  Parity status: PINNED against live reference output for configs 1, 2, 3
(down-scaled) and the surface-GF functions; config 4 (n=64 leads) is "parity unpinned"
by the reference itself (g_iter refuses n>2, wfm/__init__.py:502) — see DESIGN.md.
"""
import numpy as np


# ----------------------------------------------------------------------------------------
# index helpers  (transport/tdfci/utils.py:290-316, :318-345)
# ----------------------------------------------------------------------------------------
def mat_2d_to_4d_loops(mat, n_loc_dof):
    """(i*n+a, j*n+b) -> (i, j, a, b) with the reference's 4-deep loop (utils.py:303-314)."""
    if not isinstance(mat, np.ndarray):
        raise TypeError
    if len(mat) % n_loc_dof != 0:
        raise ValueError
    n_sites = len(mat) // n_loc_dof
    out = np.zeros((n_sites, n_sites, n_loc_dof, n_loc_dof), dtype=mat.dtype)
    for si in range(n_sites):
        for sj in range(n_sites):
            for a in range(n_loc_dof):
                for b in range(n_loc_dof):
                    out[si, sj, a, b] = mat[si * n_loc_dof + a, sj * n_loc_dof + b]
    return out


def mat_2d_to_4d(mat, n_loc_dof):
    """Same mapping as `mat_2d_to_4d_loops` via reshape/transposition."""
    if not isinstance(mat, np.ndarray):
        raise TypeError
    if len(mat) % n_loc_dof != 0:
        raise ValueError
    s = len(mat) // n_loc_dof
    return np.ascontiguousarray(mat.reshape(s, n_loc_dof, s, n_loc_dof).transpose(0, 2, 1, 3))


def mat_4d_to_2d(mat):
    """(i, j, a, b) -> (i*n+a, j*n+b)  (utils.py:318-345)."""
    if not isinstance(mat, np.ndarray):
        raise TypeError
    if mat.shape[0] != mat.shape[1] or mat.shape[2] != mat.shape[3]:
        raise ValueError
    s, n = mat.shape[0], mat.shape[-1]
    return np.ascontiguousarray(mat.transpose(0, 2, 1, 3).reshape(s * n, s * n))


# ----------------------------------------------------------------------------------------
# surface Green's functions
# ----------------------------------------------------------------------------------------
def g_closed(diag, offdiag, E, inoutsign):
    """Monatomic semi-infinite chain, per channel (wfm/__init__.py:306-340, live branch :335-340).

    lam = (E - eps)/(2 t);  g = Re(lam/t) - sign(E-eps) |Re sqrt(lam^2-1)| - i |Im sqrt(lam^2-1)|.
    `inoutsign` is validated but not used by the live branch.
    """
    if np.shape(diag) != np.shape(offdiag) or np.shape(diag) == ():
        raise ValueError
    if inoutsign not in [1, -1]:
        raise ValueError
    if np.any(diag - np.diagflat(np.diagonal(diag))):
        raise Exception("Not diagonal:\n" + str(diag))
    eps = np.diagonal(diag)
    t = np.diagonal(offdiag)
    lam = (E - eps) / (2 * t)
    first = lam / t
    root = np.lib.scimath.sqrt(lam * lam - 1)
    return np.diagflat(np.real(first) - np.sign(E - eps) * abs(np.real(root)) - 1j * abs(np.imag(root)))


def g_RiceMele(diag, offdiag, E, inoutsign):
    """Diatomic Rice-Mele closed form per spin channel (wfm/__init__.py:352-414)."""
    if np.shape(diag) != np.shape(offdiag) or np.shape(diag) == ():
        raise ValueError
    if inoutsign not in [1, -1]:
        raise ValueError
    if len(diag) % 2 != 0:
        raise ValueError("diag is not Rice-Mele type")
    ns = len(diag) // 2
    chk = 1 * offdiag
    chk[ns:, :ns] = 0 * offdiag[ns:, :ns]
    if np.any(chk):
        raise ValueError("offdiag is not Rice-Mele type")
    d = np.diagonal(diag)
    u = (d[:ns] - d[ns:]) / 2                     # :379
    v = np.diagonal(diag[:ns, ns:])               # :380
    w = np.diagonal(offdiag[ns:, :ns])            # :381
    u0 = (d[:ns] + d[ns:]) / 2                    # :384
    for val in u0:
        assert abs(val) < 1e-10                   # :385
    x = E - u0
    Z = (x + u) * (x - u) + (w + v) * (w - v)     # :394-395
    pref = 1 / (2 * w * w * (E - u0 - u))         # :398
    root = np.lib.scimath.sqrt(4 * w * w * (u + E - u0) * (u - E + u0) + np.power(Z, 2))  # :399-400
    root = np.array(root, dtype=complex)
    for s in range(ns):                           # :401-402, retarded: Im g < 0
        if np.imag(root[s] * pref[s]) > 0:
            root[s] = -root[s]
    g = pref * (Z + root)                         # :403
    gmat = np.zeros(np.shape(diag), dtype=complex)
    if inoutsign == -1:                           # :407-410 left lead -> B,B block
        gmat[ns:, ns:] = np.diagflat(g)
    else:                                         # :411-413 right lead -> A,A block
        gmat[:ns, :ns] = np.diagflat(g)
    return gmat


def g_ith(diag, offdiag, E, inoutsign, imE, ith, g_prev):
    """One fixed-point step g <- (z - h00 - t g t^dag)^-1, z = Re E + i|imE| (wfm/__init__.py:532-551)."""
    if not isinstance(ith, int):
        raise TypeError
    eye = np.eye(len(diag))
    z = complex(np.real(E), abs(imE))             # :541
    if ith == 0:
        return np.linalg.inv(eye * z - diag)      # :544
    if inoutsign == 1:                            # :547-548
        tgt = np.matmul(offdiag, np.matmul(g_prev, np.conj(offdiag.T)))
    elif inoutsign == -1:                         # :549-550
        tgt = np.matmul(np.conj(offdiag.T), np.matmul(g_prev, offdiag))
    return np.linalg.inv(eye * z - diag - tgt)    # :551


def g_iter(diag, offdiag, E, inoutsign, imE, conv_tol,
           conv_rep=5, min_iter=int(1e2), max_iter=int(1e5), full=False):
    """Fixed-point iteration with the reference's surface-DOS stopping rule (wfm/__init__.py:479-530).

    `n_limit` mirrors the reference's refusal of blocks larger than 2 (:502).
    """
    if np.shape(diag) != np.shape(offdiag):
        raise ValueError
    if inoutsign not in [1, -1]:
        raise ValueError
    if len(diag) > 2:
        raise NotImplementedError("Iterative diatomic and spin combined not supported")
    return g_iter_anysize(diag, offdiag, E, inoutsign, imE, conv_tol, conv_rep, min_iter, max_iter, full)


def g_iter_anysize(diag, offdiag, E, inoutsign, imE, conv_tol,
                   conv_rep=5, min_iter=int(1e2), max_iter=int(1e5), full=False):
    """`g_iter` without the n<=2 guard (same loop, :504-530); used to test larger blocks."""
    g = np.zeros_like(diag)
    dos = np.zeros((max_iter,), dtype=float)
    hit = np.zeros((max_iter,), dtype=int)
    check = np.nan
    for ith in range(max_iter):
        g = g_ith(diag, offdiag, E, inoutsign, imE, ith, g)
        if inoutsign == 1:
            dos[ith] = (-1 / np.pi) * np.imag(g)[0, 0]        # :516
        else:
            dos[ith] = (-1 / np.pi) * np.imag(g)[-1, -1]      # :517
        if ith > min_iter:                                     # :520
            check = abs((dos[ith] - dos[ith - 1]) / dos[ith])
            if check < conv_tol:
                hit[ith] = 1
            if np.sum(hit[ith + 1 - conv_rep:ith + 1]) == conv_rep:   # :523
                if full:
                    return g, check, ith
                return g
    raise Exception("g({:.6f}+{:.6f}j) convergence was {:.0e}, failed to meet {:.0e} tolerance after {:.0e} iterations"
                    .format(np.real(E), imE, check, conv_tol, max_iter))


# ----------------------------------------------------------------------------------------
# self energies, Hamiltonian assembly, Green's function
# ----------------------------------------------------------------------------------------
def SelfEnergies(h, tnn, tnnn, E, converger, verbose=0):
    """Sigma_L = t0^dag g_L t0, Sigma_R = tN g_R tN^dag (wfm/__init__.py:270-304)."""
    gL_args = (h[0], tnn[0], E, -1)               # :283
    gR_args = (h[-1], tnn[-1], E, 1)              # :284
    if isinstance(converger, str) and converger == "g_closed":
        gL, gR = g_closed(*gL_args), g_closed(*gR_args)
    elif isinstance(converger, str) and converger == "g_RiceMele":
        gL, gR = g_RiceMele(*gL_args), g_RiceMele(*gR_args)
    elif isinstance(converger, tuple):
        gL, gR = g_iter(*gL_args, *converger), g_iter(*gR_args, *converger)
    else:
        raise Exception("converger = " + str(converger) + " not supported")
    assert np.shape(gL) == np.shape(tnn[0])
    SigL = np.matmul(np.conj(tnn[0]).T, np.matmul(gL, tnn[0]))        # :301
    SigR = np.matmul(tnn[-1], np.matmul(gR, np.conj(tnn[-1]).T))      # :302
    return SigL, SigR


def Hmat_loops(h, tnn, tnnn):
    """Dense block-pentadiagonal H with the reference's 4-deep loop (wfm/__init__.py:168-215)."""
    if not len(tnn) + 1 == len(h):
        raise ValueError
    if not len(tnnn) + 2 == len(h):
        raise ValueError
    M = len(h)
    n = np.shape(h[0])[0]
    H = np.zeros((n * M, n * M), dtype=complex)
    for si in range(M):
        for sj in range(M):
            for a in range(n):
                for b in range(n):
                    row, col = si * n + a, sj * n + b
                    if si == sj:
                        H[row, col] = h[si][a, b]
                    elif si == sj + 1:
                        H[row, col] = (np.conj(tnn[sj]).T)[a, b]
                    elif si + 1 == sj:
                        H[row, col] = tnn[si][a, b]
                    elif si == sj + 2:
                        H[row, col] = (np.conj(tnnn[sj]).T)[a, b]
                    elif si + 2 == sj:
                        H[row, col] = tnnn[si][a, b]
    return H


def Hmat(h, tnn, tnnn):
    """Same matrix as `Hmat_loops`, block-sliced (upper = tnn[j], lower = tnn[j]^dag; :203-213)."""
    if not len(tnn) + 1 == len(h):
        raise ValueError
    if not len(tnnn) + 2 == len(h):
        raise ValueError
    M = len(h)
    n = np.shape(h[0])[0]
    H = np.zeros((n * M, n * M), dtype=complex)
    for j in range(M):
        H[j * n:(j + 1) * n, j * n:(j + 1) * n] = h[j]
    for j in range(M - 1):
        H[j * n:(j + 1) * n, (j + 1) * n:(j + 2) * n] = tnn[j]
        H[(j + 1) * n:(j + 2) * n, j * n:(j + 1) * n] = np.conj(tnn[j]).T
    for j in range(M - 2):
        H[j * n:(j + 1) * n, (j + 2) * n:(j + 3) * n] = tnnn[j]
        H[(j + 2) * n:(j + 3) * n, j * n:(j + 1) * n] = np.conj(tnnn[j]).T
    return H


def Hprime(h, tnn, tnnn, tlmag, E, converger, verbose=0, _hmat=Hmat):
    """H' = H + Sigma_L (+) Sigma_R on the corner blocks (wfm/__init__.py:217-268)."""
    n = np.shape(h[0])[0]
    Hp = _hmat(h, tnn, tnnn)                                        # :240
    SigL, SigR = SelfEnergies(h, tnn, tnnn, E, converger, verbose)  # :243 (second evaluation)
    Hp[0:n, 0:n] += SigL                                            # :246
    Hp[-n:, -n:] += SigR                                            # :247
    for s in range(n):                                              # :249-251
        if abs(np.imag(SigL[s, s])) > 1e-10 and abs(np.imag(SigR[s, s])) > 1e-10:
            assert np.sign(np.imag(SigL[s, s])) == np.sign(np.imag(SigR[s, s]))
    return Hp


def Green(h, tnn, tnnn, tl, E, converger, verbose=0, _loops=False):
    """G = inv(E - H') as [M, M, n, n] (wfm/__init__.py:139-166)."""
    n = np.shape(h[0])[0]
    Hp = Hprime(h, tnn, tnnn, tl, E, converger, verbose, _hmat=Hmat_loops if _loops else Hmat)
    G = np.linalg.inv(E * np.eye(*np.shape(Hp)) - Hp)               # :162
    return mat_2d_to_4d_loops(G, n) if _loops else mat_2d_to_4d(G, n)   # :165


def kernel(h, tnn, tnnn, tl, E, converger, Ajsigma, is_psi_jsigma, is_Rhat,
           all_debug=True, verbose=0, _loops=False):
    """Per-channel reflection / transmission for one energy and one source vector
    (wfm/__init__.py:29-137, gen-3 signature)."""
    if not isinstance(h, np.ndarray):
        raise TypeError
    if not isinstance(tnn, np.ndarray):
        raise TypeError
    if not isinstance(tnnn, np.ndarray):
        raise TypeError
    for s in range(len(Ajsigma)):                                   # :67-70
        if Ajsigma[s] != 0:
            assert h[0, s, s] == 0
    assert isinstance(Ajsigma, np.ndarray)                          # :73
    assert len(Ajsigma) == np.shape(h[0])[0]                        # :74
    N = len(h) - 2
    n = np.shape(h[0])[0]

    SigL, SigR = SelfEnergies(h, tnn, tnnn, E, converger, verbose=verbose)   # :81
    if all_debug:                                                   # :82-86
        for S in (SigL, SigR):
            assert not np.any(S - np.diagflat(np.diagonal(S)))
        assert np.any(np.imag(np.diagonal(SigL)))
    vL = -2 * np.imag(np.diagonal(SigL))                            # :91
    vR = -2 * np.imag(np.diagonal(SigR))

    G = Green(h, tnn, tnnn, tl, E, converger, verbose=verbose, _loops=_loops)   # :95
    psi = 1j * np.dot(G[:, 0], Ajsigma * vL)                        # :98
    if is_psi_jsigma:
        return psi
    if is_Rhat:                                                     # :103-105 (formula commented out upstream)
        raise NameError("name 'Rhat_matrix' is not defined")

    i_flux = np.sqrt(np.dot(Ajsigma, Ajsigma * np.real(vL)))        # :108
    Rs = np.zeros(n, dtype=float)
    Ts = np.zeros(n, dtype=float)
    for s in range(n):                                              # :114-130
        r_el = (1j * np.dot(G[0, 0, s], Ajsigma * vL) - Ajsigma[s]) * np.sqrt(np.real(vL[s])) / i_flux
        Rc = r_el * np.conjugate(r_el)
        assert abs(np.imag(Rc)) < 1e-10 or not np.isfinite(Rc)
        Rs[s] = np.real(Rc)
        t_el = 1j * np.dot(G[N + 1, 0, s], Ajsigma * vL) * np.sqrt(np.real(vR[s])) / i_flux
        Tc = t_el * np.conjugate(t_el)
        assert abs(np.imag(Tc)) < 1e-10 or not np.isfinite(Tc)
        Ts[s] = np.real(Tc)
    if all_debug:
        assert abs(np.sum(Rs) + np.sum(Ts) - 1) < 1e-10             # :135
    return Rs, Ts


def kernel_loops(h, tnn, tnnn, tl, E, converger, Ajsigma, is_psi_jsigma, is_Rhat,
                 all_debug=True, verbose=0):
    """`kernel` with the reference's Python loop nests kept (the CPU-baseline "port")."""
    return kernel(h, tnn, tnnn, tl, E, converger, Ajsigma, is_psi_jsigma, is_Rhat,
                  all_debug=all_debug, verbose=verbose, _loops=True)


# ----------------------------------------------------------------------------------------
# spin-resolved sums of the drivers  (run_wfm_gate.py:404-405, :473-477)
# ----------------------------------------------------------------------------------------
def spin_flip_sums(X, n_elec_up=None, sources=None):
    """Sums of a per-channel probability array X[..., n_src, n] (T or R) over the electron-spin sectors the
    drivers split the channels into (`np.array(range(n_loc_dof)) < n_mol_dof`, run_wfm_gate.py:473-477):
    channels [:n_elec_up] carry electron spin up (default n//2).  Returns (same, other) [..., n_src]: the weight
    leaving in the source's own sector / in the other one.  A source's sector is the one holding the larger share
    of its incident weight (`sources` [n_src, n]; default: one-hot sources in channel order)."""
    n = X.shape[-1]
    half = n // 2 if n_elec_up is None else n_elec_up
    up = X[..., :half].sum(-1)
    dn = X[..., half:].sum(-1)
    if sources is None:
        src_up = np.arange(X.shape[-2]) < half
    else:
        A2 = np.asarray(sources, dtype=float) ** 2
        src_up = A2[:, :half].sum(-1) >= A2[:, half:].sum(-1)
    return np.where(src_up, up, dn), np.where(src_up, dn, up)
