"""
TEST INFRASTRUCTURE — not product code.  Nothing under transport_b200/ may import this.

Extended-precision restatements of the recursive evaluation in oracle/rgf_oracle.py (same recursion, same
conventions: `/root/reference/transport/wfm/__init__.py:203-207` block layout, `:246-247` self-energies on the end
blocks, `:337-340` g_closed, `:91,108-130` R/T), used as the TRUTH against which both the float64 numpy oracle
and the device kernels are measured where the problem is ill-conditioned (BASELINE config 3 at full length,
M = 10 002: a 1-ulp change of E moves T by ~1e-9, so float64-vs-float64 agreement cannot say who is right).

  * `transmission_closed_ld`  numpy long double (x87 80-bit: 64-bit mantissa, eps = 1.1e-19, ~2000x finer than
    float64), vectorised over energies; Gauss-Jordan with partial pivoting written out (numpy.linalg has no
    long-double path).
  * `transmission_closed_mp`  mpmath at `dps` decimal digits (default 50), one energy at a time: slow (seconds per
    hundred sites), used on a few energies to validate the long-double truth itself.

Real energies and real lead parameters only (every reference driver).
"""
import numpy as np

LD = np.longdouble
CLD = np.clongdouble


def _inv_ld(A):
    """Batched inverse of A [nE, n, n] (clongdouble): Gauss-Jordan with partial pivoting."""
    nE, n, _ = A.shape
    aug = np.concatenate([A.astype(CLD), np.broadcast_to(np.eye(n, dtype=CLD), (nE, n, n)).copy()], axis=2)
    rows = np.arange(nE)
    for k in range(n):
        piv = k + np.argmax(np.abs(aug[:, k:, k]), axis=1)
        tmp = aug[rows, piv].copy()
        aug[rows, piv] = aug[rows, k]
        aug[rows, k] = tmp
        aug[:, k] = aug[:, k] / aug[:, k, k][:, None]
        f = aug[:, :, k].copy()
        f[:, k] = 0
        aug = aug - f[:, :, None] * aug[:, k][:, None, :]
    return aug[:, :, n:]


def _g_closed_ld(E, eps, t):
    """wfm.g_closed for real E, eps, t (wfm/__init__.py:337-340) in long double; arrays broadcast."""
    E, eps, t = LD(1) * E, LD(1) * eps, LD(1) * t
    lam = (E - eps) / (2 * t)
    z = lam * lam - 1
    root = np.sqrt(np.abs(z))
    sg = np.sign(E - eps)
    re = np.where(z >= 0, lam / t - sg * root, lam / t)
    im = np.where(z >= 0, LD(0), -root)
    return re + 1j * im.astype(CLD)


def transmission_closed_ld(h, tnn, Es, sources):
    """(R, T) [nE, n_src, n] in long double, closed-form leads (lead blocks diagonal)."""
    Es = np.asarray(np.real(Es), dtype=LD)
    M, n = h.shape[0], h.shape[-1]
    hl, tl_ = h.astype(CLD), tnn.astype(CLD)
    eye = np.eye(n, dtype=CLD)
    epsL, epsR = np.real(np.diagonal(h[0])).astype(LD), np.real(np.diagonal(h[-1])).astype(LD)
    tL, tR = np.real(np.diagonal(tnn[0])).astype(LD), np.real(np.diagonal(tnn[-1])).astype(LD)
    gL = _g_closed_ld(Es[:, None], epsL[None, :], tL[None, :])          # [nE, n]
    gR = _g_closed_ld(Es[:, None], epsR[None, :], tR[None, :])
    sL, sR = (tL * tL)[None, :] * gL, (tR * tR)[None, :] * gR          # diagonal Sigma
    zE = Es[:, None, None].astype(CLD) * eye
    A = zE - hl[M - 1]
    A = A - sR[:, :, None] * eye
    g = _inv_ld(A)
    W = g.copy()
    for j in range(M - 2, -1, -1):
        t = tl_[j]
        td = np.conj(t).T
        A = zE - hl[j] - np.matmul(np.matmul(t, g), td)
        if j == 0:
            A = A - sL[:, :, None] * eye
        g = _inv_ld(A)
        W = np.matmul(np.matmul(W, td), g)
    vL, vR = -2 * np.imag(sL), -2 * np.imag(sR)
    sources = np.atleast_2d(np.asarray(sources, dtype=LD))
    nE = len(Es)
    R = np.zeros((nE, len(sources), n), dtype=LD)
    T = np.zeros((nE, len(sources), n), dtype=LD)
    for s, Asrc in enumerate(sources):
        Av = Asrc[None, :] * vL
        flux = np.sqrt(np.sum(Asrc[None, :] * Av, axis=1))
        r = (1j * np.einsum("eab,eb->ea", g, Av.astype(CLD)) - Asrc[None, :]) * np.sqrt(vL) / flux[:, None]
        t_ = 1j * np.einsum("eab,eb->ea", W, Av.astype(CLD)) * np.sqrt(vR) / flux[:, None]
        R[:, s] = np.real(r * np.conj(r))
        T[:, s] = np.real(t_ * np.conj(t_))
    return R, T


def transmission_closed_mp(h, tnn, E, sources, dps=50):
    """(R, T) [n_src, n] for ONE real energy with mpmath at `dps` digits (returned as float64 arrays plus the
    mp values' own unitarity residual)."""
    import mpmath as mp
    mp.mp.dps = dps
    M, n = h.shape[0], h.shape[-1]

    def tomp(a):
        return mp.matrix([[mp.mpc(mp.mpf(float(x.real)), mp.mpf(float(x.imag))) for x in row] for row in a])

    E = mp.mpf(float(np.real(E)))

    def gcl(eps, t):
        lam = (E - eps) / (2 * t)
        z = lam * lam - 1
        if z >= 0:
            return mp.mpc(lam / t - mp.sign(E - eps) * mp.sqrt(z), 0)
        return mp.mpc(lam / t, -mp.sqrt(-z))

    sL = [mp.mpf(float(tnn[0][c, c].real)) ** 2 * gcl(mp.mpf(float(h[0][c, c].real)), mp.mpf(float(tnn[0][c, c].real))) for c in range(n)]
    sR = [mp.mpf(float(tnn[-1][c, c].real)) ** 2 * gcl(mp.mpf(float(h[-1][c, c].real)), mp.mpf(float(tnn[-1][c, c].real))) for c in range(n)]
    eye = mp.eye(n)
    A = E * eye - tomp(h[M - 1])
    for c in range(n):
        A[c, c] -= sR[c]
    g = A ** -1
    W = g.copy()
    scalar_hop = all(np.array_equal(tnn[j], tnn[j][0, 0] * np.eye(n)) for j in range(M - 1))
    for j in range(M - 2, -1, -1):
        if scalar_hop:
            t = mp.mpc(float(tnn[j][0, 0].real), float(tnn[j][0, 0].imag))
            A = E * eye - tomp(h[j]) - (t * mp.conj(t)) * g
        else:
            t = tomp(tnn[j])
            A = E * eye - tomp(h[j]) - t * g * t.H
        if j == 0:
            for c in range(n):
                A[c, c] -= sL[c]
        g = A ** -1
        W = (mp.conj(t) * (W * g)) if scalar_hop else (W * t.H * g)
    vL = [-2 * s.imag for s in sL]
    vR = [-2 * s.imag for s in sR]
    sources = np.atleast_2d(np.asarray(sources, dtype=float))
    R = np.zeros((len(sources), n))
    T = np.zeros((len(sources), n))
    worst = mp.mpf(0)
    for s, Asrc in enumerate(sources):
        Av = [mp.mpf(float(Asrc[c])) * vL[c] for c in range(n)]
        flux = mp.sqrt(sum(mp.mpf(float(Asrc[c])) * Av[c] for c in range(n)))
        tot = mp.mpf(0)
        for a in range(n):
            r = (mp.mpc(0, 1) * sum(g[a, b] * Av[b] for b in range(n)) - mp.mpf(float(Asrc[a]))) * mp.sqrt(vL[a]) / flux
            t_ = mp.mpc(0, 1) * sum(W[a, b] * Av[b] for b in range(n)) * mp.sqrt(vR[a]) / flux
            rr, tt = abs(r) ** 2, abs(t_) ** 2
            R[s, a], T[s, a] = float(rr), float(tt)
            tot += rr + tt
        worst = max(worst, abs(tot - 1))
    return R, T, float(worst)
