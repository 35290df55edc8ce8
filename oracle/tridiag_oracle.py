"""
TEST INFRASTRUCTURE — not product code.  Nothing under transport_b200/ imports this module.

numpy restatement of the algorithm of transport_b200/csrc/bardeen.cu's eigen stage (the reference itself
calls dense LAPACK `np.linalg.eigh`, transport/bardeen/__init__.py:757-761 and :459): eigenvalues of a real
symmetric tridiagonal matrix below a cutoff by Sturm-count multisection (division-free minor recurrence on the
power-of-two scaled matrix), eigenvectors by the twisted
factorisation (forward LDL^T and backward UDU^T pivots, twist at the smallest |gamma|).  Used to validate
the algorithm against LAPACK on the CPU before it runs on the GPU.
"""
import numpy as np


def pow2_scale(d, e):
    """exact power of two bringing the Gershgorin norm into [0.5, 1)"""
    r = np.zeros(len(d)); r[:-1] += np.abs(e); r[1:] += np.abs(e)
    tnorm = float(np.max(np.abs(d) + r))
    if not (tnorm > 0) or not np.isfinite(tnorm):
        return 1.0, tnorm
    return float(np.ldexp(1.0, -np.frexp(tnorm)[1])), tnorm


def sturm_count(ds, e2s, xs):
    """number of eigenvalues < x: sign changes of the leading principal minors of the SCALED matrix, division free,
    pair renormalised every 8 sites, a zero minor takes the sign opposite to its predecessor; vectorised over xs."""
    xs = np.atleast_1d(np.asarray(xs, dtype=float))
    p0 = np.ones_like(xs)
    p1 = ds[0] - xs
    p1 = np.where(p1 == 0, -2.0 ** -100, p1)
    cnt = (p1 < 0).astype(int)
    for i in range(1, len(ds)):
        pn = (ds[i] - xs) * p1 - e2s[i - 1] * p0            # the device uses fma(d - x, p1, -(e2 p0))
        pn = np.where(pn == 0, -p1 * 2.0 ** -100, pn)
        cnt += np.signbit(pn) != np.signbit(p1)
        p0, p1 = p1, pn
        if i % 8 == 0:
            ex = np.maximum(np.frexp(p0)[1], np.frexp(p1)[1])
            p0 = np.ldexp(p0, -ex); p1 = np.ldexp(p1, -ex)
    return cnt


def eigvals_below(d, e, cutoff, lanes=32, max_pass=40):
    d = np.asarray(d, dtype=float); e = np.asarray(e, dtype=float)
    sc, tnorm = pow2_scale(d, e)
    ds, es = d * sc, e * sc
    e2s = es * es
    r = np.zeros(len(d)); r[:-1] += np.abs(es); r[1:] += np.abs(es)
    eps = np.finfo(float).eps
    pivmin = 4 * np.finfo(float).tiny
    glo = np.min(ds - r) - 2 * eps * len(d) - 2 * pivmin
    cut = min(max(float(cutoff), -4 * tnorm), 4 * tnorm) * sc
    nb = int(sturm_count(ds, e2s, cut)[0])
    out = np.empty(nb)
    for k in range(nb):
        lo, hi = glo, cut
        for _ in range(max_pass):
            xs = lo + (hi - lo) * ((np.arange(lanes) + 1.0) * (1.0 / 33.0))
            c = sturm_count(ds, e2s, xs)
            above = np.nonzero(c > k)[0]
            nlo, nhi = lo, hi
            if len(above) == 0:
                nlo = xs[-1]
            else:
                i = above[0]
                nhi = xs[i]
                if i > 0:
                    nlo = xs[i - 1]
            lo, hi = max(lo, nlo), min(hi, nhi)
            if hi - lo <= eps * (abs(lo) + abs(hi)) + eps * 1e-3 * tnorm * sc + 2 * pivmin:
                break
        out[k] = 0.5 * (lo + hi) / sc
    return out


def _pivots_from_minors(ds, e2s_next, lam, pivmin):
    """pivots D_i = p_i / p_{i-1} of the LDL^T factorisation of T - lam along one direction, from the division-free
    minor recurrence (pair renormalised every 8 sites); e2s_next[i] couples site i to site i+1 of this direction."""
    S = len(ds)
    D = np.empty(S)
    p0, p1 = 1.0, ds[0] - lam
    if abs(p1) < pivmin:
        p1 = -pivmin
    D[0] = p1
    for i in range(1, S):
        pn = (ds[i] - lam) * p1 - e2s_next[i - 1] * p0
        if pn == 0.0:
            pn = -p1 * 2.0 ** -100
        q = pn / p1
        if not abs(q) >= pivmin:
            q = -pivmin
        D[i] = q
        p0, p1 = p1, pn
        if i % 8 == 0:
            ex = max(np.frexp(p0)[1], np.frexp(p1)[1])
            p0, p1 = np.ldexp(p0, -ex), np.ldexp(p1, -ex)
    return D


def twisted_vector(d, e, lam):
    """eigenvector of eigenvalue lam by the twisted factorisation, on the power-of-two scaled matrix"""
    d = np.asarray(d, dtype=float); e = np.asarray(e, dtype=float)
    S = len(d)
    sc, _ = pow2_scale(d, e)
    ds, es, lam = d * sc, e * sc, lam * sc
    e2s = es * es
    pivmin = 4 * np.finfo(float).tiny
    Dp = _pivots_from_minors(ds, e2s, lam, pivmin)
    Dm = _pivots_from_minors(ds[::-1], e2s[::-1], lam, pivmin)[::-1]
    gamma = Dp + Dm - (ds - lam)
    r = int(np.argmin(np.abs(gamma)))
    z = np.zeros(S)
    z[r] = 1.0
    for i in range(r - 1, -1, -1):
        z[i] = -(es[i] / Dp[i]) * z[i + 1]
    for i in range(r, S - 1):
        z[i + 1] = -(es[i] / Dm[i + 1]) * z[i]
    return z / np.sqrt(np.dot(z, z))


def eigh_below(d, e, cutoff):
    lam = eigvals_below(d, e, cutoff)
    return lam, np.array([twisted_vector(np.asarray(d, float), np.asarray(e, float), l) for l in lam])
