"""
TEST INFRASTRUCTURE — not product code.  Nothing under transport_b200/ imports this module.

numpy restatement of the algorithm of transport_b200/csrc/bardeen.cu's eigen stage (the reference itself
calls dense LAPACK `np.linalg.eigh`, transport/bardeen/__init__.py:757-761 and :459): eigenvalues of a real
symmetric tridiagonal matrix below a cutoff by Sturm-count multisection, eigenvectors by the twisted
factorisation (forward LDL^T and backward UDU^T pivots, twist at the smallest |gamma|).  Used to validate
the algorithm against LAPACK on the CPU before it runs on the GPU.
"""
import numpy as np


def sturm_count(d, e2, x, pivmin):
    """number of eigenvalues < x, vectorised over the shifts x."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    q = d[0] - x
    q = np.where(np.abs(q) < pivmin, -pivmin, q)
    cnt = (q < 0).astype(int)
    for i in range(1, len(d)):
        q = (d[i] - x) - e2[i - 1] / q
        q = np.where(np.abs(q) < pivmin, -pivmin, q)
        cnt += q < 0
    return cnt


def eigvals_below(d, e, cutoff, lanes=32, max_pass=20):
    d = np.asarray(d, dtype=float); e = np.asarray(e, dtype=float)
    e2 = e * e
    pivmin = np.finfo(float).tiny * max(1.0, e2.max() if len(e2) else 1.0)
    r = np.zeros(len(d)); r[:-1] += np.abs(e); r[1:] += np.abs(e)
    glo = np.min(d - r); glo -= 1e-12 * max(1.0, abs(glo))
    nb = int(sturm_count(d, e2, cutoff, pivmin)[0])
    out = np.empty(nb)
    for k in range(nb):
        lo, hi = glo, float(cutoff)
        for _ in range(max_pass):
            xs = lo + (hi - lo) * (np.arange(lanes) + 1.0) / (lanes + 1.0)
            c = sturm_count(d, e2, xs, pivmin)
            above = np.nonzero(c > k)[0]
            if len(above) == 0:
                lo = xs[-1]
            else:
                i = above[0]
                hi = xs[i]
                if i > 0:
                    lo = xs[i - 1]
            if hi - lo <= 2 * np.finfo(float).eps * max(abs(lo), abs(hi)):
                break
        out[k] = 0.5 * (lo + hi)
    return out


def twisted_vector(d, e, lam):
    S = len(d)
    e2 = e * e
    pivmin = np.finfo(float).tiny * max(1.0, e2.max() if len(e2) else 1.0)
    Dp = np.empty(S); Dm = np.empty(S)
    q = d[0] - lam
    for i in range(S):
        if i:
            q = (d[i] - lam) - e2[i - 1] / q
        if abs(q) < pivmin:
            q = -pivmin
        Dp[i] = q
    q = d[S - 1] - lam
    for i in range(S - 1, -1, -1):
        if i < S - 1:
            q = (d[i] - lam) - e2[i] / q
        if abs(q) < pivmin:
            q = -pivmin
        Dm[i] = q
    gamma = Dp + Dm - (d - lam)
    r = int(np.argmin(np.abs(gamma)))
    z = np.zeros(S)
    z[r] = 1.0
    for i in range(r - 1, -1, -1):
        z[i] = -(e[i] / Dp[i]) * z[i + 1]
    for i in range(r, S - 1):
        z[i + 1] = -(e[i] / Dm[i + 1]) * z[i]
    return z / np.sqrt(np.dot(z, z))


def eigh_below(d, e, cutoff):
    lam = eigvals_below(d, e, cutoff)
    return lam, np.array([twisted_vector(np.asarray(d, float), np.asarray(e, float), l) for l in lam])
