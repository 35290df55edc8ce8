"""
TEST INFRASTRUCTURE — not product code.  Nothing under transport_b200/ may import this.

numpy restatement of the *recursive* (block-tridiagonal) evaluation of the same two
Green's-function blocks the reference reads from its dense inverse
(`/root/reference/transport/wfm/__init__.py:162` dense inverse, `:116`, `:124` the two
blocks G[0,0], G[N+1,0]).  It exists because the reference's dense path cannot run
BASELINE config 3 (M = 10 002 blocks -> a 102 GB matrix) and refuses iterative leads
for n > 2 (config 4).  It is validated against `oracle/wfm_oracle.py` (itself pinned to
live reference output) in tests/test_oracle.py at every size the dense path can do;
at the sizes it cannot, parity rests on this file + the unitarity identity R+T=1.

Conventions follow the reference: tnn[j] = H_{j,j+1} (upper), lower block = tnn[j]^dagger
(`wfm/__init__.py:203-207`); Sigma_L/R added to blocks 0 and M-1 (`:246-247`);
v = -2 Im diag Sigma (`:91`); R/T formulas `:108-130`.

All functions are vectorised over a leading energy axis.
"""
import numpy as np

from . import wfm_oracle


def closed_selfenergies(h, tnn, Es):
    """Sigma_L, Sigma_R [nE, n, n] with the reference's g_closed (wfm/__init__.py:283-302, :337-340)."""
    Es = np.asarray(Es, dtype=complex)
    n = h.shape[-1]
    SigL = np.zeros((len(Es), n, n), dtype=complex)
    SigR = np.zeros_like(SigL)
    for i, E in enumerate(Es):
        gL = wfm_oracle.g_closed(h[0], tnn[0], E, -1)
        gR = wfm_oracle.g_closed(h[-1], tnn[-1], E, 1)
        SigL[i] = np.conj(tnn[0]).T @ gL @ tnn[0]
        SigR[i] = tnn[-1] @ gR @ np.conj(tnn[-1]).T
    return SigL, SigR


def rgf_blocks(h, tnn, Es, SigL, SigR):
    """G_00 and G_{M-1,0} for every energy by one right-to-left sweep.

    g_j = (E - h_j - t_j g_{j+1} t_j^dag)^-1 ; W_j = W_{j+1} t_j^dag g_j ;
    G_00 = g_0 (with Sigma_L in block 0), G_{M-1,0} = W_0.
    """
    Es = np.asarray(Es, dtype=complex)
    M, n = h.shape[0], h.shape[-1]
    eye = np.eye(n)
    zE = Es[:, None, None] * eye
    g = np.linalg.inv(zE - h[M - 1] - SigR)
    W = g.copy()
    for j in range(M - 2, -1, -1):
        t = tnn[j]
        A = zE - h[j] - t @ g @ np.conj(t).T
        if j == 0:
            A = A - SigL
        g = np.linalg.inv(A)
        W = W @ np.conj(t).T @ g
    return g, W


def rt_from_blocks(G00, GN0, SigL, SigR, sources):
    """R[nE, n_src, n], T[nE, n_src, n] from the reference's formulas (wfm/__init__.py:91, :108-130)."""
    sources = np.atleast_2d(np.asarray(sources, dtype=float))
    vL = -2 * np.imag(np.diagonal(SigL, axis1=-2, axis2=-1))   # [nE, n]
    vR = -2 * np.imag(np.diagonal(SigR, axis1=-2, axis2=-1))
    nE, n = vL.shape
    R = np.zeros((nE, len(sources), n))
    T = np.zeros((nE, len(sources), n))
    with np.errstate(divide="ignore", invalid="ignore"):
        for s, A in enumerate(sources):
            Av = A[None, :] * vL                                # [nE, n]
            i_flux = np.sqrt(np.einsum("a,ea->e", A, A[None, :] * np.real(vL)))
            r = (1j * np.einsum("eab,eb->ea", G00, Av) - A[None, :]) * np.sqrt(np.real(vL)) / i_flux[:, None]
            t = 1j * np.einsum("eab,eb->ea", GN0, Av) * np.sqrt(np.real(vR)) / i_flux[:, None]
            R[:, s] = np.real(r * np.conj(r))
            T[:, s] = np.real(t * np.conj(t))
    return R, T


def transmission_closed(h, tnn, Es, sources):
    """Full closed-form-lead path: (R, T) each [nE, n_src, n]."""
    SigL, SigR = closed_selfenergies(h, tnn, Es)
    G00, GN0 = rgf_blocks(h, tnn, Es, SigL, SigR)
    return rt_from_blocks(G00, GN0, SigL, SigR, sources)


def caroli_trace(GN0, SigL, SigR):
    """Tr[Gamma_R G Gamma_L G^dag] with matrix Gamma = i(Sigma - Sigma^dag)  (SURVEY §8a row a10)."""
    GamL = 1j * (SigL - np.conj(np.swapaxes(SigL, -1, -2)))
    GamR = 1j * (SigR - np.conj(np.swapaxes(SigR, -1, -2)))
    X = GamR @ GN0 @ GamL @ np.conj(np.swapaxes(GN0, -1, -2))
    return np.real(np.trace(X, axis1=-2, axis2=-1))


def green_column0(h, tnn, Es, SigL, SigR):
    """First block column G[:, 0] as [nE, M, n, n] (what `psi_jsigma` needs, wfm/__init__.py:98)."""
    Es = np.asarray(Es, dtype=complex)
    M, n = h.shape[0], h.shape[-1]
    eye = np.eye(n)
    zE = Es[:, None, None] * eye
    gR = [None] * M
    gR[M - 1] = np.linalg.inv(zE - h[M - 1] - SigR)
    for j in range(M - 2, 0, -1):
        t = tnn[j]
        gR[j] = np.linalg.inv(zE - h[j] - t @ gR[j + 1] @ np.conj(t).T)
    t0 = tnn[0]
    out = np.zeros((len(Es), M, n, n), dtype=complex)
    out[:, 0] = np.linalg.inv(zE - h[0] - SigL - t0 @ gR[1] @ np.conj(t0).T)
    for j in range(1, M):
        out[:, j] = gR[j] @ np.conj(tnn[j - 1]).T @ out[:, j - 1]
    return out


def green_full(h, tnn, Es, SigL, SigR):
    """All M x M blocks of inv(E - H') as [nE, M, M, n, n] by left/right sweeps."""
    Es = np.asarray(Es, dtype=complex)
    M, n = h.shape[0], h.shape[-1]
    eye = np.eye(n)
    zE = Es[:, None, None] * eye
    dag = lambda a: np.conj(np.swapaxes(a, -1, -2))
    hh = [zE - h[j] for j in range(M)]
    hh[0] = hh[0] - SigL
    hh[M - 1] = hh[M - 1] - SigR
    gL = [None] * M
    gR = [None] * M
    gL[0] = np.linalg.inv(hh[0])
    for j in range(1, M):
        gL[j] = np.linalg.inv(hh[j] - dag(tnn[j - 1]) @ gL[j - 1] @ tnn[j - 1])
    gR[M - 1] = np.linalg.inv(hh[M - 1])
    for j in range(M - 2, -1, -1):
        gR[j] = np.linalg.inv(hh[j] - tnn[j] @ gR[j + 1] @ dag(tnn[j]))
    G = np.zeros((len(Es), M, M, n, n), dtype=complex)
    for j in range(M):
        A = hh[j]
        if j > 0:
            A = A - dag(tnn[j - 1]) @ gL[j - 1] @ tnn[j - 1]
        if j < M - 1:
            A = A - tnn[j] @ gR[j + 1] @ dag(tnn[j])
        G[:, j, j] = np.linalg.inv(A)
    for c in range(M):
        for r in range(c + 1, M):       # below the diagonal
            G[:, r, c] = gR[r] @ dag(tnn[r - 1]) @ G[:, r - 1, c]
        for r in range(c - 1, -1, -1):  # above
            G[:, r, c] = gL[r] @ tnn[r] @ G[:, r + 1, c]
    return G


def sancho_rubio(h00, h01, zs, side, tol=1e-14, max_iter=200):
    """Surface GF of a semi-infinite lead by decimation (Lopez Sancho et al. 1985).

    Fixed point of the reference's own map (wfm/__init__.py:547-551):
      side=+1 (right lead):  g = (z - h00 - h01 g h01^dag)^-1
      side=-1 (left lead):   g = (z - h00 - h01^dag g h01)^-1
    Returns (g [nz,n,n], n_iter).
    """
    zs = np.asarray(zs, dtype=complex)
    n = h00.shape[0]
    eye = np.eye(n)
    dag = lambda a: np.conj(np.swapaxes(a, -1, -2))
    if side == 1:
        alpha = np.broadcast_to(h01, (len(zs), n, n)).copy()
        beta = np.broadcast_to(dag(h01), (len(zs), n, n)).copy()
    else:
        alpha = np.broadcast_to(dag(h01), (len(zs), n, n)).copy()
        beta = np.broadcast_to(h01, (len(zs), n, n)).copy()
    eps_s = np.broadcast_to(h00, (len(zs), n, n)).astype(complex).copy()
    eps = eps_s.copy()
    zI = zs[:, None, None] * eye
    it = 0
    for it in range(1, max_iter + 1):
        ginv = np.linalg.inv(zI - eps)
        agb = alpha @ ginv @ beta
        bga = beta @ ginv @ alpha
        eps_s = eps_s + agb
        eps = eps + agb + bga
        alpha = alpha @ ginv @ alpha
        beta = beta @ ginv @ beta
        if max(np.abs(alpha).max(), np.abs(beta).max()) < tol:
            break
    return np.linalg.inv(zI - eps_s), it


def rgf_blocks_telescoped(h, tnn, Es, SigL, SigR, kmax=8, gmax=64.0, return_stats=False, cond_max=None, max_extend=3,
                          pole_max=None):
    """The same two blocks by the *telescoped* sweep the sm_100a kernel `rgf_sweep_n8t` runs (scalar hopping
    t_j = tau_j I only).  Inside a chunk of sites [a..b] the left-connected g_j are never formed:

        D_{b+1} = I,  D_{b+2} := g_{b+1},   D_j = z_j D_{j+1} - |t_j|^2 D_{j+2},   z_j = E - h_j (- Sigma)
        g_j = D_{j+1} D_j^-1   =>   prod_{j=b..a} g_j = D_a^-1   (telescopes)
        X = D_a^-1,   g_a = D_{a+1} X,   W_a = (prod conj t_j) W_{b+1} X

    so a chunk of k sites costs k-1 products + 1 inverse + 2 products instead of k inverses + k products.
    A chunk is closed after `kmax` sites or as soon as max|D_j| / prod|t| exceeds `gmax` (error growth
    bound); with kmax=1 this is exactly `rgf_blocks`.  Checker for the kernel's numerics; not product code.

    `cond_max`: the kernel's guard against closing a chunk on a nearly singular D_a.  In long, nearly closed chains
    (localised regime) the right-connected g_a has poles next to the real axis, |g_a| ~ 1e6, and the next inverse undoes
    it with a cancellation that costs cond(D_a) * eps of relative accuracy (the plain recursion meets this at every
    site).  If max|D_a| max|D_a^-1| exceeds cond_max the chunk is NOT closed there: it is extended by one more site (up
    to `max_extend` times), which needs no inverse of the offending matrix at all.
    `pole_max` is the criterion the kernel uses (rgf_n8t.cuh, GMAXX): extend when max|t_{a-1} g_a| > pole_max, g_a = D_{a+1}
    D_a^-1 — the quantity whose size IS the amplification of the next chunk's rounding errors.
    """
    Es = np.asarray(Es, dtype=complex)
    M, n = h.shape[0], h.shape[-1]
    nE = len(Es)
    eye = np.broadcast_to(np.eye(n, dtype=complex), (nE, n, n))
    zE = Es[:, None, None] * np.eye(n)
    tau_all = np.array([tnn[j][0, 0] for j in range(M - 1)])
    g = None                     # g_{b+1}
    W = None                     # W_{b+1} (None = identity)
    j = M - 1
    chunks = []
    while j >= 0:
        b = j
        Dp2 = g                  # D_{b+2}
        Dp1 = eye.copy()         # D_{b+1}
        tprod = 1.0 + 0j
        scale = 1.0
        k = 0
        extended = 0
        while True:
            z = zE - h[j]
            if j == M - 1:
                z = z - SigR
            if j == 0:
                z = z - SigL
            D = z @ Dp1
            if j < M - 1:
                kap = abs(tau_all[j]) ** 2
                tprod = tprod * np.conj(tau_all[j])
                scale *= abs(tau_all[j])
                if Dp2 is not None:
                    D = D - kap * Dp2
            k += 1
            growth = np.abs(D).max() / max(scale, 1e-300)
            a = j
            if j == 0 or k >= kmax or growth > gmax:
                if pole_max is not None and j > 0 and extended < max_extend:
                    Xtry = np.linalg.inv(D)
                    gtry = (Dp1 @ Xtry) if k > 1 else Xtry
                    if np.max(np.abs(gtry)) * abs(tau_all[j - 1]) <= pole_max:
                        break
                    extended += 1
                    Dp2, Dp1 = Dp1, D
                    j -= 1
                    continue
                if cond_max is None or j == 0 or extended >= max_extend:
                    break
                Xtry = np.linalg.inv(D)
                dmax, xmax = np.abs(D).max(axis=(1, 2)) / max(scale, 1e-300), np.abs(Xtry).max(axis=(1, 2)) * max(scale, 1e-300)
                # a resonance pole: |D^-1| >> |D| = O(1); evanescent growth makes both large and is left to `gmax`
                if not np.any((dmax * xmax > cond_max) & (xmax > 4096.0 * dmax)):
                    break
                extended += 1          # nearly singular: take one more site instead of inverting here
            Dp2, Dp1 = Dp1, D
            j -= 1
        X = np.linalg.inv(D)
        g = Dp1 @ X if k > 1 else X
        W = tprod * (X if W is None else W @ X)
        chunks.append(k)
        j = a - 1
    if return_stats:
        return g, W, chunks
    return g, W
