"""Next-nearest-neighbour blocks (`tnnn`, transport/wfm/__init__.py:209-213) on the block-TRIdiagonal device kernels.

The reference assembles a block-pentadiagonal dense matrix when `tnnn` is nonzero (`Hmat` :168-215) and inverts it.
Here two neighbouring sites are paired into one super-site of size 2n, which turns (h, tnn, tnnn) with M sites into a
block-tridiagonal chain of ceil(M/2) super-sites; the sweep / Green's-function kernels run on that chain with table
self-energies placed on the sub-blocks of the original first and last site, and the outputs are un-paired again.
"""
import numpy as np

from . import green, leads
from ._lib import as_c128
from .sweep import transmission_sweep


def pair_into_superblocks(h, tnn, tnnn):
    """Block-pentadiagonal (h, tnn, tnnn) with M sites -> block-tridiagonal (H, T) with ceil(M/2)
    super-sites of size 2n.  An odd chain is padded with one decoupled dummy site far off in energy."""
    M, n = h.shape[0], h.shape[-1]
    if M % 2:
        h = np.concatenate([h, 1e6 * np.eye(n, dtype=complex)[None]])
        tnn = np.concatenate([tnn, np.zeros((1, n, n), dtype=complex)])
        tnnn = np.concatenate([tnnn, np.zeros((1, n, n), dtype=complex)])
        M += 1
    K = M // 2
    H = np.zeros((K, 2 * n, 2 * n), dtype=complex)
    T = np.zeros((K - 1, 2 * n, 2 * n), dtype=complex)
    for k in range(K):
        a, b = 2 * k, 2 * k + 1
        H[k, :n, :n] = h[a]
        H[k, n:, n:] = h[b]
        H[k, :n, n:] = tnn[a]
        H[k, n:, :n] = np.conj(tnn[a]).T
        if k + 1 < K:
            T[k, :n, :n] = tnnn[a]          # site a   -> a+2
            T[k, n:, :n] = tnn[b]           # site b   -> b+1 = a+2
            T[k, n:, n:] = tnnn[b]          # site b   -> b+2
    return H, T, M


def _paired_problem(h, tnn, tnnn, SL, SR):
    """Super-block chain + self-energy tables [nE, 2n, 2n] from the per-site tables SL, SR [nE, n, n]."""
    h, tnn, tnnn = as_c128(h), as_c128(tnn), as_c128(tnnn)
    M0, n = h.shape[0], h.shape[-1]
    H, T, M = pair_into_superblocks(h, tnn, tnnn)
    o = ((M0 - 1) % 2) * n                     # offset of the last real site inside its super-block
    if (M0 - 1) // 2 != H.shape[0] - 1:
        raise AssertionError("internal: right lead must sit in the last super-block")
    sigL = np.zeros((len(SL), 2 * n, 2 * n), dtype=complex)
    sigR = np.zeros_like(sigL)
    sigL[:, :n, :n] = SL
    sigR[:, o:o + n, o:o + n] = SR
    return H, T, sigL, sigR, o, M0, n


def sweep_with_tnnn(h, tnn, tnnn, Es, sources, converger="g_closed", tables=None):
    """R, T [nE, n_src, n] for a chain with next-nearest-neighbour blocks; the leads continue (h[0], tnn[0]) and
    (h[-1], tnn[-1]) as in wfm.SelfEnergies (:283-284; the reference ignores tnnn there too)."""
    Es = as_c128(np.atleast_1d(Es))
    SL, SR = tables if tables is not None else leads.self_energies_batched(as_c128(h), as_c128(tnn), Es, converger)
    H, T, sigL, sigR, o, M0, n = _paired_problem(h, tnn, tnnn, SL, SR)
    src = np.zeros((len(sources), 2 * n))
    src[:, :n] = np.asarray(sources, dtype=float)
    R2, T2 = transmission_sweep(H, T, Es, "table", sources=src, sigL=sigL, sigR=sigR)
    return R2[0, :, :, :n], T2[0, :, :, o:o + n]


def green_column0_with_tnnn(h, tnn, tnnn, Es, SL, SR):
    """First block column G[:, 0] [nE, M, n, n] of the pentadiagonal problem."""
    Es = as_c128(np.atleast_1d(Es))
    H, T, sigL, sigR, o, M0, n = _paired_problem(h, tnn, tnnn, SL, SR)
    col = green.green_column0(H, T, Es, sigL, sigR)                   # [nE, K, 2n, 2n]
    nE, K = col.shape[0], col.shape[1]
    out = col[:, :, :, :n].reshape(nE, K, 2, n, n).reshape(nE, 2 * K, n, n)   # rows (site a, site b) of column-site 0
    return np.ascontiguousarray(out[:, :M0])


def green_full_with_tnnn(h, tnn, tnnn, Es, SL, SR):
    """inv(E - H') as [nE, M, M, n, n] of the pentadiagonal problem."""
    Es = as_c128(np.atleast_1d(Es))
    H, T, sigL, sigR, o, M0, n = _paired_problem(h, tnn, tnnn, SL, SR)
    G = green.green_full(H, T, Es, sigL, sigR)                        # [nE, K, K, 2n, 2n]
    nE, K = G.shape[0], G.shape[1]
    G = G.reshape(nE, K, K, 2, n, 2, n).transpose(0, 1, 3, 2, 5, 4, 6).reshape(nE, 2 * K, 2 * K, n, n)
    return np.ascontiguousarray(G[:, :M0, :M0])
