"""Drop-in mirror of the two `transport.bardeen` functions on the T(E) path
(reference: transport/bardeen/__init__.py): `Hsysmat` (:690-750) and `Ts_wfm_well` (:634-685).

`Ts_wfm_well` in the reference still calls `wfm.kernel` with the old 6-argument signature and raises
TypeError against the current tree.  Here it is fixed forward: the block
arrays are built once and ONE batched sweep evaluates every (alpha, m) pair.  Next-nearest-neighbour
blocks (the only place a nonzero `tnnn` can arise, :663-664) are handled by pairing sites into
super-blocks of size 2n, which turns the pentadiagonal chain into a block-tridiagonal one.
"""
import numpy as np

from . import _lib, leads
from ._lib import as_c128, check, ptr
from .sweep import transmission_sweep


def Hsysmat(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC, bound=True) -> np.ndarray:
    """4-D tight-binding Hamiltonian [S,S,n,n] of the inf|L|C|R|inf geometry (bardeen/__init__.py:690-750)."""
    for arg in [tinfty, tL, tR, Vinfty, VL, VR]:
        if type(arg) != np.ndarray: raise TypeError
    for N in [Ninfty, NL, NR]:
        if not isinstance(N, int): raise TypeError
        if N <= 0: raise ValueError
    if np.shape(HC[0, 0]) != np.shape(tinfty): raise ValueError
    if len(HC) % 2 != 1: raise ValueError                   # NC must be odd
    if bound:
        VinfL, VinfR = Vinfty, Vinfty
    else:
        print("\n\nWARNING: NOT BOUND\n\n")
        VinfL, VinfR = VL, VR
    n = np.shape(tinfty)[0]
    NC = len(HC)
    S = 2 * Ninfty + NL + NR + NC
    mats = [as_c128(x) for x in (tinfty, tL, tR, VinfL, VinfR, VL, VR)]
    HCc = as_c128(HC)
    out = np.empty((S, S, n, n), dtype=np.complex128)
    check(_lib.load().tx_hsysmat(*[ptr(m) for m in mats], Ninfty, NL, NR, ptr(HCc), NC, n, ptr(out)))
    return out


def blocks_from_HC(tL, tR, VL, VR, HC):
    """(hblocks[N+2], tnn[N+1], tnnn[N]) from the central-region Hamiltonian HC[N,N,n,n] and the lead
    parameters, exactly as `Ts_wfm_well` builds them (bardeen/__init__.py:650-669)."""
    N = np.shape(HC)[0]
    n = np.shape(HC)[-1]
    eye = np.eye(n)
    h = np.zeros((N + 2, n, n), dtype=complex)
    tnn = np.zeros((N + 1, n, n), dtype=complex)
    tnnn = np.zeros((N, n, n), dtype=complex)
    h[0] = VL * eye
    tnn[0] = -tL * eye
    for i in range(N):
        h[1 + i] = HC[i, i]
        if i + 1 < N:
            tnn[1 + i] = HC[i, i + 1]
        if i + 2 < N:
            tnnn[1 + i] = HC[i, i + 2]
        for j in range(i + 3, N):
            assert not np.any(HC[i, j])                     # :665-666
    h[-1] = VR * eye
    tnn[-1] = -tR * eye
    return h, tnn, tnnn


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


def sweep_with_tnnn(h, tnn, tnnn, Es, sources):
    """R, T [nE, n_src, n] for a chain with next-nearest-neighbour blocks (closed-form leads on the
    original first/last sites), via super-blocks and table self-energies."""
    h, tnn, tnnn = as_c128(h), as_c128(tnn), as_c128(tnnn)
    M0, n = h.shape[0], h.shape[-1]
    Es = as_c128(np.atleast_1d(Es))
    SL, SR = leads.self_energies_batched(h, tnn, Es, "g_closed")
    H, T, M = pair_into_superblocks(h, tnn, tnnn)
    last_in_block = (M0 - 1) % 2                # position of the last real site inside its super-block
    K = H.shape[0]
    kR = (M0 - 1) // 2
    if kR != K - 1:
        raise AssertionError("internal: right lead must sit in the last super-block")
    sigL = np.zeros((len(Es), 2 * n, 2 * n), dtype=complex)
    sigR = np.zeros_like(sigL)
    sigL[:, :n, :n] = SL
    o = last_in_block * n
    sigR[:, o:o + n, o:o + n] = SR
    src = np.zeros((len(sources), 2 * n))
    src[:, :n] = np.asarray(sources, dtype=float)
    R2, T2 = transmission_sweep(H, T, Es, "table", sources=src, sigL=sigL, sigR=sigR)
    return R2[0, :, :, :n], T2[0, :, :, o:o + n]


def Ts_wfm_well(tL, tR, VL, VR, HC, Emas, verbose=0) -> np.ndarray:
    """Transmission probabilities at the bound-state energies Emas[alpha, m], final spin resolved:
    Tbmas[beta, m, alpha] (bardeen/__init__.py:634-685), one batched device sweep."""
    if np.any(tL - tR): raise NotImplementedError           # :643
    if np.shape(Emas)[0] != np.shape(HC)[-1]: raise ValueError
    n = np.shape(HC)[-1]
    n_bound = np.shape(Emas)[-1]
    tLs = tL[0, 0] if np.ndim(tL) == 2 else tL              # the reference passes tL[alpha,alpha]; tL = t*I
    tRs = tR[0, 0] if np.ndim(tR) == 2 else tR
    VLs = VL[0, 0] if np.ndim(VL) == 2 else VL
    VRs = VR[0, 0] if np.ndim(VR) == 2 else VR
    h, tnn, tnnn = blocks_from_HC(tLs, tRs, VLs, VRs, HC)
    Es = np.asarray(Emas, dtype=complex).reshape(-1)        # index alpha*n_bound + m
    sources = np.eye(n)
    if np.any(tnnn):
        _, T = sweep_with_tnnn(h, tnn, tnnn, Es, sources)
    else:
        _, T = transmission_sweep(h, tnn, Es, "g_closed", sources=sources)
        T = T[0]
    Tbmas = np.empty((n, n_bound, n), dtype=float)
    for alpha in range(n):
        for m in range(n_bound):
            Tbmas[:, m, alpha] = T[alpha * n_bound + m, alpha]
    return Tbmas
