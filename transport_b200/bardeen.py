"""Drop-in mirror of `transport.bardeen` (reference: transport/bardeen/__init__.py): `Hsysmat` (:690-750),
`Ts_wfm_well` (:634-685) and the transfer-Hamiltonian path of BASELINE config 5 — `kernel_well` (:17-257),
`kernel_well_prime` (:400-532), `current` (:537-575), `Ts_bardeen` (:577-632) — plus the batched entry
`wells_batched` the scalar ones are a batch-of-one of.  Every number comes from libtransport_b200.so
(csrc/bardeen.cu); numpy is used for argument checks and for slicing the region parameters into the
tridiagonal arrays the C ABI takes.

`Ts_wfm_well` in the reference still calls `wfm.kernel` with the old 6-argument signature and raises
TypeError against the current tree.  Here it is fixed forward: the block
arrays are built once and ONE batched sweep evaluates every (alpha, m) pair.  Next-nearest-neighbour
blocks (the only place a nonzero `tnnn` can arise, :663-664) are handled by pairing sites into
super-blocks of size 2n, which turns the pentadiagonal chain into a block-tridiagonal one.
"""
import numpy as np

from . import _lib, leads
from ._lib import as_c128, as_f64, check, ptr
from .superblocks import pair_into_superblocks, sweep_with_tnnn  # noqa: F401  (re-exported)
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


def Ts_wfm_well(tL, tR, VL, VR, HC, Emas, verbose=0) -> np.ndarray:
    """Transmission probabilities at the bound-state energies Emas[alpha, m], final spin resolved:
    Tbmas[beta, m, alpha] (bardeen/__init__.py:634-685), one batched device sweep."""
    if np.any(tL - tR): raise NotImplementedError           # :643
    if np.shape(Emas)[0] != np.shape(HC)[-1]: raise ValueError
    n = np.shape(HC)[-1]
    n_bound = np.shape(Emas)[-1]
    # the reference multiplies the parameter MATRICES elementwise with the identity (:651-653, :667-668): the
    # per-channel diagonal survives (drivers pass VL = diag(0, Delta), rbmenez_sup_inel.py:135), off-diagonals drop
    h, tnn, tnnn = blocks_from_HC(tL, tR, VL, VR, HC)
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


# ------------------------------------------------------------------------------------------------
# Bardeen transfer-Hamiltonian path (config 5)
# ------------------------------------------------------------------------------------------------
def _scalar_hopping(t):
    """bardeen/__init__.py:78-88, :444-454: hopping matrices must be spin diagonal and spin independent."""
    if np.any(t - np.diagflat(np.diagonal(t))): raise ValueError("not spin diagonal")
    diag = np.diagonal(t)
    if np.any(diag - diag[0]): raise ValueError("not spin independent")
    return t[0, 0]


def hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC):
    """The block-tridiagonal content of Hsysmat(...) (bound=True) without the S x S zeros:
    (diag[S,n,n], upper[S-1,n,n] = H[j,j+1], lower[S-1,n,n] = H[j+1,j]).  HC must be a chain
    (block tridiagonal in the site index); anything else is outside what the device path covers."""
    HC = np.asarray(HC)
    NC, n = HC.shape[0], HC.shape[-1]
    if NC % 2 != 1: raise ValueError
    for i in range(NC):
        for j in range(NC):
            if abs(i - j) > 1 and np.any(HC[i, j]):
                raise NotImplementedError("HC couples sites beyond nearest neighbours: not a tridiagonal chain")
    S = 2 * Ninfty + NL + NR + NC
    a, b = Ninfty, Ninfty + NL
    c, d = b + NC, b + NC + NR
    diag = np.zeros((S, n, n), dtype=complex)
    up = np.zeros((S - 1, n, n), dtype=complex)
    lo = np.zeros((S - 1, n, n), dtype=complex)
    diag[:a] = Vinfty; diag[a:b] = VL; diag[c:d] = VR; diag[d:] = Vinfty
    idx = np.arange(NC)
    diag[b:c] = HC[idx, idx]
    up[:a] = -tinfty; up[a:b] = -tL; up[c - 1:d - 1] = -tR; up[d - 1:] = -tinfty
    lo[:] = up
    up[b:c - 1] = HC[idx[:-1], idx[:-1] + 1]
    lo[b:c - 1] = HC[idx[:-1] + 1, idx[:-1]]
    return diag, up, lo


def _channel_tridiagonals(blocks):
    """Per-channel real symmetric tridiagonals d[n,S], e[n,S-1] of a block-tridiagonal Hamiltonian whose
    blocks are diagonal in the channel index (the reference takes H[:, :, a, a], :459 and :483)."""
    diag, up, lo = blocks
    n = diag.shape[-1]
    ch = np.arange(n)
    d, e, el = diag[:, ch, ch].T, up[:, ch, ch].T, lo[:, ch, ch].T
    if np.any(np.imag(d)) or np.any(np.imag(e)) or np.any(e != el):
        raise NotImplementedError("well Hamiltonians must be real symmetric (complex hoppings are not covered)")
    return as_f64(np.real(d)), as_f64(np.real(e))


def _is_channel_diagonal(blocks):
    n = blocks[0].shape[-1]
    off = ~np.eye(n, dtype=bool)
    return not any(np.any(b[:, off]) for b in blocks)


def wells_batched(dL, eL, dR, eR, cutL, cutR_states, cutR_count, hd0, hdU, hdL, interval, sign_fix,
                  want_states=False):
    """One device pass over P geometries x n channels (tx_bardeen_wells).  Shapes: dL, dR [P,n,S];
    eL, eR [P,n,S-1]; cut* [P,n]; hd0 [P,S,n,n]; hdU, hdL [P,S-1,n,n].
    Returns dict(EL, nL, ER, nR_states, nR_below, Mavg[, psiL, psiR]); Mavg[p, beta, m, alpha]."""
    lib = _lib.load()
    dL, eL, dR, eR = as_f64(dL), as_f64(eL), as_f64(dR), as_f64(eR)
    P, n, S = dL.shape
    cutL, cutRs, cutRc = as_f64(cutL), as_f64(cutR_states), as_f64(cutR_count)
    hd0, hdU, hdL = as_f64(hd0), as_f64(hdU), as_f64(hdL)
    assert eL.shape == (P, n, S - 1) and dR.shape == dL.shape and eR.shape == eL.shape
    assert hd0.shape == (P, S, n, n) and hdU.shape == (P, S - 1, n, n) and hdL.shape == hdU.shape
    nL = np.zeros((P, n), dtype=np.int32); nRs = np.zeros_like(nL); nRb = np.zeros_like(nL)
    args = [ptr(dL), ptr(eL), ptr(dR), ptr(eR), ptr(cutL), ptr(cutRs), ptr(cutRc), ptr(hd0), ptr(hdU), ptr(hdL), P, n, S]
    check(lib.tx_bardeen_wells(*args, 0, float(interval), int(sign_fix), None, ptr(nL), None, ptr(nRs), ptr(nRb),
                               None, None, None))
    ms = int(max(nL.max(), nRs.max(), 1))
    EL = np.zeros((P, n, ms)); ER = np.zeros((P, n, ms)); Mavg = np.zeros((P, n, ms, n))
    psiL = np.zeros((P, n, ms, S)) if want_states else None
    psiR = np.zeros((P, n, ms, S)) if want_states else None
    check(lib.tx_bardeen_wells(*args, ms, float(interval), int(sign_fix), ptr(EL), ptr(nL), ptr(ER), ptr(nRs), ptr(nRb),
                               ptr(Mavg), ptr(psiL), ptr(psiR)))
    out = dict(EL=EL, nL=nL, ER=ER, nR_states=nRs, nR_below=nRb, Mavg=Mavg)
    if want_states:
        out.update(psiL=psiL, psiR=psiR)
    return out


_region_capacity = {}     # (P, n, S) -> max_states of the last call: sweeps of similar geometries take one pass


def region_sweep(tinfty, tL, tR, Vinfty, VL, VLprime, VR, VRprime, Ninfty, NL, NR, HC, HCprime=None, cutL=None,
                 cutR_states=None, cutR_count=None, interval=1e-9, sign_fix=True, row_cut=False, Vb=None, muR=0.0,
                 kBT=0.0):
    """P geometries of the inf|L|C|R|inf chain given by their REGION PARAMETERS, assembled on the device
    (tx_bardeen_region_sweep): bound states of both wells, window-averaged matrix elements and, with `Vb`, the current.

    tinfty, tL, tR, Vinfty: [n] per-channel scalars (or [n,n] diagonal matrices as the reference passes them);
    VL, VLprime, VR, VRprime: [n] / [n,n] (shared) or [P,n];  HC, HCprime: [NC,NC,n,n] (shared) or [P,NC,NC,n,n], real
    nearest-neighbour chains;  cut*: [P,n].  Returns dict(EL, nL, ER, nR_states, nR_below, Mavg[, I], h2d_bytes, d2h_bytes)."""
    lib = _lib.load()

    def chan(x):
        x = np.asarray(x)
        if x.ndim == 2 and x.shape[0] == x.shape[1]:
            d = np.diagonal(x)
            if np.count_nonzero(x) == np.count_nonzero(d):      # diagonal matrix -> its diagonal
                x = d
        if np.any(np.imag(x)):
            raise NotImplementedError("complex region parameters are not covered by the device path")
        return as_f64(np.real(x))

    tinf_c, tL_c, tR_c, Vinf_c = chan(tinfty), chan(tL), chan(tR), chan(Vinfty)
    n = len(tinf_c)

    def pots(x):
        x = np.asarray(x)
        if x.ndim == 2 and x.shape == (n, n) and n > 1:
            return chan(x)[None]
        x = as_f64(np.real(x))
        return x.reshape(-1, n) if x.ndim > 1 or n == 1 else x[None]

    Vs = [pots(v) for v in (VL, VLprime, VR, VRprime)]
    n_pot = max(v.shape[0] for v in Vs)
    Vs = [as_f64(np.broadcast_to(v, (n_pot, n))) for v in Vs]

    def chain(H):
        H = np.asarray(H)
        if np.any(np.imag(H)):
            raise NotImplementedError("complex HC is not covered by the device path")
        H = np.real(H)
        if H.ndim == 4:
            H = H[None]
        NC = H.shape[1]
        if NC % 2 != 1: raise ValueError
        idx = np.arange(NC)
        if np.any(H[:, np.abs(idx[:, None] - idx[None, :]) > 1]):
            raise NotImplementedError("HC couples sites beyond nearest neighbours: not a tridiagonal chain")
        return (as_f64(H[:, idx, idx]), as_f64(H[:, idx[:-1], idx[:-1] + 1]), as_f64(H[:, idx[:-1] + 1, idx[:-1]]))

    c0, cU, cL = chain(HC)
    if HCprime is None:
        p0 = pU = pL = None
    else:
        p0, pU, pL = chain(HCprime)
        if p0.shape != c0.shape: raise ValueError
    n_hc, NC = c0.shape[0], c0.shape[1]
    cutL, cutRs, cutRc = as_f64(cutL), as_f64(cutR_states), as_f64(cutR_count)
    P = cutL.shape[0]
    if n_pot not in (1, P) or n_hc not in (1, P):
        raise ValueError("per-geometry parameters must have P rows")
    S = 2 * Ninfty + NL + NR + NC
    nL = np.zeros((P, n), dtype=np.int32); nRs = np.zeros_like(nL); nRb = np.zeros_like(nL)
    Vb_a = None if Vb is None else as_f64(Vb)
    nVb = 0 if Vb is None else len(Vb_a)
    import ctypes
    needed = ctypes.c_int(0)
    ms = _region_capacity.get((P, n, S), 0)
    while True:
        EL = np.zeros((P, n, max(ms, 1))); ER = np.zeros_like(EL); Mavg = np.zeros((P, n, max(ms, 1), n))
        Iab = np.zeros((P, n, n, nVb)) if nVb else None
        check(lib.tx_bardeen_region_sweep(
            ptr(tinf_c), ptr(tL_c), ptr(tR_c), ptr(Vinf_c), ptr(Vs[0]), ptr(Vs[1]), ptr(Vs[2]), ptr(Vs[3]), n_pot,
            ptr(c0), ptr(cU), ptr(cL), ptr(p0), ptr(pU), ptr(pL), n_hc, int(Ninfty), int(NL), int(NR), NC, n, P,
            ptr(cutL), ptr(cutRs), ptr(cutRc), float(interval), int(sign_fix), int(row_cut), ms, ctypes.byref(needed),
            ptr(EL), ptr(nL), ptr(ER), ptr(nRs), ptr(nRb), ptr(Mavg), ptr(Vb_a), nVb, float(muR), float(kBT), ptr(Iab)))
        if needed.value <= ms or needed.value == 0:
            break
        ms = needed.value + max(2, needed.value // 16)          # a little head room for the next geometries of a sweep
    _region_capacity[(P, n, S)] = ms
    used = max(needed.value, 1)
    out = dict(EL=EL[:, :, :used], nL=nL, ER=ER[:, :, :used], nR_states=nRs, nR_below=nRb, Mavg=Mavg[:, :, :used])
    if nVb:
        out["I"] = Iab
    out["h2d_bytes"] = int(sum(a.nbytes for a in (tinf_c, tL_c, tR_c, Vinf_c, *Vs, c0, cU, cL, cutL, cutRs, cutRc))
                           + (0 if p0 is None else p0.nbytes + pU.nbytes + pL.nbytes) + (Vb_a.nbytes if nVb else 0))
    out["d2h_bytes"] = int(EL.nbytes + ER.nbytes + Mavg.nbytes + 3 * nL.nbytes + (Iab.nbytes if nVb else 0))
    return out


def step_sweep(tinfty, tL, tR, Vinfty, VL, VRs, Ninfty, NL, NR, HC, E_cutoff, interval=1e-9, Vb=None, muR=0.0, kBT=0.0,
               row_cut=False):
    """The sweep of rbstep.py:43-104 — `kernel_well` (+ `current`) looped over the right-well step height VR — as ONE
    device pass over all step heights `VRs` [P] (n_loc_dof = 1): for every VR the wells are (VL, VR + Vinfty) and
    (VL + ... ) exactly as rbstep.py builds VLprime = Vinfty + VL ... (VLprime = VR + Vinfty on the left of the right
    well, VRprime = Vinfty on the right of the left well)."""
    VRs = as_f64(np.atleast_1d(VRs))
    P = len(VRs)
    n = np.shape(tinfty)[0]
    if n != 1:
        raise NotImplementedError("step_sweep covers n_loc_dof = 1 (rbstep.py); use region_sweep for channel-resolved wells")
    tLa, tRa = _scalar_hopping(np.asarray(tL)), _scalar_hopping(np.asarray(tR))
    cut = float(np.max(np.real(E_cutoff)))
    cutL = np.full((P, n), cut - 2 * np.real(tLa))
    cutR = np.full((P, n), cut - 2 * np.real(tRa))
    cutR_states = np.full((P, n), np.inf) if np.isnan(interval) else np.maximum(cutR, cutL) + abs(interval)
    Vinf0 = float(np.real(np.asarray(Vinfty)[0, 0]))
    VR_col = VRs[:, None]
    return region_sweep(tinfty, tL, tR, Vinfty, VL, VR_col + Vinf0, VR_col, np.full((P, n), Vinf0), Ninfty, NL, NR, HC, None,
                        cutL, cutR_states, cutR, interval, sign_fix=True, row_cut=row_cut, Vb=Vb, muR=muR, kBT=kBT)


def step_sweep_arrays(tinfty, tL, tR, Vinfty, VL, VRs, Ninfty, NL, NR, HC):
    """Host-assembled arrays of the same sweep (the inputs of `wells_batched` / tx_bardeen_wells_dev): dL, eL, dR, eR,
    h0, hU, hL stacked over the step heights.  Kept for device-resident measurements and as the cross-check of the
    device assembly."""
    arrs = {k: [] for k in ("dL", "eL", "dR", "eR", "h0", "hU", "hL")}
    for v in np.atleast_1d(VRs):
        VR = v * np.eye(np.shape(tinfty)[0])
        bL = hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, Vinfty, Ninfty, NL, NR, HC)
        bR = hsys_site_blocks(tinfty, tL, tR, Vinfty, VR + Vinfty, VR, Ninfty, NL, NR, HC)
        bS = hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC)
        a, b = _channel_tridiagonals(bL); arrs["dL"].append(a); arrs["eL"].append(b)
        a, b = _channel_tridiagonals(bR); arrs["dR"].append(a); arrs["eR"].append(b)
        hd = [np.real(x - y) for x, y in zip(bS, bL)]
        arrs["h0"].append(hd[0]); arrs["hU"].append(hd[1]); arrs["hL"].append(hd[2])
    return {k: np.ascontiguousarray(np.array(v), dtype=np.float64) for k, v in arrs.items()}


def _real_blocks(blocks):
    if any(np.any(np.imag(b)) for b in blocks):
        raise NotImplementedError("complex Hsys - HL is not covered by the device path")
    return [as_f64(np.real(b)) for b in blocks]


def kernel_well_prime(tinfty, tL, tR, Vinfty, VL, VLprime, VR, VRprime, Ninfty, NL, NR, HC, HCprime, E_cutoff,
                      interval=1e-9, verbose=0) -> tuple:
    """Oppenheimer matrix elements |M_{beta m alpha}|^2 between Sz-resolved well states, averaged over the
    final states inside the energy window (bardeen/__init__.py:400-532).  Returns (Emas[n, nb] complex,
    Mbmas[n(beta), nb, n(alpha)] float)."""
    if np.shape(HC) != np.shape(HCprime): raise ValueError
    n = np.shape(HC)[-1]
    assert np.any(HC - HCprime)                             # :432
    tLa, tRa = _scalar_hopping(tL), _scalar_hopping(tR)
    bL = hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, VRprime, Ninfty, NL, NR, HCprime)
    bR = hsys_site_blocks(tinfty, tL, tR, Vinfty, VLprime, VR, Ninfty, NL, NR, HCprime)
    bS = hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC)
    assert _is_channel_diagonal(bL)                         # is_alpha_conserving, :457
    assert _is_channel_diagonal(bR)                         # :481
    dL, eL = _channel_tridiagonals(bL)
    dR, eR = _channel_tridiagonals(bR)
    hd = _real_blocks([s - l for s, l in zip(bS, bL)])      # Hdiff = Hsys - HL, :497
    Ecut = np.real(np.diagonal(np.asarray(E_cutoff)))
    cutL = Ecut - 2 * np.real(tLa)
    cutR = Ecut - 2 * np.real(tRa)
    r = wells_batched(dL[None], eL[None], dR[None], eR[None], cutL[None], cutR[None], cutR[None],
                      hd[0][None], hd[1][None], hd[2][None], interval, sign_fix=False)
    nL = r["nL"][0]
    nb = int(nL.max())
    if np.any(nL < 1) or np.any(r["nR_states"][0] < 1):
        raise IndexError("a channel has no bound state below E_cutoff")     # the reference's Ems[-1] fails too
    # ragged channels are padded with their last bound state (:468-478)
    pick = np.minimum(np.arange(nb)[None, :], (nL - 1)[:, None])            # [alpha, m] -> stored m
    Emas = np.take_along_axis(r["EL"][0], pick, axis=1).astype(complex)
    Mavg = r["Mavg"][0]                                                     # [beta, m, alpha]
    Mbmas = np.empty((n, nb, n), dtype=float)
    for alpha in range(n):
        Mbmas[:, :, alpha] = Mavg[:, pick[alpha], alpha]
    return Emas, Mbmas * Mbmas


def kernel_well(tinfty, tL, tR, Vinfty, VL, VLprime, VR, VRprime, Ninfty, NL, NR, HC, HCobs, defines_Sz,
                E_cutoff, interval=1e-12, expval_tol=1e-12, verbose=0) -> tuple:
    """Effective Oppenheimer matrix elements of bardeen/__init__.py:17-257 for n_loc_dof = 1 (rbbarrier.py,
    rbstep.py — the geometries of BASELINE config 5).  With one local degree of freedom the observable basis is
    the eigenbasis (HCobs = HC), the basis change :160-166 is the identity and the spin classification :192-232
    is trivial, so Mbmas_eff[0, mu, 0] = |<M_mu>|^2 for mu < nu_cut and (rounding noise in the reference, exactly
    0 here) for mu >= nu_cut, the reference's row cut :181 acting on the diagonal."""
    if np.shape(HC) != np.shape(HCobs): raise ValueError
    n = np.shape(HC)[-1]
    if n != len(defines_Sz): raise ValueError
    tLa, tRa = _scalar_hopping(tL), _scalar_hopping(tR)
    if n != 1 or np.any(np.asarray(HC) - np.asarray(HCobs)):
        return _kernel_well_general(tinfty, tL, tR, Vinfty, VL, VLprime, VR, VRprime, Ninfty, NL, NR, np.asarray(HC),
                                    np.asarray(HCobs), np.asarray(defines_Sz), E_cutoff, interval, expval_tol, tLa, tRa)
    bL = hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, VRprime, Ninfty, NL, NR, HC)
    bR = hsys_site_blocks(tinfty, tL, tR, Vinfty, VLprime, VR, Ninfty, NL, NR, HC)
    bS = hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC)
    dL, eL = _channel_tridiagonals(bL)
    dR, eR = _channel_tridiagonals(bR)
    hd = _real_blocks([s - l for s, l in zip(bS, bL)])
    cut = float(np.max(np.real(E_cutoff)))                  # :174
    cutL = np.array([cut - 2 * np.real(tLa)])
    cutR = np.array([cut - 2 * np.real(tRa)])
    # the reference averages over ALL right-well states inside the window (:127-151), also above the cutoff
    cutR_states = np.array([np.inf]) if np.isnan(interval) else np.maximum(cutR, cutL) + abs(interval)
    r = wells_batched(dL[None], eL[None], dR[None], eR[None], cutL[None], cutR_states[None], cutR[None],
                      hd[0][None], hd[1][None], hd[2][None], interval, sign_fix=True)
    mu_cut, nu_cut = int(r["nL"][0, 0]), int(r["nR_below"][0, 0])
    Emas = r["EL"][0, :, :mu_cut].astype(complex)
    M = r["Mavg"][0, :, :mu_cut, :].copy()
    M[:, nu_cut:, :] = 0.0
    return Emas, M * M


def _common_spin_basis(blocks, n):
    """A unitary U that diagonalises every n x n block of `blocks` (they must be normal and commute pairwise), or
    NotImplementedError.  For the reference's drivers (rbmenez_spinless_el.py:97-108: every block is a 1 + b sigma_x) this
    is the sigma_x eigenbasis; in it the well Hamiltonians decouple into n independent real symmetric tridiagonals."""
    U = np.eye(n, dtype=complex)
    for B in blocks:
        Br = U.conj().T @ B @ U
        if np.max(np.abs(Br - np.diag(np.diagonal(Br)))) <= 1e-13 * max(1.0, np.max(np.abs(Br))):
            continue
        if np.max(np.abs(Br - Br.conj().T)) > 1e-13:
            raise NotImplementedError("non-Hermitian spin block: the well Hamiltonians do not decouple into chains")
        # refine U inside the degenerate subspaces of what has been diagonalised so far: only valid if nothing was yet
        # resolved (first non-scalar block) -- otherwise the blocks do not share an eigenbasis
        if not np.allclose(U, np.eye(n)):
            raise NotImplementedError("the spin blocks of the wells do not share one eigenbasis")
        _, U = np.linalg.eigh(B)
    for B in blocks:
        Br = U.conj().T @ B @ U
        off = Br - np.diag(np.diagonal(Br))
        if np.max(np.abs(off)) > 1e-12 * max(1.0, np.max(np.abs(Br))) or np.max(np.abs(np.imag(np.diagonal(Br)))) > 1e-12:
            raise NotImplementedError("the spin blocks of the wells do not share one real eigenbasis")
    return U


def _all_states(d, e):
    """All eigenpairs of the n_mat tridiagonals d [n_mat,S], e [n_mat,S-1] on the device: (E [n_mat,S], psi [n_mat,S,S])."""
    n_mat, S = d.shape
    top = float(np.max(np.abs(d)) + 2 * np.max(np.abs(e)) + 1.0)       # above every Gershgorin disc
    cut = np.full(n_mat, top)
    E = np.zeros((n_mat, S))
    psi = np.zeros((n_mat, S, S))
    ns = np.zeros(n_mat, dtype=np.int32)
    nb = np.zeros(n_mat, dtype=np.int32)
    check(_lib.load().tx_tridiag_eigh_below(ptr(as_f64(d)), ptr(as_f64(e)), S, n_mat, ptr(cut), ptr(cut), S, ptr(E), ptr(psi),
                                            ptr(ns), ptr(nb)))
    if np.any(ns != S):
        raise AssertionError("internal: incomplete spectrum")
    return E, psi


def _real_gemm(A, B, w=None, transA=False, transB=False):
    """opA(A) diag(w) opB(B) on the device (tx_real_gemm)."""
    A, B = as_f64(A), as_f64(B)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    assert (B.shape[1] if transB else B.shape[0]) == K
    C = np.empty((M, N))
    wv = None if w is None else as_f64(w)
    check(_lib.load().tx_real_gemm(ptr(A), ptr(B), ptr(wv), ptr(C), M, N, K, int(transA), int(transB)))
    return C


def _kernel_well_general(tinfty, tL, tR, Vinfty, VL, VLprime, VR, VRprime, Ninfty, NL, NR, HC, HCobs, defines_Sz, E_cutoff,
                         interval, expval_tol, tLa, tRa):
    """kernel_well (bardeen/__init__.py:17-257) for n_loc_dof > 1 with a separate observable basis (HCobs != HC), as
    rbmenez_spinless_el.py:45-127 calls it.

    The reference diagonalises four dense (S n)^2 Hamiltonians.  Here the spin blocks of the wells are required to share
    one eigenbasis U (every driver: a 1 + b sigma_x per site), in which H_L and H_R decouple into n real symmetric
    tridiagonals per well, and the observable Hamiltonians (HCobs) must conserve the channel (the reference asserts it,
    :104, :112), which makes them n tridiagonals in the original basis.  All eigenpairs come from the device's
    tridiagonal solver; the matrix elements <psi_n|H_sys - H_L|psi_m> (:127-151), the overlaps <psi_m|psi_mu> (:160-163)
    and the change of basis (:166) are device products; only index bookkeeping (cut-offs :174-182, classification by
    <S_z> :192-232, the final sums :236-250) runs on the host, restated from the reference including its index quirks.
    """
    n = HC.shape[-1]
    NC = HC.shape[0]
    S = 2 * Ninfty + NL + NR + NC
    if np.isnan(interval):
        raise NotImplementedError("adaptive (NaN) interval with n_loc_dof > 1")
    spin_blocks = [np.asarray(x, dtype=complex) for x in (Vinfty, VL, VLprime, VR, VRprime)]
    for i in range(NC):
        for j in range(NC):
            if abs(i - j) <= 1:
                spin_blocks.append(np.asarray(HC[i, j], dtype=complex))
    for t in (tinfty, tL, tR):
        _scalar_hopping(np.asarray(t))
    U = _common_spin_basis(spin_blocks, n)
    rot = lambda B: np.diag(np.real(np.diagonal(U.conj().T @ np.asarray(B, dtype=complex) @ U)))
    HCr = np.zeros_like(HC)
    for i in range(NC):
        for j in range(NC):
            if abs(i - j) <= 1:
                HCr[i, j] = rot(HC[i, j])
    Vi, VLr, VLpr, VRr, VRpr = rot(Vinfty), rot(VL), rot(VLprime), rot(VR), rot(VRprime)
    bL = hsys_site_blocks(tinfty, tL, tR, Vi, VLr, VRpr, Ninfty, NL, NR, HCr)
    bR = hsys_site_blocks(tinfty, tL, tR, Vi, VLpr, VRr, Ninfty, NL, NR, HCr)
    bS = hsys_site_blocks(tinfty, tL, tR, Vi, VLr, VRr, Ninfty, NL, NR, HCr)
    dL, eL = _channel_tridiagonals(bL)
    dR, eR = _channel_tridiagonals(bR)
    hd = _real_blocks([s_ - l_ for s_, l_ in zip(bS, bL)])
    # observable basis: channel conserving in the ORIGINAL basis (:104, :112)
    bLo = hsys_site_blocks(tinfty, tL, tR, Vinfty, VL, VRprime, Ninfty, NL, NR, HCobs)
    bRo = hsys_site_blocks(tinfty, tL, tR, Vinfty, VLprime, VR, Ninfty, NL, NR, HCobs)
    assert _is_channel_diagonal(bLo)
    assert _is_channel_diagonal(bRo)
    dLo, eLo = _channel_tridiagonals(bLo)
    dRo, eRo = _channel_tridiagonals(bRo)

    # ---- M_m for EVERY eigenstate of H_L (get_mstates has no cutoff, :755-761), averaged over the right-well states
    #      inside the window with the sign fix of :141
    top = float(max(np.max(np.abs(dL)), np.max(np.abs(dR))) + 2 * max(np.max(np.abs(eL)), np.max(np.abs(eR))) + 1.0)
    cut_all = np.full((1, n), top)
    r = wells_batched(dL[None], eL[None], dR[None], eR[None], cut_all, cut_all, cut_all, hd[0][None], hd[1][None], hd[2][None],
                      interval, sign_fix=True, want_states=True)
    if np.any(r["nL"] != S) or np.any(r["nR_states"] != S):
        raise AssertionError("internal: incomplete spectrum")
    EL, ER, psiL = r["EL"][0], r["ER"][0], r["psiL"][0]                # [n,S], [n,S], [n,S,S]
    # the reference averages over the right-well states of ALL channels inside the window: the device returns the mean
    # per final channel beta, the counts recombine them (elements between different rotated channels vanish but count)
    cnt = np.zeros((n, S, n))                                          # [beta, m, alpha]
    for a in range(n):
        for b in range(n):
            cnt[b, :, a] = np.count_nonzero(np.abs(EL[a][:, None] - ER[b][None, :]) < interval, axis=1)
    tot = cnt.sum(axis=0)                                              # [m, alpha]
    Mm = np.where(tot > 0, (cnt * r["Mavg"][0]).sum(axis=0) / np.where(tot > 0, tot, 1), 0.0)   # [m, alpha]

    # ---- observable states below the cutoff, merged and sorted by energy like the reference's dense eigh
    ELo, phiL = _all_states(dLo, eLo)                                  # [n,S], [n,S,S] in the original channels
    ERo, _ = _all_states(dRo, eRo)
    cutE = np.max(E_cutoff)

    def merged(Eo):
        order = np.argsort(Eo.reshape(-1), kind="stable")
        return [(int(i // S), int(i % S)) for i in order]
    mus, nus = merged(ELo), merged(ERo)
    mu_cut = (sum(1 for (a, l) in mus if ELo[a, l] + 2 * np.real(tLa) < cutE) // n) * n
    nu_cut = (sum(1 for (a, l) in nus if ERo[a, l] + 2 * np.real(tRa) < cutE) // n) * n
    mus, nus = mus[:mu_cut], nus[:nu_cut]
    rows = mus[:nu_cut]                 # the ROW index of Mnumus also runs over the left-well observable states (:166)
    if nu_cut > mu_cut:
        raise NotImplementedError("more right- than left-well states below the cutoff")

    # ---- Mnumus = C^dag diag(M) C with C[(c,k),(a,l)] = <chi_ck|phi_al> conj(U[a,c])   (:160-166)
    need = sorted(set(mus))
    per_chan = {a: [l for (a2, l) in need if a2 == a] for a in range(n)}
    O = {}
    for c in range(n):
        for a in range(n):
            if per_chan[a]:
                O[c, a] = _real_gemm(psiL[c], phiL[a][per_chan[a]], transB=True)        # [S, K_a]
    pos = {(a, l): i for a in range(n) for i, l in enumerate(per_chan[a])}
    Q = {}
    for a1 in range(n):
        for a2 in range(n):
            if not per_chan[a1] or not per_chan[a2]:
                continue
            acc = np.zeros((len(per_chan[a1]), len(per_chan[a2])), dtype=complex)
            for c in range(n):
                acc += U[a1, c] * np.conj(U[a2, c]) * _real_gemm(O[c, a1], O[c, a2], w=Mm[:, c], transA=True)
            Q[a1, a2] = acc
    Mnumus = np.empty((len(rows), mu_cut))
    for i, (a1, l1) in enumerate(rows):
        for jx, (a2, l2) in enumerate(mus):
            q = Q[a1, a2][pos[a1, l1], pos[a2, l2]]
            Mnumus[i, jx] = np.real(np.conj(q) * q)

    # ---- classification by <S_z> (:192-232) and the final sums (:236-250), with the reference's index conventions
    expvals = np.linalg.eigvalsh(np.asarray(defines_Sz, dtype=complex))

    def classify(a):
        sz = np.real(np.asarray(defines_Sz)[a, a])
        idx = None
        for k in range(n):
            if abs(sz - expvals[k]) < expval_tol:
                idx = k
        if idx is None:
            raise Exception("<Sz> = " + str(sz) + " not an eigenval of Sz")
        return idx
    mu_cls = [classify(a) for (a, l) in mus]
    nu_cls = [classify(a) for (a, l) in nus]
    mucount = np.bincount(mu_cls, minlength=n)
    nucount = np.bincount(nu_cls, minlength=n)
    nmu, nnu = int(mucount.min()), int(nucount.min())
    E_out = np.zeros((n, nmu), dtype=complex)
    M4 = np.zeros((n, max(nucount.max(), 1), n, max(mucount.max(), 1)))    # [Sz(nu), rank of nu, Sz(mu), rank of mu]
    seen_mu = np.zeros(n, dtype=int)
    for jx, (a, l) in enumerate(mus):
        ca = mu_cls[jx]
        if seen_mu[ca] < nmu:
            E_out[ca, seen_mu[ca]] = ELo[a, l]
        seen_nu = np.zeros(n, dtype=int)
        for i in range(len(nus)):
            cb = nu_cls[i]
            M4[cb, seen_nu[cb], ca, seen_mu[ca]] = Mnumus[i, jx]
            seen_nu[cb] += 1
        seen_mu[ca] += 1
    Mbmas_eff = np.empty((n, nmu, n))
    for beta in range(n):
        for alpha in range(n):
            Mbmas_eff[beta, :, alpha] = M4[alpha, :nnu, beta, :nmu].sum(axis=0)      # the alpha/beta swap of :238
    return E_out, Mbmas_eff


def current(Emas, Mbmas, Vb, tbulk, muR, kBT, verbose=0) -> np.ndarray:
    """Current as a function of bias voltage, Iab[alpha, beta, Vb] (bardeen/__init__.py:537-575)."""
    for arr in [Emas, Mbmas, Vb]:
        if not isinstance(arr, np.ndarray): raise TypeError
    n, nb = np.shape(Emas)
    E = as_f64(np.real(Emas))
    M = as_f64(Mbmas)
    V = as_f64(Vb)
    Iab = np.empty((n, n, len(V)), dtype=float)
    check(_lib.load().tx_bardeen_current(ptr(E), ptr(M), 1, n, nb, ptr(V), len(V), float(muR), float(kBT), ptr(Iab)))
    return Iab


def Ts_bardeen(Emas, Mbmas, tL, tR, VL, VR, NL, NR, verbose=0) -> np.ndarray:
    """Transmission coefficients from the averaged matrix elements (bardeen/__init__.py:577-632)."""
    if Mbmas.dtype != float: raise TypeError
    n, nb = np.shape(Emas)
    if len(np.shape(Mbmas)) != 3: raise ValueError
    if np.shape(Mbmas)[0] != n: raise ValueError
    conv = []
    for c in [tL, VL, tR, VR]:
        if np.any(c - np.diagflat(np.diagonal(c))): raise ValueError
        conv.append(as_f64(np.real(np.diagonal(c))))
    E = as_f64(np.real(Emas))
    M = as_f64(Mbmas)
    T = np.empty_like(M)
    flags = np.zeros(2 + n * n, dtype=np.int32)
    check(_lib.load().tx_bardeen_ts(ptr(E), ptr(M), n, nb, ptr(conv[0]), ptr(conv[1]), ptr(conv[2]), ptr(conv[3]),
                                    int(NL), int(NR), ptr(T), ptr(flags)))
    if flags[0] or flags[1] or np.any(flags[2:] == 0): raise ValueError     # :611, :624
    return T
