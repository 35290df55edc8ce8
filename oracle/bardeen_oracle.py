"""
TEST INFRASTRUCTURE — not product code.  Nothing under transport_b200/ imports this module.

CPU restatement (numpy + LAPACK `eigh`) of the Bardeen transfer-Hamiltonian path of cpbunker/transport
(BASELINE config 5).  Reference: transport/bardeen/__init__.py

    Hsysmat            :690-750
    kernel_well        :17-257     (restated for any n_loc_dof; the CUDA path covers n_loc_dof = 1)
    kernel_well_prime  :400-532
    matrix_element     :835-857
    current            :537-575    nFD :1014-1020
    Ts_bardeen         :577-632

Parity is PINNED: tests/golden/bardeen_*.npz are outputs of the unmodified reference
(oracle/make_golden.py) and tests/test_oracle.py checks this file against them.

Caveat stated by the reference's own construction: the Oppenheimer matrix elements are overlaps of
exponentially small eigenvector tails (|M|^2 ~ 1e-15 .. 1e-17 at the driver parameters), so two correct
eigensolvers agree on |M|^2 only to ~1e-7 relative (LAPACK's absolute 1e-16 error on a 1e-8 tail);
`eigensolver_spread` below measures that spread between LAPACK drivers.
"""
import numpy as np


def Hsysmat(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC, bound=True):
    """bardeen/__init__.py:690-750, vectorised by region."""
    NC = len(HC)
    if NC % 2 != 1:
        raise ValueError
    n = np.shape(tinfty)[0]
    VinfL, VinfR = (Vinfty, Vinfty) if bound else (VL, VR)
    S = 2 * Ninfty + NL + NR + NC
    H = np.zeros((S, S, n, n), dtype=complex)
    # region boundaries in 0-based site index: [0,Ninfty) inf | [Ninfty,Ninfty+NL) L | C | R | inf
    a, b = Ninfty, Ninfty + NL
    c, d = b + NC, b + NC + NR
    for j in range(S):
        if j < a:
            H[j, j] += VinfL; H[j, j + 1] += -tinfty; H[j + 1, j] += -tinfty
        elif j < b:
            H[j, j] += VL; H[j, j + 1] += -tL; H[j + 1, j] += -tL
        elif j < c:
            pass
        elif j < d:
            H[j, j] += VR; H[j, j - 1] += -tR; H[j - 1, j] += -tR
        else:
            H[j, j] += VinfR; H[j, j - 1] += -tinfty; H[j - 1, j] += -tinfty
    H[b:c, b:c] = HC
    return H


def mat_4d_to_2d(mat):
    S, _, n, _ = mat.shape
    return mat.transpose(0, 2, 1, 3).reshape(S * n, S * n)


def _scalar_hopping(t):
    """:78-88 / :444-454: hopping matrices must be spin diagonal and spin independent."""
    if np.any(t - np.diagflat(np.diagonal(t))):
        raise ValueError("not spin diagonal")
    diag = np.diagonal(t)
    if np.any(diag - diag[0]):
        raise ValueError("not spin independent")
    return t[0, 0]


def _window_average(E_m, E_n, interval, element, sign_fix):
    """The averaging loop shared by :127-151 and :502-519: mean of element(n) over |E_m - E_n| < interval,
    nearest n if interval is nan, 0.0 if nothing qualifies."""
    Mns = []
    for n in range(len(E_n)):
        if abs(E_m - E_n[n]) < interval:
            el = element(n)
            if sign_fix and np.real(el) < 0:
                el = -el
            Mns.append(el)
    if np.isnan(interval) and Mns == []:
        n_nearest = int(np.argmin(abs(E_m - E_n)))
        el = element(n_nearest)
        if sign_fix and np.real(el) < 0:
            el = -el
        Mns.append(el)
    return sum(Mns) / len(Mns) if Mns else 0.0


def kernel_well_prime(tinfty, tL, tR, Vinfty, VL, VLprime, VR, VRprime, Ninfty, NL, NR, HC, HCprime, E_cutoff,
                      interval=1e-9, verbose=0):
    """bardeen/__init__.py:400-532."""
    if np.shape(HC) != np.shape(HCprime):
        raise ValueError
    n = np.shape(HC)[-1]
    assert np.any(HC - HCprime)
    tLa, tRa = _scalar_hopping(tL), _scalar_hopping(tR)
    S = 2 * Ninfty + NL + NR + len(HC)

    def well_states(H4, ta):                                 # :457-478 / :481-503
        Es, psis = [], []
        for a in range(n):
            Ea, pa = np.linalg.eigh(H4[:, :, a, a])
            keep = Ea + 2 * ta < E_cutoff[a, a]
            Es.append(Ea[keep]); psis.append(pa.T[keep])
        nb = max(len(e) for e in Es)
        E_arr = np.empty((n, nb), dtype=complex)
        psi_arr = np.empty((n, nb, S), dtype=complex)
        for a in range(n):
            pad = nb - len(Es[a])
            E_arr[a] = np.append(Es[a], np.full((pad,), Es[a][-1]))
            psi_arr[a] = np.append(psis[a], np.full((pad, S), psis[a][-1]), axis=0)
        return E_arr, psi_arr

    HL = Hsysmat(tinfty, tL, tR, Vinfty, VL, VRprime, Ninfty, NL, NR, HCprime)
    Emas, psimas = well_states(HL, tLa)
    HR = Hsysmat(tinfty, tL, tR, Vinfty, VLprime, VR, Ninfty, NL, NR, HCprime)
    Enbs, psinbs = well_states(HR, tRa)
    Hsys = Hsysmat(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC)
    Hdiff = Hsys - HL                                        # [S,S,n,n]
    nbl = Emas.shape[1]
    Mbmas = np.empty((n, nbl, n), dtype=float)
    for alpha in range(n):
        for m in range(nbl):
            for beta in range(n):
                op = Hdiff[:, :, beta, alpha]                # matrix_element :835-857
                y = op @ psimas[alpha, m]
                val = _window_average(Emas[alpha, m], Enbs[beta], interval,
                                      lambda k: np.dot(np.conj(psinbs[beta, k]), y), False)
                Mbmas[beta, m, alpha] = np.real(val)         # complex -> float store (:498, :521)
    return Emas, Mbmas * Mbmas


def kernel_well(tinfty, tL, tR, Vinfty, VL, VLprime, VR, VRprime, Ninfty, NL, NR, HC, HCobs, defines_Sz,
                E_cutoff, interval=1e-12, expval_tol=1e-12, verbose=0):
    """bardeen/__init__.py:17-257."""
    if np.shape(HC) != np.shape(HCobs):
        raise ValueError
    n = np.shape(HC)[-1]
    if n != len(defines_Sz):
        raise ValueError
    S = 2 * Ninfty + NL + NR + len(HC)
    tLa, tRa = _scalar_hopping(tL), _scalar_hopping(tR)

    def mstates(H4):                                         # get_mstates :755-761
        E, psi = np.linalg.eigh(mat_4d_to_2d(H4))
        return E.astype(complex), psi.T

    Hsys = Hsysmat(tinfty, tL, tR, Vinfty, VL, VR, Ninfty, NL, NR, HC)
    HL = Hsysmat(tinfty, tL, tR, Vinfty, VL, VRprime, Ninfty, NL, NR, HC)
    Ems, psims = mstates(HL)
    Emus, psimus = mstates(Hsysmat(tinfty, tL, tR, Vinfty, VL, VRprime, Ninfty, NL, NR, HCobs))
    Ens, psins = mstates(Hsysmat(tinfty, tL, tR, Vinfty, VLprime, VR, Ninfty, NL, NR, HC))
    Enus, psinus = mstates(Hsysmat(tinfty, tL, tR, Vinfty, VLprime, VR, Ninfty, NL, NR, HCobs))

    Hdiff = mat_4d_to_2d(Hsys - HL)
    Mms = np.empty_like(Ems)
    for m in range(len(Ems)):                                # :127-151
        y = Hdiff @ psims[m]
        Mms[m] = _window_average(Ems[m], Ens, interval, lambda k: np.dot(np.conj(psins[k]), y), True)
    change_basis = np.conj(psims) @ psimus.T                 # :160-163, [m, mu]
    Mnumus = np.conj(change_basis.T) @ (Mms[:, None] * change_basis)       # :166
    Mnumus = np.real(np.conj(Mnumus) * Mnumus).astype(float)

    cut = np.max(E_cutoff)                                   # :174-182
    mu_cut = (len(Emus[Emus + 2 * tLa < cut]) // n) * n
    nu_cut = (len(Enus[Enus + 2 * tRa < cut]) // n) * n
    Emus, psimus = Emus[:mu_cut], psimus[:mu_cut]
    Enus, psinus = Enus[:nu_cut], psinus[:nu_cut]
    Mnumus = Mnumus[:nu_cut, :mu_cut]

    expvals_exact = np.linalg.eigvalsh(defines_Sz)           # :192
    sz_mu = np.real(np.einsum("mja,ab,mjb->m", np.conj(psimus.reshape(mu_cut, S, n)), defines_Sz,
                              psimus.reshape(mu_cut, S, n)))
    sz_nu = np.real(np.einsum("mja,ab,mjb->m", np.conj(psinus.reshape(nu_cut, S, n)), defines_Sz,
                              psinus.reshape(nu_cut, S, n)))

    def classify(sz):                                        # :203-212: the LAST matching eigenvalue wins
        idx = 2
        for k in range(n):
            if abs(sz - expvals_exact[k]) < expval_tol:
                idx = k
        if idx not in range(n):
            raise Exception("<Sz> = " + str(sz) + " not an eigenval of Sz")
        return idx

    E_mualphas = np.zeros((n, mu_cut), dtype=complex)
    M4 = np.empty((n, nu_cut, n, mu_cut), dtype=float)       # [beta, nu, alpha, mu], uninitialised like :190
    M4[:] = np.nan
    mucounter = np.zeros(n, dtype=int)
    nucounter = np.zeros(n, dtype=int)
    for mu in range(mu_cut):
        a = classify(sz_mu[mu])
        E_mualphas[a, mucounter[a]] = Emus[mu]
        mucounter[a] += 1
        nucounter = np.zeros(n, dtype=int)
        for nu in range(nu_cut):
            b = classify(sz_nu[nu])
            nucounter[b] += 1
            M4[b, nucounter[b] - 1, a, mucounter[a] - 1] = Mnumus[nu, mu]
    nmu, nnu = int(np.min(mucounter)), int(np.min(nucounter))
    E_out = E_mualphas[:, :nmu]
    Mbmas_eff = np.empty((n, nmu, n), dtype=float)
    for beta in range(n):                                    # :236-250 (note the alpha/beta swap at :238)
        for alpha in range(n):
            for mu in range(nmu):
                Mbmas_eff[beta, mu, alpha] = np.sum(M4[alpha, :nnu, beta, mu])
    return E_out, Mbmas_eff


def nFD(epsilon, mu, kBT):
    """:1014-1020."""
    return 1 / (np.exp((epsilon - mu) / kBT) + 1)


def current(Emas, Mbmas, Vb, tbulk, muR, kBT, verbose=0):
    """bardeen/__init__.py:537-575."""
    n, nb = np.shape(Emas)
    Emas = np.real(Emas)
    Iab = np.empty((n, n, len(Vb)), dtype=float)
    for alpha in range(n):
        for beta in range(n):
            for i in range(len(Vb)):
                muL = muR + Vb[i]
                fL, fR = nFD(Emas[alpha], muL, kBT), nFD(Emas[alpha], muR, kBT)
                stat = fL * (1 - fR) - fR * (1 - fL)
                Iab[alpha, beta, i] = 2 * np.pi * np.dot(stat, Mbmas[alpha, :, beta])
    return Iab


def Ts_bardeen(Emas, Mbmas, tL, tR, VL, VR, NL, NR, verbose=0):
    """bardeen/__init__.py:577-632."""
    if Mbmas.dtype != float:
        raise TypeError
    n, nb = np.shape(Emas)
    if len(np.shape(Mbmas)) != 3 or np.shape(Mbmas)[0] != n:
        raise ValueError
    conv = []
    for c in [tL, VL, tR, VR]:
        if np.any(c - np.diagflat(np.diagonal(c))):
            raise ValueError
        conv.append(np.diagonal(c))
    tLa, VLa, tRa, VRa = conv
    kmas = np.arccos((Emas - VLa[:, None]) / (-2 * tLa[:, None]))
    if abs(np.max(np.imag(kmas))) > 1e-10:
        raise ValueError
    kmas = np.real(kmas)
    Tbmas = np.empty_like(Mbmas)
    for alpha in range(n):
        for beta in range(n):
            dos = 1 / np.sqrt(1 - np.power((Emas[alpha] - VRa[beta]) / (2 * tRa[beta]), 2))
            if abs(np.max(np.imag(dos))) > 1e-10:
                raise ValueError
            Tbmas[alpha, :, beta] = NL / (kmas[alpha] * tLa[alpha]) * NR / tRa[beta] * np.real(dos) * Mbmas[alpha, :, beta]
    return Tbmas


def eigensolver_spread(H2d):
    """max relative difference of squared eigenvector components >1e-40 between LAPACK drivers (needs scipy)."""
    from scipy.linalg import eigh
    _, v1 = eigh(H2d, driver="evd")
    _, v2 = eigh(H2d, driver="evr")
    a, b = v1 * v1, v2 * v2
    mask = a > 1e-40
    return float(np.max(np.abs(a - b)[mask] / a[mask]))
