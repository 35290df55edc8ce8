"""
TEST INFRASTRUCTURE — not product code.  Nothing under transport_b200/ may import this.

numpy restatement of the reference's matrix-Gamma Landauer current, `/root/reference/fcdmft/__init__.py`:
`landauer` :224-292 (Meir-Wingreen Eq. 7 for a noninteracting region), the closed-form branch of `surface_gf`
:298-342 it takes for diagonal leads, `decompose_gf` :525-574 (retarded / advanced parts) and the trapezoid rule the
package integrates with (`np.trapz`, :426).  SURVEY.md §8f row 4.

Pinning: tests/golden/landauer_fcdmft.npz holds the output of the UNMODIFIED `fcdmft.landauer` (imported with stubs for
its pyscf-dependent subpackage, oracle/make_golden_r2.py:landauer_fcdmft); tests/test_oracle.py checks this file
against it, tests/test_gpu_parity.py checks the device's Caroli trace and `tx_landauer_current` against the same file.

Array convention of the reference: "spinful arrays" [spin, orb, orb, energy].
"""
import numpy as np


def surface_gf_diag(energies, H, V):
    """Closed form for diagonal, orbital-independent leads (fcdmft/__init__.py:312-342)."""
    energies = np.real(energies)
    gf = np.zeros((*np.shape(H), len(energies)), dtype=complex)
    for s in range(np.shape(H)[0]):
        for o in range(np.shape(H)[1]):
            pref = (energies - H[s, o, o]) / (2 * V[s, o, o])
            gf[s, o, o, :] = pref - np.lib.scimath.sqrt(pref * pref * (1 - 4 * V[s, o, o] * V[s, o, o] / np.power(energies - H[s, o, o], 2)))
    return gf


def _dot_op(a, op, backwards=False):
    """dot_spinful_arrays with a frequency-independent operator (:492-498)."""
    out = np.zeros_like(a, dtype=complex)
    for s in range(a.shape[0]):
        for iw in range(a.shape[-1]):
            out[s, :, :, iw] = (op[s] @ a[s, :, :, iw]) if backwards else (a[s, :, :, iw] @ op[s])
    return out


def fermi(energies, mu, kBT):
    """:240-247 (a step function, 1 at E <= mu, for kBT == 0)."""
    if kBT == 0.0:
        n = np.zeros_like(energies, dtype=int)
        n[energies <= mu] = 1
        return n
    return 1 / (np.exp((energies - mu) / kBT) + 1)


def transmission(energies, MBGF, LL, RL):
    """Tr[G_adv Lambda_R G_ret Lambda_L](E) of :250-270: Lambda = -2 Im(coup gf coup), G_ret = MBGF, G_adv = MBGF^dag."""
    LL_diag, LL_hop, LL_coup, _ = LL
    RL_diag, RL_hop, RL_coup, _ = RL
    n_imp = np.shape(LL_coup)[1]
    G_ret = MBGF[:, :n_imp, :n_imp]
    G_adv = np.conj(np.swapaxes(G_ret, 1, 2))
    lam = []
    for diag, hop, coup in ((LL_diag, LL_hop, LL_coup), (RL_diag, RL_hop, RL_coup)):
        hyb = _dot_op(_dot_op(surface_gf_diag(energies, diag, hop), coup), coup, backwards=True)
        lam.append(-2 * np.imag(hyb))
    T = np.zeros(len(energies), dtype=complex)
    for iw in range(len(energies)):
        T[iw] = np.trace(G_adv[0, :, :, iw] @ lam[1][0, :, :, iw] @ G_ret[0, :, :, iw] @ lam[0][0, :, :, iw])
    return T


def landauer(energies, kBT, MBGF, LL, RL):
    """jE(E) = Tr[...] (n_L - n_R) / (2 pi)   (:270)."""
    nL, nR = fermi(energies, LL[3], kBT), fermi(energies, RL[3], kBT)
    return transmission(energies, MBGF, LL, RL) * (nL - nR) / (2 * np.pi)


def current(energies, jE):
    """Trapezoid rule, np.trapz as the package integrates (:426)."""
    return np.trapezoid(np.real(jE), np.real(energies))
