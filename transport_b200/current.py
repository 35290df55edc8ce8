"""Landauer current from the transmission function, on the device (csrc/green.cu:landauer_kernel).

SURVEY.md §8f row 4.  The reference's matrix-Gamma Landauer code lives in `fcdmft/__init__.py`
(`landauer:224-292`: jE = Tr[G_adv Lambda_R G_ret Lambda_L] (n_L - n_R)/(2 pi), integrated with
`np.trapz`, `:426`); here the trace is the Caroli trace / channel sum that the sweep already produced.
"""
import numpy as np

from . import _lib
from ._lib import as_f64, check, ptr


def landauer_current(T_total, energies, mu_L, mu_R, kBT=0.0, want_density=False):
    """I[b] = trapz(T(E) (n_L - n_R)/(2 pi), E) for every bias point b.

    T_total [nE] total transmission (e.g. `want_total` / `want_caroli` of `transmission_sweep`),
    energies [nE] real and increasing, mu_L/mu_R scalars or [n_bias]."""
    T = as_f64(np.real(T_total))
    E = as_f64(np.real(energies))
    if T.shape != E.shape or T.ndim != 1:
        raise ValueError("T_total and energies must be 1-D of equal length")
    muL, muR = np.broadcast_arrays(np.atleast_1d(np.asarray(mu_L, dtype=float)), np.atleast_1d(np.asarray(mu_R, dtype=float)))
    muL, muR = as_f64(muL), as_f64(muR)
    nb = len(muL)
    I = np.empty(nb, dtype=np.float64)
    jE = np.empty((nb, len(E)), dtype=np.float64) if want_density else None
    check(_lib.load().tx_landauer_current(ptr(T), ptr(E), len(E), ptr(muL), ptr(muR), nb, float(kBT), ptr(jE), ptr(I)))
    return (I, jE) if want_density else I
