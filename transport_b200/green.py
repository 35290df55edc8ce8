"""Green's-function blocks and dense assembly on the device (csrc/green.cu)."""
import numpy as np

from . import _lib
from ._lib import as_c128, check, ptr


def _args(h, tnn, Es, sigL, sigR):
    h, tnn = as_c128(h), as_c128(tnn)
    Es = as_c128(np.atleast_1d(Es))
    sigL, sigR = as_c128(sigL), as_c128(sigR)
    M, n = h.shape[0], h.shape[-1]
    if tnn.shape != (M - 1, n, n) or sigL.shape != (len(Es), n, n) or sigR.shape != sigL.shape:
        raise ValueError("shape mismatch")
    return h, tnn, Es, sigL, sigR, M, n


def green_column0(h, tnn, Es, sigL, sigR):
    """G[:,0] for every energy -> [nE, M, n, n]."""
    h, tnn, Es, sigL, sigR, M, n = _args(h, tnn, Es, sigL, sigR)
    out = np.empty((len(Es), M, n, n), dtype=np.complex128)
    check(_lib.load().tx_green_column0(ptr(h), ptr(tnn), n, M, ptr(Es), len(Es), ptr(sigL), ptr(sigR), ptr(out)))
    return out


def green_full(h, tnn, Es, sigL, sigR):
    """inv(E - H') as [nE, M, M, n, n]."""
    h, tnn, Es, sigL, sigR, M, n = _args(h, tnn, Es, sigL, sigR)
    out = np.empty((len(Es), M, M, n, n), dtype=np.complex128)
    check(_lib.load().tx_green_full(ptr(h), ptr(tnn), n, M, ptr(Es), len(Es), ptr(sigL), ptr(sigR), ptr(out)))
    return out


def hmat(h, tnn, tnnn):
    """Dense H[(M n), (M n)] from the block arrays."""
    h, tnn, tnnn = as_c128(h), as_c128(tnn), as_c128(tnnn)
    M, n = h.shape[0], h.shape[-1]
    out = np.empty((M * n, M * n), dtype=np.complex128)
    check(_lib.load().tx_hmat(ptr(h), ptr(tnn), ptr(tnnn) if M > 2 else None, n, M, ptr(out)))
    return out
