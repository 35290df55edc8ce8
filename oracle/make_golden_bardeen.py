"""
TEST INFRASTRUCTURE.  Generates tests/golden/bardeen_*.npz by running the UNMODIFIED reference's
`transport.bardeen` (imported from /root/reference through oracle/ref_loader.py) at the driver
parameters of BASELINE config 5: rbbarrier.py:45-96, rbstep.py:43-104, rbmenez_nosup_el.py:55-140 and the
bias sweep of rbbarrier.py:447-493.  Runs only in the build container; the fixtures are committed.

    python -m oracle.make_golden_bardeen
"""
import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def chain_HC(NC, VC, tC, n, extra=None):
    HC = np.zeros((NC, NC, n, n), dtype=complex)
    for j in range(NC):
        HC[j, j] = VC if extra is None else VC + extra
    for j in range(NC - 1):
        HC[j, j + 1] = -tC
        HC[j + 1, j] = -tC
    return HC


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    bardeen = load_reference(with_bardeen=True)["bardeen"]
    os.makedirs(OUT, exist_ok=True)

    # ---- rbbarrier.py:45-96 : spinless rectangular barrier, kernel_well + Ts_bardeen -----------
    n = 1
    tLR = 1.0 * np.eye(n); tinfty = 1.0 * tLR; VLR = 0.0 * tLR; Vinfty = 0.5 * tLR; Ninfty = 20
    VC = 0.4 * tLR
    HC = chain_HC(11, VC, 1.0 * tLR, n)
    for NLR in (50, 200):
        t0 = time.time()
        E, M = quiet(bardeen.kernel_well, tinfty, tLR, tLR, Vinfty, VLR, Vinfty, VLR, Vinfty, Ninfty, NLR, NLR,
                     HC, np.copy(HC), np.array([[1]]), E_cutoff=VC, verbose=0)
        secs = time.time() - t0
        T = quiet(bardeen.Ts_bardeen, E, M, tLR, tLR, VLR, VLR, NLR, NLR)
        np.savez(os.path.join(OUT, "bardeen_barrier_N%d.npz" % NLR), tinfty=tinfty, tL=tLR, tR=tLR, Vinfty=Vinfty,
                 VL=VLR, VLprime=Vinfty, VR=VLR, VRprime=Vinfty, Ninfty=Ninfty, NL=NLR, NR=NLR, HC=HC, HCobs=HC,
                 defines_Sz=np.array([[1]]), E_cutoff=VC, interval=1e-12, E=E, M=M, T=T, ref_seconds=secs)
        print("barrier", NLR, E.shape, "%.2fs" % secs, T[0, :3, 0])

    # ---- rbstep.py:43-104 : potential step in the right well, interval 1e-2 ---------------------
    VL = 0.0 * tLR
    for tag, VRs in (("m005", -0.005), ("p05", 0.05), ("m05", -0.05)):
        VR = VRs * tLR
        NLR = 200
        t0 = time.time()
        E, M = quiet(bardeen.kernel_well, tinfty, tLR, tLR, Vinfty, VL, VR + Vinfty, VR, Vinfty, Ninfty, NLR, NLR,
                     HC, np.copy(HC), np.copy(tLR), E_cutoff=0.1 * np.eye(n), interval=1e-2, verbose=0)
        secs = time.time() - t0
        T = quiet(bardeen.Ts_bardeen, E, M, tLR, tLR, VL, VR, NLR, NLR)
        np.savez(os.path.join(OUT, "bardeen_step_%s.npz" % tag), tinfty=tinfty, tL=tLR, tR=tLR, Vinfty=Vinfty,
                 VL=VL, VLprime=VR + Vinfty, VR=VR, VRprime=Vinfty, Ninfty=Ninfty, NL=NLR, NR=NLR, HC=HC, HCobs=HC,
                 defines_Sz=np.copy(tLR), E_cutoff=0.1 * np.eye(n), interval=1e-2, E=E, M=M, T=T, ref_seconds=secs)
        print("step", VRs, E.shape, "%.2fs" % secs, M[0, :3, 0])

    # ---- rbmenez_nosup_el.py:55-140 : n_loc_dof = 2, Kondo spin-flip term in Hsys - HL ----------
    n = 2
    tLR = 1.0 * np.eye(n); tinfty = 1.0 * tLR; VLR = 0.0 * tLR; Vinfty = 0.5 * tLR
    VC = 0.4 * tLR
    for tag, Jval, NLR in (("J05", -0.05, 200), ("J5", -0.5, 60)):
        kondo = (Jval / 2) * np.array([[0, 1], [1, 0]])
        HC = chain_HC(11, VC, 1.0 * tLR, n, extra=kondo)
        HCprime = chain_HC(11, VC, 1.0 * tLR, n)
        t0 = time.time()
        E, M = quiet(bardeen.kernel_well_prime, tinfty, tLR, tLR, Vinfty, VLR, Vinfty, VLR, Vinfty, Ninfty, NLR, NLR,
                     HC, HCprime, E_cutoff=0.4 * np.eye(n), verbose=0)
        secs = time.time() - t0
        T = quiet(bardeen.Ts_bardeen, E, M, tLR, tLR, VLR, VLR, NLR, NLR)
        Vb = np.linspace(-0.05, 0.05, 99)                    # rbbarrier.py:447-450
        I1 = quiet(bardeen.current, E, M, Vb, tLR, 0.0, 0.0001)
        I2 = quiet(bardeen.current, E, M, Vb, tLR, -1.8, 0.01)
        np.savez(os.path.join(OUT, "bardeen_prime_%s.npz" % tag), tinfty=tinfty, tL=tLR, tR=tLR, Vinfty=Vinfty,
                 VL=VLR, VLprime=Vinfty, VR=VLR, VRprime=Vinfty, Ninfty=Ninfty, NL=NLR, NR=NLR, HC=HC,
                 HCprime=HCprime, E_cutoff=0.4 * np.eye(n), interval=1e-9, E=E, M=M, T=T, Vb=Vb,
                 I_mu0_kT1em4=I1, I_mum18_kT1em2=I2, ref_seconds=secs)
        print("prime", Jval, NLR, E.shape, "%.2fs" % secs, M[:, 0, :].ravel(), np.abs(I2).max())

    # ragged bound-state counts (channel-dependent cutoff) and asymmetric wells
    HC = chain_HC(5, VC, 1.0 * tLR, n, extra=0.1 * np.array([[0.3, 1], [1, -0.2]]))
    HCprime = chain_HC(5, VC, 1.0 * tLR, n, extra=0.1 * np.array([[0.3, 0], [0, -0.2]]))
    Ecut = np.diag([0.3, 0.2])
    E, M = quiet(bardeen.kernel_well_prime, tinfty, tLR, tLR, Vinfty, VLR, Vinfty, 0.002 * np.eye(n), Vinfty, 7, 40, 43,
                 HC, HCprime, E_cutoff=Ecut, interval=3e-3, verbose=0)
    np.savez(os.path.join(OUT, "bardeen_prime_ragged.npz"), tinfty=tinfty, tL=tLR, tR=tLR, Vinfty=Vinfty,
             VL=VLR, VLprime=Vinfty, VR=0.002 * np.eye(n), VRprime=Vinfty, Ninfty=7, NL=40, NR=43, HC=HC,
             HCprime=HCprime, E_cutoff=Ecut, interval=3e-3, E=E, M=M)
    print("ragged", E.shape, np.count_nonzero(M), M.max())


if __name__ == "__main__":
    main()
