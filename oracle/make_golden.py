"""
TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_loader.py) on seeded inputs.  Runs only in
the build container; the fixtures are committed so the GPU box never needs the reference.

    python -m oracle.make_golden
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=SyntaxWarning)

from oracle.ref_loader import load_reference  # noqa: E402
from transport_b200 import configs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sweep(wfm, h, tnn, tnnn, tl, Es, converger, sources, all_debug=False):
    n = h.shape[-1]
    R = np.zeros((len(Es), len(sources), n))
    T = np.zeros_like(R)
    for i, E in enumerate(Es):
        for s, src in enumerate(sources):
            R[i, s], T[i, s] = wfm.kernel(h, tnn, tnnn, tl, E, converger, src, False, False,
                                          all_debug=all_debug, verbose=0)
    return R, T


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference(with_bardeen=True, with_scripts=True)
    wfm, gate, bardeen = ref["wfm"], ref["run_wfm_gate"], ref["bardeen"]
    rng = np.random.default_rng(20261019)

    # ---- cfg1: barrier ------------------------------------------------------------------
    h, tnn, tnnn = configs.barrier_blocks(Vb=0.1, NC=14)
    Es = np.logspace(-4, 0, 199).astype(complex) - 2.0
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Es, "g_closed", [np.array([1.0])], all_debug=True)
    np.savez(os.path.join(OUT, "barrier_run_wfm.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Es, R=R, T=T)

    h, tnn, tnnn = configs.barrier_blocks(Vb=0.4, NC=11)
    Es_full = configs.barrier_energies(10_000)
    Es = Es_full[::50]
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Es, "g_closed", [np.array([1.0])], all_debug=True)
    np.savez(os.path.join(OUT, "cfg1_barrier.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Es, stride=50, R=R, T=T)

    # ---- cfg2: two spin-1/2 impurities, n=8, M=8 -------------------------------------------
    Js_full, Es_full = configs.cfg2_grid(1000, 1000)
    Jsel = Js_full[[0, 250, 500, 750, 999]]
    Esel = Es_full[::40]
    srcs = [np.eye(8)[a] for a in range(8)]
    Rs, Ts, hs = [], [], []
    for J in Jsel:
        h, tnn, tnnn = gate.get_hblocks(1, 1.0, J, 0.0, 0.0, 0.0, 3, the_dist=2)
        hb, tb, _ = configs.cicc_blocks(1, 1.0, J, 0.0, 0.0, 0.0, 3, 2)
        assert np.array_equal(h, hb) and np.array_equal(tnn, tb), "cicc builder deviates from reference"
        R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_closed", srcs, all_debug=True)
        Rs.append(R), Ts.append(T), hs.append(h)
    np.savez(os.path.join(OUT, "cfg2_cicc.npz"), Js=Jsel, Es=Esel, h=np.array(hs), tnn=tnn, tnnn=tnnn,
             R=np.array(Rs), T=np.array(Ts))
    # the survey's spot check: M=5, source 1, E=-1.99
    h, tnn, tnnn = gate.get_hblocks(1, 1.0, -0.5, 0, 0, 0.0, 1, the_dist=1)
    R, T = wfm.kernel(h, tnn, tnnn, 1.0, complex(-1.99), "g_closed", np.eye(8)[1], False, False, all_debug=True)
    np.savez(os.path.join(OUT, "cicc_spot.npz"), h=h, tnn=tnn, tnnn=tnnn, E=complex(-1.99), R=R, T=T)
    # spin-1 impurities (n = 18)
    h, tnn, tnnn = gate.get_hblocks(2, 1.0, -0.4, 0.1, 0.0, 0.0, 2, the_dist=1)
    hb, tb, _ = configs.cicc_blocks(2, 1.0, -0.4, 0.1, 0.0, 0.0, 2, 1)
    assert np.allclose(h, hb, atol=0, rtol=0) and np.array_equal(tnn, tb)
    Esel = np.array([-1.9, -1.2, -0.3], dtype=complex)
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_closed", [np.eye(18)[0], np.eye(18)[7]], all_debug=True)
    np.savez(os.path.join(OUT, "cicc_spin1_n18.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Esel, R=R, T=T)

    # ---- cfg3 (down-scaled): disordered chain, n=8 ----------------------------------------
    for N, W in ((6, 0.05), (30, 0.05), (62, 0.5)):
        h, tnn, tnnn = configs.disordered_blocks(N=N, n=8, W=W, seed=1234)
        Esel = np.linspace(-1.9, 1.9, 7).astype(complex)
        srcs = [np.eye(8)[0], np.eye(8)[5]]
        R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_closed", srcs, all_debug=True)
        np.savez(os.path.join(OUT, "cfg3_disorder_N%d.npz" % N), h=h, tnn=tnn, tnnn=tnnn, Es=Esel,
                 sources=np.array(srcs), R=R, T=T)

    # ---- general (matrix) hopping in the scattering region, assorted block sizes -----------
    for n in (1, 2, 3, 4, 5, 6, 8, 12, 16):
        M = 6
        A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
        h = 0.3 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
        h[0] = 0
        h[-1] = np.diag(0.1 * rng.standard_normal(n))
        tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
        tnn[1:-1] += 0.2 * (rng.standard_normal((M - 3, n, n)) + 1j * rng.standard_normal((M - 3, n, n)))
        tnnn = np.zeros((M - 2, n, n), dtype=complex)
        Esel = np.array([-1.7, -0.9, 0.2, 1.1], dtype=complex)
        srcs = [np.eye(n)[0], np.ones(n) / np.sqrt(n)]
        R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_closed", srcs, all_debug=True)
        np.savez(os.path.join(OUT, "general_hop_n%d.npz" % n), h=h, tnn=tnn, tnnn=tnnn, Es=Esel,
                 sources=np.array(srcs), R=R, T=T)

    # lead hopping that is diagonal but channel dependent and != 1, plus a non-diagonal tnn[0]
    n, M = 4, 5
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = 0.2 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
    h[0] = np.diag([0.0, 0.0, 0.05, -0.05])
    h[-1] = np.diag([0.0, 0.1, 0.0, -0.1])
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    tnn[0] = np.diag([-1.0, -0.8, -1.2, -1.0])
    tnn[-1] = np.diag([-1.0, -1.1, -0.9, -1.0])
    tnn[0][0, 1] = 0.1
    tnn[-1][2, 3] = -0.05j
    tnnn = np.zeros((M - 2, n, n), dtype=complex)
    Esel = np.array([-1.5, -0.4, 0.7, 2.3], dtype=complex)
    srcs = [np.eye(n)[0], np.eye(n)[1]]
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_closed", srcs, all_debug=False)
    SL = np.array([wfm.SelfEnergies(h, tnn, tnnn, E, "g_closed")[0] for E in Esel])
    SR = np.array([wfm.SelfEnergies(h, tnn, tnnn, E, "g_closed")[1] for E in Esel])
    np.savez(os.path.join(OUT, "nondiag_lead_hop_n4.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Esel,
             sources=np.array(srcs), R=R, T=T, SigL=SL, SigR=SR)

    # ---- Rice-Mele leads (n=4, M=3), `run_wfm_single.py:82-200` ------------------------------
    h, tnn, tnnn = configs.ricemele_kondo_blocks(J=-0.5, u=0.0, v=-1.0, w=-1.0)
    Esel = (np.logspace(-4, 0, 25) - 2.0).astype(complex)
    srcs = [np.eye(4)[0]]
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_RiceMele", srcs, all_debug=True)
    np.savez(os.path.join(OUT, "ricemele_kondo.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Esel,
             sources=np.array(srcs), R=R, T=T)
    h, tnn, tnnn = configs.ricemele_kondo_blocks(J=-0.05, u=0.0, v=-0.8, w=-1.1)
    Esel = np.linspace(-1.85, -0.35, 16).astype(complex)
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_RiceMele", srcs, all_debug=True)
    np.savez(os.path.join(OUT, "ricemele_kondo_vw.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Esel,
             sources=np.array(srcs), R=R, T=T)

    # ---- kondo single impurity, monatomic, n=2 ------------------------------------------------
    h, tnn, tnnn = configs.kondo_blocks(J=-0.5)
    Esel = (np.logspace(-6, 0, 31) - 2.0).astype(complex)
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, "g_closed", [np.eye(2)[0], np.eye(2)[1]], all_debug=True)
    np.savez(os.path.join(OUT, "kondo_n2.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Esel, R=R, T=T)

    # ---- iterative converger through kernel (n<=2) --------------------------------------------
    Esel = np.array([-1.6, -0.7, 0.4], dtype=complex)
    conv = (1e-2, 1e-3)
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Esel, conv, [np.eye(2)[0]], all_debug=False)
    np.savez(os.path.join(OUT, "kondo_n2_iterative.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Esel,
             imE=conv[0], conv_tol=conv[1], R=R, T=T)

    # ---- surface Green's functions ------------------------------------------------------------
    Es = np.linspace(-2.6, 2.6, 53).astype(complex)
    diag = np.diag([0.0, 0.3, -0.2]).astype(complex)
    off = np.diag([-1.0, -1.0, -1.0]).astype(complex)
    off2 = np.diag([-1.0, -0.7, -1.4]).astype(complex)
    gc = np.array([wfm.g_closed(diag, off, E, 1) for E in Es])
    gc2 = np.array([wfm.g_closed(diag, off2, E, -1) for E in Es])
    Ec = Es + 0.05j
    gc3 = np.array([wfm.g_closed(diag, off, E, 1) for E in Ec])
    np.savez(os.path.join(OUT, "g_closed.npz"), Es=Es, diag=diag, off=off, off2=off2, g=gc, g2=gc2,
             Ec=Ec, g3=gc3)

    rm_diag = np.array([[0.2, -0.8], [-0.8, -0.2]], dtype=complex)
    rm_off = np.array([[0, 0], [-1.1, 0]], dtype=complex)
    Es = np.linspace(-2.2, 2.2, 45).astype(complex)
    gl = np.array([wfm.g_RiceMele(rm_diag, rm_off, E, -1) for E in Es])
    gr = np.array([wfm.g_RiceMele(rm_diag, rm_off, E, 1) for E in Es])
    np.savez(os.path.join(OUT, "g_ricemele.npz"), Es=Es, diag=rm_diag, off=rm_off, gL=gl, gR=gr)

    # g_ith chain and g_iter with its stopping rule (`surfacegf.py:189-222` usage)
    h00 = np.array([[0.1, -0.6], [-0.6, -0.1]], dtype=complex)
    h01 = np.array([[0.0, 0.0], [-1.0, 0.0]], dtype=complex)
    Es = np.linspace(-1.8, 1.8, 7).astype(complex)
    chain = np.zeros((len(Es), 2, 12, 2, 2), dtype=complex)
    for i, E in enumerate(Es):
        for si, side in enumerate((-1, 1)):
            g = np.zeros_like(h00)
            for it in range(12):
                g = wfm.g_ith(h00, h01, E, side, 1e-2, it, g)
                chain[i, si, it] = g
    giter = np.zeros((len(Es), 2, 2, 2), dtype=complex)
    gconv = np.zeros((len(Es), 2))
    gnit = np.zeros((len(Es), 2), dtype=int)
    for i, E in enumerate(Es):
        for si, side in enumerate((-1, 1)):
            g, c, nit = wfm.g_iter(h00, h01, E, side, 1e-2, 1e-3, full=True)
            giter[i, si], gconv[i, si], gnit[i, si] = g, c, nit
    np.savez(os.path.join(OUT, "g_iter.npz"), Es=Es, h00=h00, h01=h01, imE=1e-2, conv_tol=1e-3,
             chain=chain, g=giter, conv=gconv, n_iter=gnit)

    # ---- Hmat / Hprime / Green / psi ----------------------------------------------------------
    n, M = 3, 5
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = 0.3 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
    h[0] = 0
    h[-1] = 0
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    tnn[1:-1] += 0.1 * rng.standard_normal((M - 3, n, n))
    tnnn = 0.05 * (rng.standard_normal((M - 2, n, n)) + 1j * rng.standard_normal((M - 2, n, n)))
    E = complex(-0.8)
    H = wfm.Hmat(h, tnn, tnnn)
    Hp = wfm.Hprime(h, tnn, tnnn, 1.0, E, "g_closed")
    G = wfm.Green(h, tnn, tnnn, 1.0, E, "g_closed")
    tz = np.zeros_like(tnnn)
    G0 = wfm.Green(h, tnn, tz, 1.0, E, "g_closed")
    src = np.array([1.0, 0.0, 0.5])
    psi = wfm.kernel(h, tnn, tz, 1.0, E, "g_closed", src, True, False, all_debug=False)
    Rn, Tn = wfm.kernel(h, tnn, tnnn, 1.0, E, "g_closed", src, False, False, all_debug=False)
    np.savez(os.path.join(OUT, "green_n3.npz"), h=h, tnn=tnn, tnnn=tnnn, E=E, H=H, Hp=Hp, G=G, G_nn_only=G0,
             src=src, psi=psi, R_tnnn=Rn, T_tnnn=Tn)

    # ---- bardeen.Hsysmat ---------------------------------------------------------------------
    n = 2
    tinf, tL, tR = 1.0 * np.eye(n), 0.9 * np.eye(n), 1.1 * np.eye(n)
    Vinf, VL, VR = 0.5 * np.eye(n), 0.01 * np.eye(n), 0.02 * np.eye(n)
    HC = np.zeros((3, 3, n, n), dtype=complex)
    for j in range(3):
        HC[j, j] = 0.4 * np.eye(n) + 0.05 * np.array([[0, 1], [1, 0]])
    for j in range(2):
        HC[j, j + 1] = -1.0 * np.eye(n)
        HC[j + 1, j] = -1.0 * np.eye(n)
    Hs = bardeen.Hsysmat(tinf, tL, tR, Vinf, VL, VR, 4, 6, 5, HC)
    np.savez(os.path.join(OUT, "hsysmat.npz"), tinf=tinf, tL=tL, tR=tR, Vinf=Vinf, VL=VL, VR=VR,
             Ninf=4, NL=6, NR=5, HC=HC, H=Hs)
    print("golden fixtures written to", OUT)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("total bytes", tot)


if __name__ == "__main__":
    main()
