"""
TEST INFRASTRUCTURE.  Round-2 additions to tests/golden/: outputs of the UNMODIFIED reference (imported from
/root/reference through oracle/ref_loader.py).  Runs only in the build container.

    python -m oracle.make_golden_r2 [name ...]
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=SyntaxWarning)

from oracle.make_golden import sweep  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402
from transport_b200 import configs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cicc_spin32(ref):
    """Two spin-3/2 impurities: n_loc_dof = 2 (2s+1)^2 = 32, the large-block kernel's size class, straight from
    run_wfm_gate.get_hblocks (run_wfm_gate.py:80-143)."""
    wfm, gate = ref["wfm"], ref["run_wfm_gate"]
    h, tnn, tnnn = gate.get_hblocks(3, 1.0, -0.3, 0.05, 0.0, 0.0, 1, the_dist=1)
    hb, tb, _ = configs.cicc_blocks(3, 1.0, -0.3, 0.05, 0.0, 0.0, 1, 1)
    assert np.allclose(h, hb, atol=0, rtol=0) and np.array_equal(tnn, tb), "cicc builder deviates from reference"
    assert h.shape[-1] == 32
    Es = np.array([-1.95, -1.4, -0.6], dtype=complex)
    srcs = np.eye(32)[[0, 13, 31]]
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Es, "g_closed", list(srcs), all_debug=True)
    np.savez(os.path.join(OUT, "cicc_spin32_n32.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Es, sources=srcs, R=R, T=T)


def cicc_large_spins(ref):
    """Two spin-2, 5/2, 3, 7/2 impurities: n_loc_dof = 2 (2s+1)^2 = 50, 72, 98, 128 -- block edges beyond 64 (the 512-thread
    instantiation of the large-block kernel) and edges that are not a multiple of the 8 x 8 tensor-core tile (padded work
    blocks), straight from run_wfm_gate.get_hblocks (run_wfm_gate.py:80-143).  The blocks are not stored (1 MB):
    the fixture keeps the builder arguments, and this script asserts that transport_b200.configs.cicc_blocks reproduces
    the reference's blocks bit for bit."""
    wfm, gate = ref["wfm"], ref["run_wfm_gate"]
    for TwoS in (4, 5, 6, 7):
        args = (TwoS, 1.0, -0.3, 0.05, 0.0, 0.0, 1)
        h, tnn, tnnn = gate.get_hblocks(*args, the_dist=1)
        hb, tb, _ = configs.cicc_blocks(*args, 1)
        assert np.allclose(h, hb, atol=0, rtol=0) and np.array_equal(tnn, tb), "cicc builder deviates from reference"
        n = h.shape[-1]
        assert n == 2 * (TwoS + 1) ** 2
        Es = np.array([-1.95, -1.4, -0.6], dtype=complex)
        idx = [0, n // 3 + 1, n - 1]
        srcs = np.eye(n)[idx]
        R, T = sweep(wfm, h, tnn, tnnn, 1.0, Es, "g_closed", list(srcs), all_debug=True)
        np.savez(os.path.join(OUT, "cicc_builder_n%d.npz" % n), builder_args=np.array(args), the_dist=1, Es=Es, sources=srcs, R=R, T=T)


def nnn_chains(ref):
    """Nonzero next-nearest-neighbour blocks through the reference's own pentadiagonal assembly (wfm.Hmat :209-213):
    kernel (R, T), the wavefunction output and the full Green's function, even and odd chain lengths."""
    wfm = ref["wfm"]
    rng = np.random.default_rng(20261020)
    for n, M in ((2, 6), (2, 7), (3, 5), (1, 8)):
        A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
        h = 0.3 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
        h[0] = 0
        h[-1] = np.diag(0.05 * rng.standard_normal(n))
        tnn = np.tile(-np.eye(n, dtype=complex), (M - 1, 1, 1))
        tnn[1:-1] += 0.1 * (rng.standard_normal((M - 3, n, n)) + 1j * rng.standard_normal((M - 3, n, n)))
        tnnn = 0.15 * (rng.standard_normal((M - 2, n, n)) + 1j * rng.standard_normal((M - 2, n, n)))
        Es = np.array([-1.7, -0.9, 0.4, 1.3], dtype=complex)
        srcs = np.eye(n)
        R, T = sweep(wfm, h, tnn, tnnn, 1.0, Es, "g_closed", list(srcs), all_debug=True)
        psi = np.array([wfm.kernel(h, tnn, tnnn, 1.0, E, "g_closed", srcs[0], True, False, all_debug=False) for E in Es])
        G = np.array([wfm.Green(h, tnn, tnnn, 1.0, E, "g_closed") for E in Es])
        np.savez(os.path.join(OUT, "nnn_n%d_M%d.npz" % (n, M)), h=h, tnn=tnn, tnnn=tnnn, Es=Es, sources=srcs, R=R, T=T,
                 psi=psi, G=G)


def well_channel_potentials(ref):
    """Blocks as bardeen.Ts_wfm_well builds them (:650-669) with a channel-dependent lead potential VL = VR = diag(0, Delta)
    (rbmenez_sup_inel.py:135, rbmenez_spinless_inel.py:147) through the reference's wfm.kernel.  The reference asserts
    that the SOURCE channel's lead potential is 0 (wfm/__init__.py:67-70), so only alpha = 0 can be generated."""
    wfm = ref["wfm"]
    n, NC, Delta = 2, 5, 0.3
    HC = np.zeros((NC, NC, n, n), dtype=complex)
    for j in range(NC):
        HC[j, j] = 0.4 * np.eye(n) + 0.1 * np.array([[0, 1], [1, 0]])
    for j in range(NC - 1):
        HC[j, j + 1] = HC[j + 1, j] = -1.0 * np.eye(n)
    tL = 1.0 * np.eye(n)
    VLR = np.diag([0.0, Delta])
    hblocks = np.zeros((NC + 2, n, n), dtype=complex)
    tnn = np.zeros((NC + 1, n, n), dtype=complex)
    tnnn = np.zeros((NC, n, n), dtype=complex)
    hblocks[0] = VLR * np.eye(n)
    tnn[0] = -tL * np.eye(n)
    for i in range(NC):
        hblocks[1 + i] = HC[i, i]
        if i + 1 < NC:
            tnn[1 + i] = HC[i, i + 1]
    hblocks[-1] = VLR * np.eye(n)
    tnn[-1] = -tL * np.eye(n)
    Emas = np.array([[-1.8, -1.3, -0.6, 0.2], [-1.5, -1.0, -0.4, 0.5]])
    T0 = np.array([wfm.kernel(hblocks, tnn, tnnn, 1.0, complex(E), "g_closed", np.eye(n)[0], False, False, all_debug=False)[1]
                   for E in Emas[0]])
    np.savez(os.path.join(OUT, "well_channel_potentials.npz"), HC=HC, tL=tL, VLR=VLR, Emas=Emas, T_alpha0=T0)


def _load_fcdmft():
    """The unmodified /root/reference/fcdmft/__init__.py with stubs for matplotlib and for its pyscf-dependent subpackage
    `fcdmft.dmft` (landauer with diagonal leads takes the closed-form branch of surface_gf and never touches it)."""
    import importlib.util
    import types
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    root = os.path.join(os.environ.get("TRANSPORT_REFERENCE_ROOT", "/root/reference"), "fcdmft")
    spec = importlib.util.spec_from_file_location("fcdmft", os.path.join(root, "__init__.py"), submodule_search_locations=[root])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fcdmft"] = mod
    stub = types.ModuleType("fcdmft.dmft")
    stub.__path__ = []
    sys.modules["fcdmft.dmft"] = stub
    spec.loader.exec_module(mod)
    return mod


def landauer_fcdmft(ref):
    """Matrix-Gamma Landauer current of a noninteracting two-orbital region (the reference hard-codes two orbitals at :273)
    between two 1D leads through the reference's
    `fcdmft.landauer` (:224-292): energies, the retarded Green's function handed to it, its jE(E) at two temperatures and
    the trapezoid integrals.  The same system as a wfm chain: blocks [lead cell, lead cell, region, lead cell, lead cell],
    hoppings [V, c_L, c_R, V]."""
    fc = _load_fcdmft()
    rng = np.random.default_rng(7)
    n = 2
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    h_SR = 0.4 * (A + A.conj().T) / 2
    # real symmetric couplings, strong enough that |G| < 1: the reference's decompose_gf asserts an identity to an ABSOLUTE
    # 1e-15 (:569), which rounding breaks for larger Green's functions
    cL = 2.0 * (np.diag([0.8, 0.6]) + 0.1 * np.array([[0, 1], [1, 0]]))
    cR = 2.0 * (np.diag([0.7, 0.9]) + 0.05 * np.array([[0, 1], [1, 0]]))
    V = 1.0
    energies = np.linspace(-1.8, 1.8, 240)      # even count: E = 0 would divide by zero in the reference (:337)
    eye = np.eye(n)
    LL = (0.0 * eye[None], V * eye[None], cL[None], 0.35)
    RL = (0.0 * eye[None], V * eye[None], cR[None], -0.25)
    gL = fc.surface_gf(energies, 0.0, LL[0], LL[1])
    gR = fc.surface_gf(energies, 0.0, RL[0], RL[1])
    G = np.zeros((1, n, n, len(energies)), dtype=complex)
    for iw, E in enumerate(energies):
        SigL = cL @ gL[0, :, :, iw] @ cL
        SigR = cR @ gR[0, :, :, iw] @ cR
        G[0, :, :, iw] = np.linalg.inv(E * eye - h_SR - SigL - SigR)
    out = {}
    for tag, kBT in (("kT0", 0.0), ("kT005", 0.05)):
        jE = fc.landauer(energies, 0.0, kBT, G, LL, RL)
        out["jE_" + tag] = jE
        out["I_" + tag] = np.trapezoid(np.real(jE), energies)
    h = np.zeros((5, n, n), dtype=complex)
    h[2] = h_SR
    tnn = np.array([V * eye, cL, cR, V * eye], dtype=complex)
    np.savez(os.path.join(OUT, "landauer_fcdmft.npz"), energies=energies, MBGF=G, h_SR=h_SR, cL=cL, cR=cR, V=V, muL=0.35,
             muR=-0.25, h=h, tnn=tnn, **out)


def kernel_well_n2(ref):
    """bardeen.kernel_well with n_loc_dof = 2 and a separate observable basis (HCobs != HC), the call of
    rbmenez_spinless_el.py:45-127 (HC = V_C + (J/2) sigma_x on the central sites, HCobs without the spin mixing).
    With exactly degenerate observable channels (HCobs = V_C 1) the reference classifies the ROWS of its basis-changed
    matrix with the right-well list's <S_z> sequence (:213-227) while they index left-well states (:166), and LAPACK's order
    inside a degenerate pair is arbitrary: its split between the two spin channels is then an artefact of the eigensolver
    (totals are not).  Fixtures: two cases with a small Zeeman term in HCobs (non-degenerate, ordering unique) and one
    degenerate case in which the two lists happen to come out in the same order."""
    import contextlib
    import io
    bardeen = ref["bardeen"]
    n = 2
    tLR = 1.0 * np.eye(n)
    Vinfty = 0.5 * tLR
    VLR = 0.0 * tLR
    sx, sz = np.array([[0, 1], [1, 0]]), np.diag([1.0, -1.0])

    def chains(J, B):
        NC = 11
        HC = np.zeros((NC, NC, n, n), dtype=complex)
        HCobs = np.zeros_like(HC)
        for i in range(NC):
            HCobs[i, i] = 0.4 * tLR + B * sz
            HC[i, i] = HCobs[i, i] + (J / 2) * sx
            if i + 1 < NC:
                HC[i, i + 1] = HC[i + 1, i] = HCobs[i, i + 1] = HCobs[i + 1, i] = -1.0 * tLR
        return HC, HCobs
    for tag, J, B, NLR, Ninf in (("zeeman_J5", -0.5, 0.03, 40, 10), ("zeeman_J2", -0.2, -0.05, 36, 8), ("degenerate_J05", -0.05, 0.0, 30, 8)):
        HC, HCobs = chains(J, B)
        with contextlib.redirect_stdout(io.StringIO()):
            E, M = bardeen.kernel_well(1.0 * tLR, tLR, tLR, Vinfty, VLR, Vinfty, VLR, Vinfty, Ninf, NLR, NLR, HC, HCobs, sz,
                                       E_cutoff=0.1 * np.eye(n), verbose=0)
        np.savez(os.path.join(OUT, "bardeen_well_n2_%s.npz" % tag), tinfty=1.0 * tLR, tL=tLR, tR=tLR, Vinfty=Vinfty, VL=VLR,
                 VLprime=Vinfty, VR=VLR, VRprime=Vinfty, Ninfty=Ninf, NL=NLR, NR=NLR, HC=HC, HCobs=HCobs, defines_Sz=sz,
                 E_cutoff=0.1 * np.eye(n), interval=1e-12, E=E, M=M)
        print(tag, E.shape, M.shape)


MAKERS = {"kernel_well_n2": kernel_well_n2, "cicc_spin32": cicc_spin32, "cicc_large_spins": cicc_large_spins, "nnn_chains": nnn_chains, "well_channel_potentials": well_channel_potentials,
          "landauer_fcdmft": landauer_fcdmft}


def main(names):
    ref = load_reference(with_bardeen=True, with_scripts=True)
    for name in names or sorted(MAKERS):
        MAKERS[name](ref)
        print("wrote", name)


if __name__ == "__main__":
    main(sys.argv[1:])
