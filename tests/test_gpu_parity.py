"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
(1) the golden vectors produced by the unmodified reference and (2) the numpy oracle on seeded
inputs.  Bar (BASELINE.json north_star): relative error <= 1e-10 on every R/T channel and
|sum R + sum T - 1| <= 1e-10.
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rt_err
from oracle import rgf_oracle, wfm_oracle
from transport_b200 import _lib, configs, green, leads, wfm
from transport_b200.sweep import transmission_sweep

pytestmark = pytest.mark.gpu

RTOL = 1e-10      # the parity gate of north_star
UNIT = 1e-10      # unitarity gate (wfm/__init__.py:135)


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on a GPU box")


def _check(R, T, gR, gT, resid=None):
    assert rt_err(T, gT) < RTOL, rt_err(T, gT)
    assert rt_err(R, gR) < RTOL, rt_err(R, gR)
    if resid is not None:
        ok = np.isfinite(resid)
        assert np.max(np.abs(resid[ok])) < UNIT


# ------------------------------------------------------------------------------- golden vectors
def test_barrier_goldens_n1():
    for name in ("barrier_run_wfm.npz", "cfg1_barrier.npz"):
        g = load_golden(name)
        R, T, res = transmission_sweep(g["h"], g["tnn"], g["Es"], sources=[[1.0]], want_resid=True)
        _check(R[0], T[0], g["R"], g["T"], res)
        # T spans 1e-6..1 here; every value must hold to 1e-10 relative, not absolute
        assert np.max(np.abs(T[0] - g["T"]) / g["T"]) < RTOL


def test_cfg2_cicc_golden_all_sources():
    g = load_golden("cfg2_cicc.npz")
    # batched over the 5 Hamiltonians x 25 energies x all 8 one-hot sources in one launch
    R, T, res = transmission_sweep(g["h"], g["tnn"], g["Es"], want_resid=True)
    _check(R, T, g["R"], g["T"], res)
    # the affine form H(J) = H0 + J*H1 gives the same answer
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    R2, T2 = transmission_sweep(h0, tnn, g["Es"], h1=h1, params=g["Js"])
    _check(R2, T2, g["R"], g["T"])
    # explicit one-hot amplitude vectors take the general-source path
    R3, T3 = transmission_sweep(g["h"], g["tnn"], g["Es"], sources=np.eye(8)[[1, 6]])
    _check(R3, T3, g["R"][:, :, [1, 6]], g["T"][:, :, [1, 6]])


def test_cicc_spot_value_through_reference_signature():
    g = load_golden("cicc_spot.npz")
    Rs, Ts = wfm.kernel(g["h"], g["tnn"], g["tnnn"], 1.0, complex(g["E"]), "g_closed", np.eye(8)[1], False, False,
                        all_debug=True)
    assert Rs.dtype == np.float64 and Rs.shape == (8,)
    _check(Rs, Ts, g["R"], g["T"])


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                                        if os.path.basename(p).startswith(("general_hop", "cfg3_", "nondiag",
                                                                           "kondo_n2.npz", "cicc_spin"))))
def test_closed_lead_goldens(name):
    """Every closed-form-lead golden of the live reference, through the C ABI; cicc_spin1 (n = 18) and cicc_spin32
    (n = 32) pin the CTA-per-point large-block kernel to the reference itself."""
    g = load_golden(name)
    srcs = g["sources"] if "sources" in g.files else np.eye(g["h"].shape[-1])
    if name.startswith("cicc_spin1_"):
        srcs = np.eye(18)[[0, 7]]
    R, T, res = transmission_sweep(g["h"], g["tnn"], g["Es"], sources=srcs, want_resid=True)
    _check(R[0], T[0], g["R"], g["T"], res if not name.startswith("nondiag") else None)


@pytest.mark.parametrize("name", ["ricemele_kondo.npz", "ricemele_kondo_vw.npz"])
def test_ricemele_goldens(name):
    g = load_golden(name)
    for i, E in enumerate(g["Es"][::3]):
        Rs, Ts = wfm.kernel(g["h"], g["tnn"], g["tnnn"], 1.0, E, "g_RiceMele", g["sources"][0], False, False,
                            all_debug=True)
        _check(Rs, Ts, g["R"][3 * i, 0], g["T"][3 * i, 0])
    SL, SR = leads.self_energies_batched(g["h"], g["tnn"], g["Es"], "g_RiceMele")
    R, T = transmission_sweep(g["h"], g["tnn"], g["Es"], "table", sources=g["sources"], sigL=SL, sigR=SR)
    _check(R[0], T[0], g["R"], g["T"])


def test_iterative_converger_golden():
    g = load_golden("kondo_n2_iterative.npz")
    conv = (float(g["imE"]), float(g["conv_tol"]))
    for i, E in enumerate(g["Es"]):
        Rs, Ts = wfm.kernel(g["h"], g["tnn"], g["tnnn"], 1.0, E, conv, np.eye(2)[0], False, False, all_debug=False)
        assert rt_err(Ts, g["T"][i, 0]) < 1e-9 and rt_err(Rs, g["R"][i, 0]) < 1e-9


def test_surface_gf_goldens():
    g = load_golden("g_closed.npz")
    for i in range(0, len(g["Es"]), 4):
        E = g["Es"][i]
        assert np.allclose(wfm.g_closed(g["diag"], g["off"], E, 1), g["g"][i], rtol=1e-13, atol=1e-15)
        assert np.allclose(wfm.g_closed(g["diag"], g["off2"], E, -1), g["g2"][i], rtol=1e-13, atol=1e-15)
    (gc,) = leads.surface_gf(g["diag"], g["off"], g["Ec"], 1, _lib.TX_GF_CLOSED)
    assert np.allclose(gc, g["g3"], rtol=1e-13, atol=1e-15)
    g = load_golden("g_ricemele.npz")
    (gl,) = leads.surface_gf(g["diag"], g["off"], g["Es"], -1, _lib.TX_GF_RICEMELE)
    (gr,) = leads.surface_gf(g["diag"], g["off"], g["Es"], 1, _lib.TX_GF_RICEMELE)
    assert np.allclose(gl, g["gL"], rtol=1e-12, atol=1e-14)
    assert np.allclose(gr, g["gR"], rtol=1e-12, atol=1e-14)
    assert np.allclose(wfm.g_RiceMele(g["diag"], g["off"], g["Es"][3], -1), g["gL"][3], rtol=1e-12, atol=1e-14)


def test_g_ith_g_iter_goldens_and_sancho():
    g = load_golden("g_iter.npz")
    imE, tol = float(g["imE"]), float(g["conv_tol"])
    for si, side in enumerate((-1, 1)):
        cur = None
        for it in range(12):
            cur = leads.g_ith(g["h00"], g["h01"], g["Es"], side, imE, it, cur)
            assert np.allclose(cur, g["chain"][:, si, it], rtol=1e-12, atol=1e-14)
        (gg, conv, nit, ok) = leads.surface_gf(g["h00"], g["h01"], g["Es"], side, _lib.TX_GF_FIXEDPOINT,
                                               imE=imE, conv_tol=tol, full=True)
        assert ok.all()
        assert np.array_equal(nit, g["n_iter"][:, si])
        assert np.allclose(gg, g["g"][:, si], rtol=1e-11, atol=1e-14)
    out, c, nit = wfm.g_iter(g["h00"], g["h01"], g["Es"][2], 1, imE, tol, full=True)
    assert nit == g["n_iter"][2, 1] and np.allclose(out, g["g"][2, 1], rtol=1e-11)
    step = wfm.g_ith(g["h00"], g["h01"], g["Es"][0], -1, imE, 0, np.zeros((2, 2), complex))
    assert np.allclose(step, g["chain"][0, 0, 0], rtol=1e-12)
    with pytest.raises(Exception, match="failed to meet"):
        wfm.g_iter(g["h00"], g["h01"], g["Es"][2], 1, 1e-6, 1e-14, max_iter=300)
    # Sancho-Rubio: fixed point of the reference's own map ...
    for side in (-1, 1):
        (gs, conv, nit, ok) = leads.surface_gf(g["h00"], g["h01"], g["Es"], side, _lib.TX_GF_SANCHO,
                                               imE=imE, conv_tol=1e-14, max_iter=200, full=True)
        assert ok.all() and nit.max() < 40
        for i, E in enumerate(g["Es"]):
            nxt = wfm_oracle.g_ith(g["h00"], g["h01"], E, side, imE, 1, gs[i])
            assert np.linalg.norm(nxt - gs[i]) / np.linalg.norm(gs[i]) < 1e-12
        ref, _ = rgf_oracle.sancho_rubio(g["h00"], g["h01"], g["Es"] + 1j * imE, side)
        assert np.allclose(gs, ref, rtol=1e-11, atol=1e-13)


def test_sancho_ribbon_leads_match_closed_form_channels():
    """a square-lattice ribbon lead is separable, g = U diag(g_closed(eps_m)) U^T."""
    n = 16
    h, tnn, _ = configs.ribbon_blocks(width=n, L=2, W=0.0)
    h00, h01 = h[0], tnn[0]
    eps, U = np.linalg.eigh(h00.real)
    Es = np.array([-2.9, -1.3, 0.45, 2.2], dtype=complex)
    eta = 1e-9
    (gs, conv, nit, ok) = leads.surface_gf(h00, h01, Es, 1, _lib.TX_GF_SANCHO, imE=eta, conv_tol=1e-15,
                                           max_iter=200, full=True)
    assert ok.all()
    for i, E in enumerate(Es):
        gd = np.diagonal(wfm_oracle.g_closed(np.diag(eps).astype(complex), -np.eye(n, dtype=complex), E, 1))
        want = (U * gd) @ U.T
        assert np.max(np.abs(gs[i] - want)) < 1e-6     # O(sqrt(eta)) near band edges, exact as eta -> 0


def test_green_hmat_hprime_psi_golden():
    g = load_golden("green_n3.npz")
    h, tnn, tnnn, E = g["h"], g["tnn"], g["tnnn"], complex(g["E"])
    assert np.array_equal(wfm.Hmat(h, tnn, tnnn), g["H"])
    assert np.allclose(wfm.Hprime(h, tnn, tnnn, 1.0, E, "g_closed"), g["Hp"], rtol=1e-13, atol=1e-15)
    tz = np.zeros_like(tnnn)
    G = wfm.Green(h, tnn, tz, 1.0, E, "g_closed")
    assert G.shape == g["G_nn_only"].shape
    assert np.allclose(G, g["G_nn_only"], rtol=1e-10, atol=1e-13)
    psi = wfm.kernel(h, tnn, tz, 1.0, E, "g_closed", g["src"], True, False, all_debug=False)
    assert np.allclose(psi, g["psi"], rtol=1e-10, atol=1e-13)
    SL, SR = wfm.SelfEnergies(h, tnn, tz, E, "g_closed")
    oL, oR = wfm_oracle.SelfEnergies(h, tnn, tz, E, "g_closed")
    assert np.allclose(SL, oL, rtol=1e-13, atol=1e-15) and np.allclose(SR, oR, rtol=1e-13, atol=1e-15)
    with pytest.raises(NameError):
        wfm.kernel(h, tnn, tz, 1.0, E, "g_closed", g["src"], False, True, all_debug=False)


# ------------------------------------------------------------------------------- seeded vs oracle
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8, 13, 16])
def test_device_block_inverse(n):
    """The sweep's in-register Gauss-Jordan (implicit partial pivoting) against numpy's LU inverse."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((101, n, n)) + 1j * rng.standard_normal((101, n, n))
    A[7] = np.eye(n)[::-1]                        # permutation matrix: every pivot is off-diagonal
    A[8] = np.diag(10.0 ** np.linspace(-6, 6, n)) @ A[8]
    out = np.empty_like(A)
    _lib.check(_lib.load().tx_debug_invert(_lib.ptr(A), _lib.ptr(out), len(A), n))
    want = np.linalg.inv(A)
    scale = np.abs(want).max(axis=(1, 2), keepdims=True)
    assert np.max(np.abs(out - want) / scale) < 1e-10 * max(1.0, np.linalg.cond(A).max() * 1e-4)
    assert np.abs(out[7] - A[7].T).max() == 0.0


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 11, 16])
@pytest.mark.parametrize("general", [False, True])
def test_seeded_random_against_dense_oracle(n, general):
    rng = np.random.default_rng(100 * n + general)
    M = 7
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = 0.4 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
    h[0] = 0
    h[-1] = np.diag(0.05 * rng.standard_normal(n))
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    if general:
        tnn[1:-1] += 0.3 * (rng.standard_normal((M - 3, n, n)) + 1j * rng.standard_normal((M - 3, n, n)))
    else:
        tnn *= np.array([1.0, 0.9, 1.1 + 0.2j, 0.7, -1.0, 1.0])[:, None, None]
    tnnn = np.zeros((M - 2, n, n), dtype=complex)
    Es = np.linspace(-1.8, 1.7, 9).astype(complex)
    srcs = np.vstack([np.eye(n)[0], rng.standard_normal(n)])
    R, T, res = transmission_sweep(h, tnn, Es, sources=srcs, want_resid=True)
    oR = np.zeros_like(R[0])
    oT = np.zeros_like(T[0])
    for i, E in enumerate(Es):
        for s, src in enumerate(srcs):
            oR[i, s], oT[i, s] = wfm_oracle.kernel(h, tnn, tnnn, 1.0, E, "g_closed", src, False, False, all_debug=False)
    _check(R[0], T[0], oR, oT, res)
    # all one-hot sources in one go (the static path)
    Ra, Ta = transmission_sweep(h, tnn, Es)
    _check(Ra[0, :, 0], Ta[0, :, 0], oR[:, 0], oT[:, 0])


def test_pivoting_needed_block():
    """A block whose diagonal vanishes at the chosen energy: unpivoted elimination would divide by 0."""
    n, M = 4, 4
    E = 0.5
    h = np.zeros((M, n, n), dtype=complex)
    X = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
    h[1] = E * np.eye(n) + 0.7 * X
    h[2] = E * np.eye(n) + 0.3 * X[::-1, ::-1]
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    tnnn = np.zeros((M - 2, n, n), dtype=complex)
    R, T, res = transmission_sweep(h, tnn, np.array([E], dtype=complex), want_resid=True)
    oR, oT = rgf_oracle.transmission_closed(h, tnn, [E], np.eye(n))
    _check(R[0], T[0], oR, oT, res)
    for a in range(n):
        dR, dT = wfm_oracle.kernel(h, tnn, tnnn, 1.0, E, "g_closed", np.eye(n)[a], False, False)
        _check(R[0, 0, a], T[0, 0, a], dR, dT)


def test_ragged_batches_and_edge_sizes():
    h, tnn, tnnn = configs.cicc_blocks(1, 1.0, -0.3, 0.0, 0.0, 0.0, 3, 2)
    for nE in (1, 7, 33):          # not multiples of the 8 points a warp holds
        Es = np.linspace(-1.95, -0.2, nE).astype(complex)
        R, T = transmission_sweep(h, tnn, Es)
        oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, np.eye(8))
        _check(R[0], T[0], oR, oT)
    # smallest legal chain: M = 2 (two lead cells, no scattering region)
    h2 = np.zeros((2, 2, 2), dtype=complex)
    t2 = -np.ones((1, 1, 1)) * np.eye(2, dtype=complex)[None]
    R, T = transmission_sweep(h2, t2, np.array([-0.5 + 0j]))
    assert np.allclose(T[0, 0], np.eye(2), atol=1e-13) and np.allclose(R[0, 0], 0, atol=1e-13)
    # energies outside the band: closed source channel -> NaN, as the reference's 0/0
    R, T = transmission_sweep(h, tnn, np.array([2.5 + 0j]))
    assert np.isnan(T).all()


def test_exactly_singular_block_raises_linalgerror():
    """E - H' exactly singular: numpy's dense inverse raises LinAlgError (wfm/__init__.py:162)."""
    h = np.zeros((3, 2, 2), dtype=complex)
    tnn = -np.tile(np.eye(2, dtype=complex), (2, 1, 1))
    zero = np.zeros((1, 2, 2), dtype=complex)
    with pytest.raises(np.linalg.LinAlgError):
        transmission_sweep(h, tnn, np.array([0.0 + 0j]), "table", sigL=zero, sigR=zero)


def test_long_chain_unitarity_and_oracle():
    """cfg3 down-scaled: N=1000 disordered sites (the dense reference path cannot do this size)."""
    h, tnn, _ = configs.disordered_blocks(N=1000, n=8, W=0.05, seed=7)
    Es = np.linspace(-1.9, 1.9, 40).astype(complex)
    R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
    assert np.max(np.abs(res)) < UNIT
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Es[::8], np.eye(8))
    _check(R[0, ::8], T[0, ::8], oR, oT)


def test_green_column_and_full_against_oracle():
    rng = np.random.default_rng(5)
    for n in (2, 6, 20):
        M = 5
        A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
        h = 0.3 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
        h[0] = 0
        h[-1] = 0
        tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1)) + 0.1 * rng.standard_normal((M - 1, n, n))
        tnn[0] = -np.eye(n)
        tnn[-1] = -np.eye(n)
        Es = np.array([-1.2, 0.3], dtype=complex)
        SL, SR = rgf_oracle.closed_selfenergies(h, tnn, Es)
        dSL, dSR = leads.self_energies_batched(h, tnn, Es, "g_closed")
        assert np.allclose(dSL, SL, rtol=1e-13, atol=1e-15) and np.allclose(dSR, SR, rtol=1e-13, atol=1e-15)
        col = green.green_column0(h, tnn, Es, SL, SR)
        assert np.allclose(col, rgf_oracle.green_column0(h, tnn, Es, SL, SR), rtol=1e-10, atol=1e-13)
        full = green.green_full(h, tnn, Es, SL, SR)
        want = np.array([wfm_oracle.Green(h, tnn, np.zeros((M - 2, n, n), complex), 1.0, E, "g_closed") for E in Es])
        assert np.allclose(full, want, rtol=1e-10, atol=1e-12)


# ------------------------------------------------------------------------------- bardeen adaptor, large blocks
def test_hsysmat_golden():
    from transport_b200 import bardeen
    g = load_golden("hsysmat.npz")
    H = bardeen.Hsysmat(g["tinf"], g["tL"], g["tR"], g["Vinf"], g["VL"], g["VR"], int(g["Ninf"]), int(g["NL"]),
                        int(g["NR"]), g["HC"])
    assert H.shape == g["H"].shape and np.array_equal(H, g["H"])
    H2 = bardeen.Hsysmat(g["tinf"], g["tL"], g["tR"], g["Vinf"], g["VL"], g["VR"], 2, 3, 4, g["HC"], bound=False)
    assert np.array_equal(np.diagonal(H2[0, 0]), np.diagonal(g["VL"]))
    with pytest.raises(ValueError):
        bardeen.Hsysmat(g["tinf"], g["tL"], g["tR"], g["Vinf"], g["VL"], g["VR"], 2, 3, 4, g["HC"][:2, :2])
    with pytest.raises(TypeError):
        bardeen.Hsysmat(g["tinf"], g["tL"], g["tR"], g["Vinf"], g["VL"], g["VR"], 2.0, 3, 4, g["HC"])


@pytest.mark.parametrize("with_nnn", [False, True])
def test_ts_wfm_well_against_dense_oracle(with_nnn):
    """rb* scripts' entry into T(E) (bardeen/__init__.py:634-685), incl. next-nearest-neighbour blocks."""
    from transport_b200 import bardeen
    rng = np.random.default_rng(11)
    n, NC = 2, 5
    HC = np.zeros((NC, NC, n, n), dtype=complex)
    for j in range(NC):
        HC[j, j] = 0.3 * np.eye(n) + 0.05 * np.array([[0, 1], [1, 0]])
    for j in range(NC - 1):
        HC[j, j + 1] = -1.0 * np.eye(n)
        HC[j + 1, j] = -1.0 * np.eye(n)
    if with_nnn:
        for j in range(NC - 2):
            blk = 0.08 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
            HC[j, j + 2] = blk
            HC[j + 2, j] = np.conj(blk).T
    tL = tR = 1.0 * np.eye(n)
    VL = VR = 0.0 * np.eye(n)
    Emas = np.array([[-1.7, -1.1, -0.2], [-1.6, -0.9, 0.4]])
    Tb = bardeen.Ts_wfm_well(tL, tR, VL, VR, HC, Emas)
    assert Tb.shape == (n, 3, n)
    h, tnn, tnnn = bardeen.blocks_from_HC(1.0, 1.0, 0.0, 0.0, HC)
    for a in range(n):
        for m in range(3):
            _, oT = wfm_oracle.kernel(h, tnn, tnnn, 1.0, complex(Emas[a, m]), "g_closed", np.eye(n)[a], False, False,
                                      all_debug=False)
            assert rt_err(Tb[:, m, a], oT) < RTOL
    with pytest.raises(NotImplementedError):
        bardeen.Ts_wfm_well(tL, 0.9 * tR, VL, VR, HC, Emas)


@pytest.mark.parametrize("n,scalar", [(20, True), (24, False), (33, True), (40, False), (48, True), (56, True), (64, True),
                                      (64, False)])
def test_large_block_sweep_and_caroli(n, scalar):
    """16 < n <= 64 goes through the CTA-per-point kernel; parity vs the recursive numpy oracle."""
    rng = np.random.default_rng(n)
    M = 5
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = 0.2 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
    h[0] = np.diag(np.zeros(n))
    h[-1] = np.diag(0.02 * rng.standard_normal(n))
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    if not scalar:
        tnn[1:-1] += 0.1 * (rng.standard_normal((M - 3, n, n)) + 1j * rng.standard_normal((M - 3, n, n)))
    Es = np.array([-1.4, 0.3, 1.2], dtype=complex)
    srcs = np.eye(n)[[0, n // 2]]
    R, T, res, tot, car = transmission_sweep(h, tnn, Es, sources=srcs, want_resid=True, want_total=True, want_caroli=True)
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, srcs)
    _check(R[0], T[0], oR, oT, res)
    SL, SR = rgf_oracle.closed_selfenergies(h, tnn, Es)
    G00, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es, SL, SR)
    assert np.allclose(car[0], rgf_oracle.caroli_trace(GN0, SL, SR), rtol=1e-10)
    assert np.allclose(tot[0], T[0].sum(axis=(1, 2)), rtol=1e-12)


def test_ribbon_sancho_leads_cfg4_downscaled():
    """cfg4 shape at reduced size: ribbon width 24, Sancho-Rubio lead tables -> sweep -> Caroli trace.
    Parity: with the SAME self-energy tables the sweep must match the numpy
    restatement to 1e-10; the decimation itself is compared at eta=1e-3 to 1e-9 and at the cfg4 value
    eta=1e-8 to 1e-6 (the fixed point's condition number grows like 1/eta, so 1e-16 round-off differences
    between two correct implementations show up at ~1e-8 there).  Clean ribbon: exact channel count."""
    n, L = 24, 6
    Es = np.array([-2.3, -0.7, 0.15, 1.9], dtype=complex)
    for W in (0.0, 1.0):
        h, tnn, _ = configs.ribbon_blocks(width=n, L=L, W=W, seed=3)
        for eta, tol in ((1e-3, 1e-9), (1e-8, 1e-6)):     # two correct decimations differ by round-off x |G|^2 <= eps / eta^2
            SL, SR = leads.self_energies_batched(h, tnn, Es, ("sancho", eta, 1e-14))
            gL, _ = rgf_oracle.sancho_rubio(h[0], tnn[0], Es + 1j * eta, -1)
            gR, _ = rgf_oracle.sancho_rubio(h[-1], tnn[-1], Es + 1j * eta, 1)
            oSL = np.conj(tnn[0]).T @ gL @ tnn[0]
            oSR = tnn[-1] @ gR @ np.conj(tnn[-1]).T
            for dev_S, ora_S in ((SL, oSL), (SR, oSR)):
                assert np.abs(dev_S - ora_S).max() / np.abs(ora_S).max() < tol
            # fixed-point residual of the reference's own map at the same z
            (gdev, _, _, ok) = leads.surface_gf(h[-1], tnn[-1], Es, 1, _lib.TX_GF_SANCHO, imE=eta, conv_tol=1e-14,
                                                max_iter=200, full=True)
            assert ok.all()
            for i, E in enumerate(Es):
                resid = lambda g: (np.linalg.norm(wfm_oracle.g_ith(h[-1], tnn[-1], E, 1, eta, 1, g) - g)
                                   / np.linalg.norm(g))
                # within two digits of the fixed-point quality the float64 numpy decimation reaches at this eta
                assert resid(gdev[i]) < max(1e-12, 100 * resid(gR[i]))
        # same tables on both sides -> the sweep itself is held to 1e-10
        R, T, car = transmission_sweep(h, tnn, Es, "table", sigL=oSL, sigR=oSR, sources=np.eye(n)[:1], want_caroli=True)
        _, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es, oSL, oSR)
        want = rgf_oracle.caroli_trace(GN0, oSL, oSR)
        assert np.allclose(car[0], want, rtol=1e-10, atol=1e-12)
        # device tables end to end
        R, T, car2 = transmission_sweep(h, tnn, Es, "table", sigL=SL, sigR=SR, sources=np.eye(n)[:1], want_caroli=True)
        assert np.allclose(car2[0], want, rtol=1e-6, atol=1e-8)
        if W == 0.0:
            eps = np.linalg.eigvalsh(h[0].real)
            open_ch = np.array([np.sum(np.abs(E.real - eps) < 2.0) for E in Es])
            assert np.allclose(car[0], open_ch, atol=1e-4)      # clean ribbon: T = number of open channels


@pytest.mark.parametrize("n", [3, 24, 64, 72, 96])
def test_one_decimation_yields_both_lead_surfaces(n):
    """tx_surface_gf_pair: the left and right continuations of one periodic lead from a single Sancho-Rubio decimation
    (other surface block = bulk + h00 - surface) against the two independent decimations, the numpy oracle and the
    fixed-point maps of both sides (wfm/__init__.py:547-551); general complex hopping, not just ribbons"""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    h00 = 0.4 * (A + A.conj().T) / 2 / np.sqrt(n)
    t = -np.eye(n) + 0.2 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)
    Es = np.array([-1.3, 0.1, 0.9], dtype=complex)
    eta = 1e-3
    gL, gR, sL, sR, conv, nit, ok = leads.surface_gf_pair(h00, t, Es, imE=eta, conv_tol=1e-14, want_g=True)
    assert ok.all()
    for side, g, sig in ((-1, gL, sL), (1, gR, sR)):
        ((g1, s1), _, _, ok1) = leads.surface_gf(h00, t, Es, side, _lib.TX_GF_SANCHO, imE=eta, conv_tol=1e-14, max_iter=200,
                                                 full=True, want_sigma=True)
        assert ok1.all()
        # (two correct decimations differ by round-off amplified by |G|^2 ~ 1/eta^2: 5e-11 here)
        assert np.abs(g - g1).max() < 1e-9 * np.abs(g1).max() and np.abs(sig - s1).max() < 1e-9 * np.abs(s1).max()
        go, _ = rgf_oracle.sancho_rubio(h00, t, Es + 1j * eta, side)
        assert np.abs(g - go).max() < 1e-9 * np.abs(go).max()
        for i, E in enumerate(Es):
            step = wfm_oracle.g_ith(h00, t, E, side, eta, 1, g[i])
            assert np.linalg.norm(step - g[i]) < 1e-9 * np.linalg.norm(g[i])
    # identical lead blocks: the host sweep and self_energies_batched take the single decimation by themselves
    M = 4
    h = np.array([h00] * M)
    h[1:-1] += 0.1 * np.eye(n)
    tnn = np.array([t] * (M - 1))
    SL, SR = leads.self_energies_batched(h, tnn, Es, ("sancho", eta, 1e-14))
    assert np.array_equal(SL, sL) and np.array_equal(SR, sR)
    car = transmission_sweep(h, tnn, Es, ("sancho", eta), sources=np.eye(n)[:1], want_rt=False, want_caroli=True)
    car_t = transmission_sweep(h, tnn, Es, "table", sigL=sL, sigR=sR, sources=np.eye(n)[:1], want_rt=False, want_caroli=True)
    assert np.allclose(car, car_t, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("n", [4, 24, 64, 72, 99])
def test_hermitian_hopping_decimation_two_products(n):
    """Sancho-Rubio with a Hermitian hopping block (alpha = beta throughout: two products per iteration, detected on the
    device) against the general iteration (tx_set_option(6, 0)), the numpy oracle and the fixed-point map; both surfaces of
    such a lead coincide"""
    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    h00 = 0.4 * (A + A.conj().T) / 2 / np.sqrt(n)
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    t = -np.eye(n) + 0.2 * (B + B.conj().T) / 2 / np.sqrt(n)          # Hermitian, not scalar
    Es = np.array([-1.3, 0.1, 0.9], dtype=complex)
    eta = 1e-3
    lib = _lib.load()
    res = {}
    try:
        for mode in (1, 0):
            _lib.check(lib.tx_set_option(6, mode))
            res[mode] = leads.surface_gf_pair(h00, t, Es, imE=eta, conv_tol=1e-14, want_g=True)
            assert res[mode][-1].all()
    finally:
        _lib.check(lib.tx_set_option(6, 1))
    gL, gR, sL, sR = res[1][:4]
    assert np.abs(gL - gR).max() < 1e-9 * np.abs(gR).max()
    assert np.array_equal(res[1][5], res[0][5])                       # same number of iterations
    for a, b in zip(res[1][:4], res[0][:4]):
        assert np.abs(a - b).max() < 1e-9 * np.abs(b).max()
    go, _ = rgf_oracle.sancho_rubio(h00, t, Es + 1j * eta, 1)
    assert np.abs(gR - go).max() < 1e-9 * np.abs(go).max()
    for i, E in enumerate(Es):
        resid = lambda g: np.linalg.norm(wfm_oracle.g_ith(h00, t, E, 1, eta, 1, g) - g) / np.linalg.norm(g)
        assert resid(gR[i]) < max(1e-9, 100 * resid(go[i]))         # as good a fixed point as the float64 numpy decimation
    # single-side entry takes the same path
    ((g1, s1), _, _, ok1) = leads.surface_gf(h00, t, Es, -1, _lib.TX_GF_SANCHO, imE=eta, conv_tol=1e-14, max_iter=200,
                                             full=True, want_sigma=True)
    assert ok1.all() and np.abs(g1 - gL).max() < 1e-9 * np.abs(gL).max() and np.abs(s1 - sL).max() < 1e-9 * np.abs(sL).max()


def test_all_kernel_variants_agree():
    """Every kernel of the n=8 sweep (register-resident, plain DMMA, telescoped) gives the golden answer;
    the DMMA kernel is also exercised with padding (n=5..7), table leads and mixed Hamiltonians per warp."""
    lib = _lib.load()
    g = load_golden("cfg2_cicc.npz")
    try:
        for v in (0, 6, 13):
            _lib.check(lib.tx_set_variant(v))
            R, T, res = transmission_sweep(g["h"], g["tnn"], g["Es"], want_resid=True)
            _check(R, T, g["R"], g["T"], res)
    finally:
        _lib.check(lib.tx_set_variant(13))
    rng = np.random.default_rng(3)
    for n in (5, 6, 7, 8):
        M = 6
        P = 3
        A = rng.standard_normal((P, M, n, n)) + 1j * rng.standard_normal((P, M, n, n))
        h = 0.3 * (A + np.conj(np.swapaxes(A, 2, 3))) / 2
        h[:, 0] = 0
        h[:, -1] = 0
        tnn = -np.tile(np.eye(n, dtype=complex), (P, M - 1, 1, 1)) * np.array([1.0, 0.8, 1.1])[:, None, None, None]
        tnn[:, 0] = -np.eye(n)
        tnn[:, -1] = -np.eye(n)
        Es = np.linspace(-1.6, 1.5, 5).astype(complex)     # 5 energies: warps mix Hamiltonians
        R, T = transmission_sweep(h, tnn, Es)
        for p in range(P):
            oR, oT = rgf_oracle.transmission_closed(h[p], tnn[p], Es, np.eye(n))
            _check(R[p], T[p], oR, oT)
        SL, SR = rgf_oracle.closed_selfenergies(h[0], tnn[0], Es)
        R2, T2 = transmission_sweep(h, tnn, Es, "table", sigL=SL, sigR=SR, sources=np.eye(n)[:2])
        for p in range(P):
            oR, oT = rgf_oracle.transmission_closed(h[p], tnn[p], Es, np.eye(n)[:2])
            _check(R2[p], T2[p], oR, oT)


def test_landauer_current_and_rhat_extensions():
    """reflection operator from G_00 and the Landauer current integral."""
    from transport_b200 import current
    h, tnn, tnnn = configs.cicc_blocks(1, 1.0, -0.4, 0.0, 0.0, 0.0, 3, 2)
    Es = np.linspace(-1.95, -0.05, 400).astype(complex)
    R, T, tot = transmission_sweep(h, tnn, Es, want_total=True)
    mu_L = np.array([-1.0, -0.8, -0.5, -1.2])
    mu_R = np.array([-1.2, -1.0, -0.9, -1.2])
    for kBT in (0.0, 0.03):
        I, jE = current.landauer_current(tot[0], Es.real, mu_L, mu_R, kBT, want_density=True)
        for b in range(len(mu_L)):
            if kBT == 0.0:
                nL, nR = (Es.real <= mu_L[b]).astype(float), (Es.real <= mu_R[b]).astype(float)
            else:
                nL = 1 / (np.exp((Es.real - mu_L[b]) / kBT) + 1)
                nR = 1 / (np.exp((Es.real - mu_R[b]) / kBT) + 1)
            want_j = tot[0] * (nL - nR) / (2 * np.pi)
            assert np.allclose(jE[b], want_j, rtol=1e-13, atol=1e-16)
            assert np.isclose(I[b], np.trapezoid(want_j, Es.real), rtol=1e-12, atol=1e-15)
    assert abs(I[-1]) < 1e-15                      # no bias, no current
    # Rhat: all channels share ka_L here, so |Rhat[beta, alpha]|^2 are the reflection probabilities
    E = complex(-1.3)
    Rhat = wfm.Rhat_matrix(h, tnn, tnnn, 1.0, E)
    Rk = np.array([wfm.kernel(h, tnn, tnnn, 1.0, E, "g_closed", np.eye(8)[a], False, False)[0] for a in range(8)])
    assert np.allclose(np.abs(Rhat.T) ** 2, Rk, rtol=1e-10, atol=1e-13)


def test_cfg2_full_size_properties():
    """BASELINE config 2 at full size (1000 J x 1000 E x 8 sources) through the host C ABI: unitarity on all
    8e6 (point, source) pairs, the golden rows, T_total = sum of channels, and sweep linearity in the batch
    (a warp-aligned sub-batch reproduces the same bits)."""
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    Js, Es = configs.cfg2_grid(1000, 1000)
    R, T, res, tot = transmission_sweep(h0, tnn, Es, h1=h1, params=Js, want_resid=True, want_total=True)
    assert R.shape == (1000, 1000, 8, 8)
    assert np.isfinite(R).all() and np.isfinite(T).all()
    assert np.max(np.abs(res)) < UNIT
    assert np.allclose(tot, T.sum(axis=(2, 3)), rtol=1e-12, atol=1e-15)
    assert T.min() >= 0 and R.min() >= 0
    g = load_golden("cfg2_cicc.npz")
    for k, jidx in enumerate([0, 250, 500, 750, 999]):
        _check(R[jidx, ::40], T[jidx, ::40], g["R"][k], g["T"][k])
    # a sub-batch aligned to the eight energies of a warp reproduces the same bits; an unaligned, short one
    # (served by the point-per-4-lanes kernel) the same numbers
    R2, T2 = transmission_sweep(h0, tnn, Es[96:136], h1=h1, params=Js[400:403])
    assert np.array_equal(T2, T[400:403, 96:136]) and np.array_equal(R2, R[400:403, 96:136])
    R3, T3 = transmission_sweep(h0, tnn, Es[100:133], h1=h1, params=Js[400:403])
    assert rt_err(T3, T[400:403, 100:133]) < 1e-11 and rt_err(R3, R[400:403, 100:133]) < 1e-11


def test_cfg1_full_size_and_cfg3_long_chain_properties():
    """config 1 at full size against the golden stride; config 3 at N=10^4 sites on a strided energy grid:
    finite, non-negative, unitarity within the conditioning of the problem at that length (a 1-ulp change of
    E moves the oracle's own T by ~1e-9, see profiles/README.md), exact agreement of a re-run (determinism)."""
    g = load_golden("cfg1_barrier.npz")
    h, tnn, _ = configs.barrier_blocks(0.4, 11)
    Es = configs.barrier_energies(10_000)
    R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
    assert np.max(np.abs(res)) < UNIT
    _check(R[0, ::50], T[0, ::50], g["R"], g["T"])
    below = (Es.real + 2.0) < 0.4                      # under the barrier top: pure tunnelling, monotone in E
    assert np.all(np.diff(T[0, below, 0, 0]) > 0)
    h, tnn, _ = configs.disordered_blocks(N=10_000, n=8, W=0.05, seed=1234)
    Es = np.linspace(-1.9, 1.9, 100_000).astype(complex)[::500]
    R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
    assert np.isfinite(R).all() and np.isfinite(T).all() and T.min() >= 0
    assert np.max(np.abs(res)) < 1e-7
    R2, T2 = transmission_sweep(h, tnn, Es)
    assert np.array_equal(T, T2) and np.array_equal(R, R2)


def test_diagonal_tail_fast_path_matches_block_path():
    """Sites right of the last channel-coupling block are swept per channel (option 0); forcing the block path
    on every site must give the same numbers, for tails of every length including the whole chain."""
    lib = _lib.load()
    Es = np.linspace(-1.9, -0.1, 21).astype(complex)
    cases = []
    for NB in (1, 3, 40):
        cases.append(configs.cicc_blocks(1, 1.0, -0.7, 0.05, 0.3, 0.0, NB, 2)[:2])      # barrier tail at V_B
    h, tnn, _ = configs.barrier_blocks(0.4, 9, n=8)                                       # fully diagonal chain
    cases.append((h, tnn))
    h, tnn, _ = configs.disordered_blocks(N=12, n=7, W=0.2, seed=5)                       # no diagonal tail, padded
    cases.append((h, tnn))
    # scalar runs: impurities far from the left lead (run of 22 clean sites incl. the lead cell), and a
    # 30-site evanescent barrier (V = 1.2 > band top at these energies) between two channel-coupling blocks
    cases.append(configs.cicc_blocks(1, 1.0, -0.7, 0.05, 0.3, 0.0, 25, 2, offset=20)[:2])
    rng = np.random.default_rng(17)
    M = 40
    h = np.zeros((M, 8, 8), dtype=complex)
    for j in (3, 36):
        A = rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8))
        h[j] = 0.3 * (A + A.conj().T) / 2
    h[5:35] = 1.2 * np.eye(8)
    h[1] = 0.1 * np.eye(8)
    tnn = -np.tile(np.eye(8, dtype=complex), (M - 1, 1, 1)) * (1.0 + 0.05 * rng.standard_normal(M - 1))[:, None, None]
    tnn[0] = -np.eye(8)
    tnn[-1] = -np.eye(8)
    cases.append((h, tnn))
    for h, tnn in cases:
        n = h.shape[-1]
        try:
            _lib.check(lib.tx_set_option(0, 0))
            Rb, Tb = transmission_sweep(h, tnn, Es)
        finally:
            _lib.check(lib.tx_set_option(0, 1))
        Rf, Tf, res = transmission_sweep(h, tnn, Es, want_resid=True)
        assert rt_err(Tf, Tb) < 1e-11 and rt_err(Rf, Rb) < 1e-11      # two float64 evaluation orders of the same chain
        assert np.max(np.abs(res)) < UNIT
        oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, np.eye(n))
        _check(Rf[0], Tf[0], oR, oT)


# ------------------------------------------------------------------------------- telescoped sweep (rgf_n8t.cuh)
def _telescoped_cases():
    rng = np.random.default_rng(41)
    cases = {}
    cases["disorder_n8"] = configs.disordered_blocks(N=40, n=8, W=0.3, seed=3)[:2]
    cases["disorder_n6_padded"] = configs.disordered_blocks(N=21, n=6, W=0.2, seed=4)[:2]
    # evanescent barrier of mixed heights between two channel-coupling regions: the growth bound has to close chunks
    h, tnn, _ = configs.disordered_blocks(N=50, n=8, W=0.2, seed=5)
    for j in range(15, 35):
        h[j] += np.diag(rng.uniform(0.0, 5.0, 8))
    cases["mixed_barrier"] = (h, tnn)
    # complex, non-unit scalar hoppings (|t| far from 1: the recurrence's scaling factors matter)
    h, tnn, _ = configs.disordered_blocks(N=30, n=7, W=0.5, tl=2.7, seed=6)
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, len(tnn)))
    ph[0] = ph[-1] = 1.0
    tnn = tnn * ph[:, None, None] * rng.uniform(0.5, 1.5, len(tnn))[:, None, None]
    tnn[0] = tnn[-1] = -2.7 * np.eye(7)
    cases["complex_hopping_n7"] = (h, tnn)
    return cases


@pytest.mark.parametrize("kmax,gexp", [(1, 4), (2, 4), (3, 0), (8, 4), (32, 4), (32, 2), (64, 5)])
def test_telescoped_sweep_chunk_lengths(kmax, gexp):
    """Every chunk length of the telescoped sweep, with growth bounds around the default 2^4, reproduces the plain
    recursion (kmax = 1 IS the plain recursion: one inverse per site) and the numpy oracle -- including the
    1e-24 tunnelling entries of the mixed barrier, which are measured relative to themselves.  (Without a growth
    bound the component-wise accuracy of such entries degrades as growth^2; the numpy restatement shows the same.)"""
    lib = _lib.load()
    try:
        _lib.check(lib.tx_set_option(1, kmax))
        _lib.check(lib.tx_set_option(2, gexp))
        for name, (h, tnn) in _telescoped_cases().items():
            n = h.shape[-1]
            scale = float(np.abs(tnn[0][0, 0]))
            Es = (np.linspace(-1.9, 1.9, 24) * scale).astype(complex)
            R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
            oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, np.eye(n))
            _check(R[0], T[0], oR, oT, res)
    finally:
        _lib.check(lib.tx_set_option(1, 32))
        _lib.check(lib.tx_set_option(2, 4))


def test_telescoped_sweep_edge_cases():
    """Energy axes that are / are not multiples of the eight points of a warp, several Hamiltonians, complex
    energies (the general fragment path), table-mode leads, amplitude-vector sources, a decoupled site (t = 0)."""
    rng = np.random.default_rng(43)
    h, tnn, _ = configs.disordered_blocks(N=18, n=8, W=0.3, seed=8)
    for nE in (8, 16, 64, 65, 100):
        Es = np.linspace(-1.9, 1.9, nE).astype(complex)
        R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
        oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, np.eye(8))
        _check(R[0], T[0], oR, oT, res)
    # three Hamiltonians x 72 energies, amplitude-vector sources
    P = 3
    hs = np.stack([configs.disordered_blocks(N=18, n=8, W=0.1 * (p + 1), seed=20 + p)[0] for p in range(P)])
    Es = np.linspace(-1.8, 1.8, 72).astype(complex)
    srcs = np.vstack([np.eye(8)[3], rng.standard_normal(8)])
    R, T, res = transmission_sweep(hs, tnn, Es, sources=srcs, want_resid=True)
    for p in range(P):
        oR, oT = rgf_oracle.transmission_closed(hs[p], tnn, Es, srcs)
        _check(R[p], T[p], oR, oT, res[p])
    # complex energies: retarded GF at finite eta (R + T < 1: absorption, so no unitarity check)
    Ec = np.linspace(-1.5, 1.5, 16) + 0.01j
    R, T = transmission_sweep(h, tnn, Ec)
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Ec, np.eye(8))
    _check(R[0], T[0], oR, oT)
    # table-mode leads (n = 8 and padded n = 6) with the closed-form self-energies as tables
    for n in (8, 6):
        hh, tt, _ = configs.disordered_blocks(N=14, n=n, W=0.3, seed=9)
        Es = np.linspace(-1.7, 1.7, 24).astype(complex)
        SL, SR = rgf_oracle.closed_selfenergies(hh, tt, Es)
        R, T, res = transmission_sweep(hh, tt, Es, "table", sigL=SL, sigR=SR, want_resid=True)
        oR, oT = rgf_oracle.transmission_closed(hh, tt, Es, np.eye(n))
        _check(R[0], T[0], oR, oT, res)
    # a decoupled site: t_5 = 0 cuts the chain, T = 0 and R = 1 exactly as the dense inverse gives
    t0 = tnn.copy()
    t0[5] = 0
    Es = np.linspace(-1.5, 1.5, 16).astype(complex)
    R, T, res = transmission_sweep(h, t0, Es, want_resid=True)
    oR, oT = rgf_oracle.transmission_closed(h, t0, Es, np.eye(8))
    assert np.all(T == 0)
    _check(R[0], T[0], oR, oT, res)


def test_separable_multichannel_leads():
    """TX_GF_SEPARABLE (scalar hopping block): g = U diag(g_closed(eps_m)) U^dag with the device Jacobi
    eigensolver, against the same expression built with numpy's eigh and the oracle's g_closed; against
    Sancho-Rubio at small eta away from the band edges; refusal of a non-scalar hopping block."""
    rng = np.random.default_rng(11)
    for n, tval in ((5, -1.0), (24, -1.3), (64, -1.0)):
        if n == 64:
            h00 = configs.ribbon_blocks(width=64, L=1)[0][0]
        else:
            X = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
            h00 = 0.5 * (X + X.conj().T) / 2
        h01 = tval * np.eye(n, dtype=complex)
        Es = np.linspace(-3.4, 3.4, 21).astype(complex) * abs(tval)
        eps, U = np.linalg.eigh(h00)
        for side in (-1, 1):
            ((g, sig),) = leads.surface_gf(h00, h01, Es, side, _lib.TX_GF_SEPARABLE, want_sigma=True)
            for i, E in enumerate(Es):
                gm = np.array([wfm_oracle.g_closed(np.array([[e_m]], dtype=complex), np.array([[tval]], dtype=complex), E, side)[0, 0]
                               for e_m in eps])
                want = (U * gm) @ U.conj().T
                scale = np.abs(want).max()
                assert np.abs(g[i] - want).max() < 1e-11 * scale, (n, i, np.abs(g[i] - want).max() / scale)
                assert np.abs(sig[i] - abs(tval) ** 2 * want).max() < 1e-11 * scale * abs(tval) ** 2
    # agreement with the decimation as eta -> 0 (g_closed is the eta = 0 limit; O(sqrt(eta)) near band edges)
    h00 = configs.ribbon_blocks(width=16, L=1)[0][0]
    h01 = -np.eye(16, dtype=complex)
    Es = np.array([-2.9, -1.3, 0.45, 2.2], dtype=complex)
    (g_sep,) = leads.surface_gf(h00, h01, Es, 1, _lib.TX_GF_SEPARABLE)
    (g_sr, conv, nit, ok) = leads.surface_gf(h00, h01, Es, 1, _lib.TX_GF_SANCHO, imE=1e-9, conv_tol=1e-15, max_iter=200, full=True)
    assert ok.all()
    assert np.abs(g_sep - g_sr).max() < 1e-6
    with pytest.raises(_lib.TransportError):
        bad = h01.copy()
        bad[0, 1] = 0.1
        leads.surface_gf(h00, bad, Es, 1, _lib.TX_GF_SEPARABLE)


@pytest.fixture(params=["staged", "in_place"])
def inverse_mode(request, monkeypatch):
    """blocks staged from global memory (n >= 56 in the kernels) or resident in shared memory (n <= 48)"""
    if request.param == "in_place":
        monkeypatch.setenv("TX_INVERT_INPLACE", "1")
    return request.param


@pytest.mark.parametrize("n", [8, 16, 24, 40, 48, 56, 64, 13])
def test_staged_inverse_blocked_gauss_jordan(n, inverse_mode):
    """the CTA inverse of the large-block paths (blocked tensor-core Gauss-Jordan with implicit partial pivoting for
    n % 8 == 0, n >= 16) against numpy, including blocks that cannot be inverted without pivoting"""
    from transport_b200._lib import as_c128, check, ptr
    rng = np.random.default_rng(n)
    count = 6
    A = rng.normal(size=(count, n, n)) + 1j * rng.normal(size=(count, n, n))
    A[1, :n // 2, :n // 2] = 0.0                                  # zero leading block: pivoting required
    A[2] = np.eye(n)[::-1] * (1 + 0.5j) + 1e-3 * A[2]             # anti-diagonal dominant
    A[3] = np.diag(np.exp(rng.uniform(-8, 8, size=n))) @ A[3]     # badly row-scaled
    A[4] = 3.0 * np.eye(n) - 0.5 * (A[4] + A[4].conj().T) / np.sqrt(n) + 1e-8j * np.eye(n)   # E - h - Sigma like
    out = np.empty_like(A)
    check(_lib.load().tx_debug_invert_staged(ptr(as_c128(A)), ptr(out), count, n))
    for b in range(count):
        ref = np.linalg.inv(A[b])
        scale = np.abs(ref).max()
        assert np.max(np.abs(out[b] - ref)) < 1e-11 * scale * np.linalg.cond(A[b]) ** 0.5, (n, b)
        assert np.max(np.abs(out[b] @ A[b] - np.eye(n))) < 1e-9, (n, b)
    S = A[0].copy()
    S[:, 3] = 0.0                                                 # exactly singular
    with pytest.raises(_lib.TransportError) as ei:
        check(_lib.load().tx_debug_invert_staged(ptr(as_c128(S[None])), ptr(out[:1]), 1, n))
    assert ei.value.code == _lib.TX_ERR_SINGULAR


@pytest.mark.parametrize("name", ["nnn_n1_M8.npz", "nnn_n2_M6.npz", "nnn_n2_M7.npz", "nnn_n3_M5.npz"])
def test_next_nearest_neighbour_blocks_through_reference_signatures(name):
    """nonzero tnnn through wfm.kernel / wfm.Green / the psi output (reference: block-pentadiagonal dense assembly,
    wfm/__init__.py:209-213; here: sites paired into super-blocks), against the live reference's output."""
    g = load_golden(name)
    h, tnn, tnnn = g["h"], g["tnn"], g["tnnn"]
    for i, E in enumerate(g["Es"]):
        for s, src in enumerate(g["sources"]):
            Rs, Ts = wfm.kernel(h, tnn, tnnn, 1.0, E, "g_closed", src, False, False, all_debug=True)
            _check(Rs, Ts, g["R"][i, s], g["T"][i, s])
        psi = wfm.kernel(h, tnn, tnnn, 1.0, E, "g_closed", g["sources"][0], True, False, all_debug=False)
        assert np.allclose(psi, g["psi"][i], rtol=1e-10, atol=1e-13)
        G = wfm.Green(h, tnn, tnnn, 1.0, E, "g_closed")
        assert G.shape == g["G"][i].shape and np.allclose(G, g["G"][i], rtol=1e-10, atol=1e-13)
    Rh = wfm.Rhat_matrix(h, tnn, tnnn, 1.0, g["Es"][1])
    n = h.shape[-1]
    ka = np.arccos((g["Es"][1] - np.diagonal(h[0])) / (-2.0))
    assert np.allclose(Rh, 2j * g["G"][1][0, 0] * np.sin(ka)[None, :] - np.eye(n), rtol=1e-10, atol=1e-13)


def test_ts_wfm_well_channel_dependent_lead_potentials():
    """VL = VR = diag(0, Delta) as rbmenez_sup_inel.py:135 passes them: every channel keeps its own lead on-site energy
    (bardeen/__init__.py:651, :667 multiply the matrices elementwise with the identity).  alpha = 0 against the live
    reference's wfm.kernel on the same blocks, alpha = 1 (which the reference's kernel refuses, :67-70) against the
    recursive numpy oracle."""
    from transport_b200 import bardeen
    g = load_golden("well_channel_potentials.npz")
    HC, tL, VLR, Emas = g["HC"], g["tL"], g["VLR"], g["Emas"]
    Tb = bardeen.Ts_wfm_well(tL, tL, VLR, VLR, HC, Emas)
    assert Tb.shape == (2, 4, 2)
    assert rt_err(Tb[:, :, 0].T, g["T_alpha0"]) < RTOL
    h, tnn, tnnn = bardeen.blocks_from_HC(tL, tL, VLR, VLR, HC)
    assert h[0][1, 1] == VLR[1, 1] and h[-1][1, 1] == VLR[1, 1] and h[0][0, 1] == 0
    _, oT = rgf_oracle.transmission_closed(h, tnn, Emas[1].astype(complex), np.eye(2)[[1]])
    assert rt_err(Tb[:, :, 1].T, oT[:, 0]) < RTOL
    # the old behaviour (lead potential of channel 0 used for every channel) is measurably different
    h_wrong, _, _ = bardeen.blocks_from_HC(1.0, 1.0, 0.0, 0.0, HC)
    _, oT_wrong = rgf_oracle.transmission_closed(h_wrong, tnn, Emas[1].astype(complex), np.eye(2)[[1]])
    assert np.max(np.abs(oT_wrong - oT)) > 1e-3


def test_landauer_current_pinned_to_reference_fcdmft():
    """SURVEY §8f row 4, pinned: the reference's matrix-Gamma Landauer current (fcdmft.landauer, fcdmft/__init__.py:224-292,
    golden produced by the unmodified function) against the device path — Caroli trace of the same system as a wfm chain
    (sweep epilogue, matrix Gamma) -> tx_landauer_current (Fermi window, trapezoid rule)."""
    from transport_b200 import current
    g = load_golden("landauer_fcdmft.npz")
    E = g["energies"]
    car = transmission_sweep(g["h"], g["tnn"], E.astype(complex), sources=np.eye(2)[:1], want_rt=False, want_caroli=True)
    for tag, kBT in (("kT0", 0.0), ("kT005", 0.05)):
        want_j = np.real(g["jE_" + tag])
        I, jE = current.landauer_current(car[0], E, float(g["muL"]), float(g["muR"]), kBT, want_density=True)
        assert np.allclose(jE[0], want_j, rtol=1e-11, atol=1e-15)
        assert np.isclose(I[0], float(g["I_" + tag]), rtol=1e-11)
    # the Fermi-window / trapezoid kernel alone, on the reference's own transmission function
    T_ref = np.real(g["jE_kT0"]) * 2 * np.pi        # inside the bias window n_L - n_R = 1 at kBT = 0
    window = (E <= float(g["muL"])) & (E > float(g["muR"]))
    assert np.allclose(car[0][window], T_ref[window], rtol=1e-11)


def test_cfg3_full_length_against_extended_precision():
    """BASELINE config 3 at full length (n = 8, M = 10 002) against the extended-precision truth of
    tests/golden/hp_cfg3_M10002.npz (mpmath at 50 digits, oracle/make_golden_hp.py) on the energies with the worst device
    unitarity residual plus a uniform stride.  The band-edge energies of this chain are localised (T down to 1e-18) and
    ill-conditioned: kappa(E) = |dT/T| / |dE/E| reaches 7e9, so the 1e-10 gate becomes 1e-10 + C kappa(E) eps per energy.
    Measured: device <= 49 kappa eps (median 0.4), the float64 numpy recursion up to 470 kappa eps (it inverts the nearly
    singular blocks the kernel's conditioning guard steps around)."""
    g = load_golden("hp_cfg3_M10002.npz")
    h, tnn, _ = configs.disordered_blocks(N=10_000, n=8, W=0.05, seed=1234)
    # the whole 1e5-energy axis in one sweep, as BASELINE names it (a point's chunk boundaries depend on its warp mates)
    Es = np.linspace(-1.9, 1.9, 100_000).astype(complex)
    assert np.array_equal(Es[g["idx"]], g["Es"].astype(complex))
    R, T, res = transmission_sweep(h, tnn, Es, want_resid=True)
    assert np.max(np.abs(res)) < 5e-9                      # all 1e5 energies (measured 1.2e-9 at an energy with kappa ~ 1e7)
    R, T, res = R[:, g["idx"]], T[:, g["idx"]], res[:, g["idx"]]

    def rel(a, b):
        den = np.maximum(np.abs(b), 1e-13 * np.abs(b).max(-1, keepdims=True))
        return (np.abs(a - b) / den).max(axis=(1, 2))
    dev = np.maximum(rel(T[0], g["T_hp"]), rel(R[0], g["R_hp"]))
    bound = 1e-10 + 100.0 * g["kappa"] * 2.2e-16
    assert np.all(dev <= bound), (float(np.max(dev / bound)), int(np.argmax(dev / bound)))
    # pointwise both float64 implementations scatter inside C kappa eps; the worst case of the device stays an order of
    # magnitude under the worst case of the plain float64 recursion (which inverts through the poles)
    keps = g["kappa"] * 2.2e-16
    assert np.max(dev / (1e-10 + keps)) < 0.5 * np.max(g["f64_err"] / (1e-10 + keps))
    assert np.median(dev) < 1e-10
    # unitarity: bounded by the same conditioning
    assert np.all(np.abs(res[0]).max(axis=-1) <= 1e-10 + 100.0 * g["kappa"] * 2.2e-16)


def test_cfg4_sancho_leads_against_extended_precision():
    """Sancho-Rubio self-energy of the config-4 ribbon lead (n = 64) at eta = 1e-8 against the long-double decimation of
    tests/golden/hp_cfg4_leads.npz (oracle/make_golden_hp.py).  At eta = 1e-8 the surface Green's function itself is
    conditioned like 1/eta: the float64 numpy decimation is off by 8e-9 = 0.8 eps/eta, the device by the same amount."""
    g = load_golden("hp_cfg4_leads.npz")
    h, tnn, _ = configs.ribbon_blocks(width=64, L=32, W=1.0, seed=1234)
    SL, _ = leads.self_energies_batched(h, tnn, g["Es"].astype(complex), ("sancho", float(g["eta"])))
    den = np.abs(g["sigL_hp"]).max(axis=(1, 2), keepdims=True)
    err = (np.abs(SL - g["sigL_hp"]) / den).max(axis=(1, 2))
    assert np.max(err) < 4 * 2.2e-16 / float(g["eta"])              # a few eps / eta
    assert np.max(err) < 3 * np.max(g["f64_err"]) + 1e-10           # the level of the float64 numpy decimation
