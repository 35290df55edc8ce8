"""CPU tests: the numpy oracle (oracle/) against the golden vectors produced by the unmodified
reference (oracle/make_golden.py), and the RGF restatement against the dense oracle."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden, rt_err
from oracle import rgf_oracle, wfm_oracle
from transport_b200 import configs

TOL = 1e-12


def _sweep(h, tnn, tnnn, Es, converger, sources, fn=wfm_oracle.kernel):
    n = h.shape[-1]
    R = np.zeros((len(Es), len(sources), n))
    T = np.zeros_like(R)
    for i, E in enumerate(Es):
        for s, src in enumerate(sources):
            R[i, s], T[i, s] = fn(h, tnn, tnnn, 1.0, E, converger, np.asarray(src, dtype=float), False, False,
                                  all_debug=False)
    return R, T


def test_barrier_survey_spot_values():
    g = load_golden("barrier_run_wfm.npz")
    # SURVEY.md §8c spot values measured from the live reference
    assert g["T"][0, 0, 0] == pytest.approx(2.436987210571056e-06, rel=1e-13)
    assert g["T"][100, 0, 0] == pytest.approx(3.629186977233515e-04, rel=1e-13)
    assert g["T"][198, 0, 0] == pytest.approx(0.9956549695595422, rel=1e-13)
    R, T = _sweep(g["h"], g["tnn"], g["tnnn"], g["Es"], "g_closed", [np.array([1.0])])
    assert rt_err(T, g["T"]) < TOL and rt_err(R, g["R"]) < TOL


def test_cfg1_barrier():
    g = load_golden("cfg1_barrier.npz")
    h, tnn, tnnn = configs.barrier_blocks(0.4, 11)
    assert np.array_equal(h, g["h"]) and np.array_equal(tnn, g["tnn"])
    assert np.array_equal(configs.barrier_energies(10_000)[::50], g["Es"])
    R, T = _sweep(h, tnn, tnnn, g["Es"], "g_closed", [np.array([1.0])])
    assert rt_err(T, g["T"]) < TOL and rt_err(R, g["R"]) < TOL


def test_cfg2_cicc_dense_and_loops():
    g = load_golden("cfg2_cicc.npz")
    srcs = np.eye(8)
    for k, J in enumerate(g["Js"]):
        h, tnn, tnnn = configs.cicc_blocks(1, 1.0, J, 0.0, 0.0, 0.0, 3, 2)
        assert np.array_equal(h, g["h"][k])
        Es = g["Es"][::6]
        R, T = _sweep(h, tnn, tnnn, Es, "g_closed", srcs)
        assert rt_err(T, g["T"][k][::6]) < TOL and rt_err(R, g["R"][k][::6]) < TOL
    # the loop-for-loop port (CPU baseline) gives the same numbers
    R, T = _sweep(h, tnn, tnnn, g["Es"][:2], "g_closed", srcs[:2], fn=wfm_oracle.kernel_loops)
    assert rt_err(T, g["T"][-1][:2, :2]) < TOL and rt_err(R, g["R"][-1][:2, :2]) < TOL


def test_cicc_spot_value_from_survey():
    g = load_golden("cicc_spot.npz")
    want_T = [0, 0.447045935133, 0.127810093027, 0, 0.119217541793, 0, 0, 0]
    assert np.allclose(g["T"], want_T, atol=1e-11)
    R, T = wfm_oracle.kernel(g["h"], g["tnn"], g["tnnn"], 1.0, complex(g["E"]), "g_closed", np.eye(8)[1], False, False)
    assert rt_err(T, g["T"]) < TOL and rt_err(R, g["R"]) < TOL


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                                        if os.path.basename(p).startswith(("general_hop", "cfg3_", "cicc_spin",
                                                                           "nondiag", "kondo_n2.npz"))))
def test_closed_lead_goldens_dense_and_rgf(name):
    g = load_golden(name)
    srcs = g["sources"] if "sources" in g.files else np.eye(g["h"].shape[-1])[:g["R"].shape[1]]
    if name.startswith("cicc_spin1_"):
        srcs = [np.eye(18)[0], np.eye(18)[7]]
    R, T = _sweep(g["h"], g["tnn"], g["tnnn"], g["Es"], "g_closed", srcs)
    assert rt_err(T, g["T"]) < TOL and rt_err(R, g["R"]) < TOL
    # recursive restatement against the same reference output
    R2, T2 = rgf_oracle.transmission_closed(g["h"], g["tnn"], g["Es"], srcs)
    assert rt_err(T2, g["T"]) < 1e-11 and rt_err(R2, g["R"]) < 1e-11


@pytest.mark.parametrize("n", [50, 72, 98, 128])
def test_rgf_oracle_against_reference_goldens_of_large_spins(n):
    """two spin-2 ... spin-7/2 impurities (run_wfm_gate.get_hblocks, n_loc_dof = 50 ... 128): the recursive oracle against
    the live reference's R/T; the blocks are rebuilt by configs.cicc_blocks (held bit-equal to the reference's by the
    generating script, oracle/make_golden_r2.py cicc_large_spins)"""
    from transport_b200 import configs
    g = load_golden("cicc_builder_n%d.npz" % n)
    a = g["builder_args"]
    h, tnn, tnnn = configs.cicc_blocks(int(a[0]), a[1], a[2], a[3], a[4], a[5], int(a[6]), int(g["the_dist"]))
    assert h.shape[-1] == n and not np.any(tnnn)
    R, T = rgf_oracle.transmission_closed(h, tnn, g["Es"], g["sources"])
    assert rt_err(T, g["T"]) < 1e-10 and rt_err(R, g["R"]) < 1e-10


def test_nondiag_lead_selfenergies():
    g = load_golden("nondiag_lead_hop_n4.npz")
    SL, SR = rgf_oracle.closed_selfenergies(g["h"], g["tnn"], g["Es"])
    assert np.allclose(SL, g["SigL"], rtol=1e-13, atol=1e-15)
    assert np.allclose(SR, g["SigR"], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", ["ricemele_kondo.npz", "ricemele_kondo_vw.npz"])
def test_ricemele(name):
    g = load_golden(name)
    R, T = _sweep(g["h"], g["tnn"], g["tnnn"], g["Es"], "g_RiceMele", g["sources"])
    assert rt_err(T, g["T"]) < TOL and rt_err(R, g["R"]) < TOL


def test_iterative_converger_kernel():
    g = load_golden("kondo_n2_iterative.npz")
    conv = (float(g["imE"]), float(g["conv_tol"]))
    R, T = _sweep(g["h"], g["tnn"], g["tnnn"], g["Es"], conv, [np.eye(2)[0]])
    assert rt_err(T, g["T"]) < 1e-11 and rt_err(R, g["R"]) < 1e-11


def test_surface_gf_closed_forms():
    g = load_golden("g_closed.npz")
    for i, E in enumerate(g["Es"]):
        assert np.allclose(wfm_oracle.g_closed(g["diag"], g["off"], E, 1), g["g"][i], rtol=1e-14, atol=0)
        assert np.allclose(wfm_oracle.g_closed(g["diag"], g["off2"], E, -1), g["g2"][i], rtol=1e-14, atol=0)
    for i, E in enumerate(g["Ec"]):
        assert np.allclose(wfm_oracle.g_closed(g["diag"], g["off"], E, 1), g["g3"][i], rtol=1e-14, atol=0)
    g = load_golden("g_ricemele.npz")
    for i, E in enumerate(g["Es"]):
        assert np.allclose(wfm_oracle.g_RiceMele(g["diag"], g["off"], E, -1), g["gL"][i], rtol=1e-14, atol=0)
        assert np.allclose(wfm_oracle.g_RiceMele(g["diag"], g["off"], E, 1), g["gR"][i], rtol=1e-14, atol=0)


def test_g_ith_and_g_iter():
    g = load_golden("g_iter.npz")
    imE, tol = float(g["imE"]), float(g["conv_tol"])
    for i, E in enumerate(g["Es"]):
        for si, side in enumerate((-1, 1)):
            cur = np.zeros_like(g["h00"])
            for it in range(12):
                cur = wfm_oracle.g_ith(g["h00"], g["h01"], E, side, imE, it, cur)
                assert np.allclose(cur, g["chain"][i, si, it], rtol=1e-13, atol=1e-15)
            out, conv, nit = wfm_oracle.g_iter(g["h00"], g["h01"], E, side, imE, tol, full=True)
            assert nit == g["n_iter"][i, si]
            assert np.allclose(out, g["g"][i, si], rtol=1e-12, atol=1e-15)
    with pytest.raises(NotImplementedError):
        wfm_oracle.g_iter(np.eye(3), np.eye(3), 0.0, 1, 1e-2, 1e-3)
    # decimation reaches the fixed point of the reference's own map
    zs = g["Es"] + 1j * imE
    for side in (-1, 1):
        gs, _ = rgf_oracle.sancho_rubio(g["h00"], g["h01"], zs, side)
        for i, z in enumerate(zs):
            nxt = wfm_oracle.g_ith(g["h00"], g["h01"], z.real, side, imE, 1, gs[i])
            assert np.linalg.norm(nxt - gs[i]) / np.linalg.norm(gs[i]) < 1e-12


def test_hmat_hprime_green_psi():
    g = load_golden("green_n3.npz")
    h, tnn, tnnn, E = g["h"], g["tnn"], g["tnnn"], complex(g["E"])
    assert np.array_equal(wfm_oracle.Hmat(h, tnn, tnnn), g["H"])
    assert np.array_equal(wfm_oracle.Hmat_loops(h, tnn, tnnn), g["H"])
    assert np.allclose(wfm_oracle.Hprime(h, tnn, tnnn, 1.0, E, "g_closed"), g["Hp"], rtol=1e-14, atol=0)
    G = wfm_oracle.Green(h, tnn, tnnn, 1.0, E, "g_closed")
    assert np.allclose(G, g["G"], rtol=1e-12, atol=1e-14)
    assert np.array_equal(wfm_oracle.mat_2d_to_4d_loops(g["H"], 3), wfm_oracle.mat_2d_to_4d(g["H"], 3))
    assert np.array_equal(wfm_oracle.mat_4d_to_2d(wfm_oracle.mat_2d_to_4d(g["H"], 3)), g["H"])
    tz = np.zeros_like(tnnn)
    psi = wfm_oracle.kernel(h, tnn, tz, 1.0, E, "g_closed", g["src"], True, False, all_debug=False)
    assert np.allclose(psi, g["psi"], rtol=1e-12, atol=1e-14)
    R, T = wfm_oracle.kernel(h, tnn, tnnn, 1.0, E, "g_closed", g["src"], False, False, all_debug=False)
    assert rt_err(R, g["R_tnnn"]) < TOL and rt_err(T, g["T_tnnn"]) < TOL
    # recursive full Green's function / first column against the reference's dense inverse
    SL, SR = rgf_oracle.closed_selfenergies(h, tnn, [E])
    Gf = rgf_oracle.green_full(h, tnn, [E], SL, SR)[0]
    assert np.allclose(Gf, g["G_nn_only"], rtol=1e-11, atol=1e-13)
    col = rgf_oracle.green_column0(h, tnn, [E], SL, SR)[0]
    assert np.allclose(col, g["G_nn_only"][:, 0], rtol=1e-11, atol=1e-13)
    with pytest.raises(ValueError):
        wfm_oracle.Hmat(h, tnn[:-1], tnnn)
    with pytest.raises(TypeError):
        wfm_oracle.kernel(list(h), tnn, tnnn, 1.0, E, "g_closed", g["src"], False, False)


def test_exact_lattice_identities():
    """Builder-added KATs: clean chain T=1; delta barrier T = 1/(1+(V/(2 t sin ka))^2)."""
    Es = np.linspace(-1.9, 1.9, 21).astype(complex)
    h, tnn, tnnn = configs.barrier_blocks(0.0, 5)
    R, T = rgf_oracle.transmission_closed(h, tnn, Es, [[1.0]])
    assert np.max(np.abs(T - 1)) < 1e-13 and np.max(np.abs(R)) < 1e-13
    V = 0.7
    h, tnn, tnnn = configs.barrier_blocks(V, 1)
    R, T = rgf_oracle.transmission_closed(h, tnn, Es, [[1.0]])
    ka = np.arccos(np.real(Es) / -2.0)
    want = 1.0 / (1.0 + (V / (2 * np.sin(ka))) ** 2)
    assert np.max(np.abs(T[:, 0, 0] - want)) < 1e-13
    assert np.max(np.abs(R + T - 1)) < 1e-13


def test_caroli_trace_equals_channel_sum():
    g = load_golden("cfg2_cicc.npz")
    h, tnn = g["h"][2], g["tnn"]
    Es = g["Es"][::5]
    SL, SR = rgf_oracle.closed_selfenergies(h, tnn, Es)
    G00, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es, SL, SR)
    _, T = rgf_oracle.rt_from_blocks(G00, GN0, SL, SR, np.eye(8))
    assert np.allclose(T.sum(axis=(1, 2)), rgf_oracle.caroli_trace(GN0, SL, SR), rtol=1e-12)


@pytest.mark.parametrize("kmax,gmax", [(1, 64.0), (4, 64.0), (8, 64.0), (32, 32.0), (64, 4.0)])
def test_telescoped_restatement_matches_plain_recursion(kmax, gmax):
    """`rgf_blocks_telescoped` (the restatement of the sm_100a kernel's chunked recurrence) against the plain
    recursion, itself pinned to the dense reference path above: same blocks to round-off for every chunk length."""
    cases = [configs.cicc_blocks(1, 1.0, -0.6, 0.0, 0.0, 0.0, 3, 2)[:2],
             configs.disordered_blocks(N=60, n=8, W=0.3, seed=2)[:2],
             configs.disordered_blocks(N=25, n=5, W=0.5, tl=2.7, seed=3)[:2],
             configs.barrier_blocks(0.4, 11, 1.0, 1)[:2]]
    for h, tnn in cases:
        n = h.shape[-1]
        scale = float(abs(tnn[0][0, 0]))
        Es = (np.linspace(-1.9, 1.9, 12) * scale).astype(complex)
        SL, SR = rgf_oracle.closed_selfenergies(h, tnn, Es)
        G0, W0 = rgf_oracle.rgf_blocks(h, tnn, Es, SL, SR)
        G1, W1, chunks = rgf_oracle.rgf_blocks_telescoped(h, tnn, Es, SL, SR, kmax, gmax, return_stats=True)
        assert sum(chunks) == len(h) and max(chunks) <= kmax
        R0, T0 = rgf_oracle.rt_from_blocks(G0, W0, SL, SR, np.eye(n))
        R1, T1 = rgf_oracle.rt_from_blocks(G1, W1, SL, SR, np.eye(n))
        assert rt_err(T1, T0) < 1e-10 and rt_err(R1, R0) < 1e-10
        if kmax == 1:
            assert np.allclose(G1, G0, rtol=1e-13, atol=1e-15) and np.allclose(W1, W0, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("name", ["nnn_n1_M8.npz", "nnn_n2_M6.npz", "nnn_n2_M7.npz", "nnn_n3_M5.npz"])
def test_next_nearest_neighbour_goldens(name):
    """nonzero tnnn (wfm/__init__.py:209-213): the dense restatement against the live reference's output"""
    g = load_golden(name)
    R, T = _sweep(g["h"], g["tnn"], g["tnnn"], g["Es"], "g_closed", g["sources"])
    assert rt_err(T, g["T"]) < TOL and rt_err(R, g["R"]) < TOL
    G = wfm_oracle.Green(g["h"], g["tnn"], g["tnnn"], 1.0, g["Es"][1], "g_closed")
    assert np.allclose(G, g["G"][1], rtol=1e-12, atol=1e-14)


def test_landauer_oracle_against_reference_fcdmft_golden():
    """oracle/landauer_oracle.py against the unmodified fcdmft.landauer (fcdmft/__init__.py:224-292); the same system as a
    wfm chain gives the same Tr[G_adv Lambda_R G_ret Lambda_L] as Caroli trace (current conservation)."""
    from oracle import landauer_oracle as lo
    g = load_golden("landauer_fcdmft.npz")
    eye = np.eye(2)
    LL = (0 * eye[None], float(g["V"]) * eye[None], g["cL"][None], float(g["muL"]))
    RL = (0 * eye[None], float(g["V"]) * eye[None], g["cR"][None], float(g["muR"]))
    for tag, kBT in (("kT0", 0.0), ("kT005", 0.05)):
        jE = lo.landauer(g["energies"], kBT, g["MBGF"], LL, RL)
        assert np.max(np.abs(jE - g["jE_" + tag])) < 1e-15
        assert abs(lo.current(g["energies"], jE) - float(g["I_" + tag])) < 1e-15
    Es = g["energies"].astype(complex)
    SL, SR = rgf_oracle.closed_selfenergies(g["h"], g["tnn"], Es)
    _, GN0 = rgf_oracle.rgf_blocks(g["h"], g["tnn"], Es, SL, SR)
    T = np.real(lo.transmission(g["energies"], g["MBGF"], LL, RL))
    assert np.allclose(rgf_oracle.caroli_trace(GN0, SL, SR), T, rtol=1e-12)


def test_extended_precision_oracles_agree_with_float64_and_each_other():
    """oracle/rgf_oracle_hp.py (the truth behind tests/golden/hp_cfg3_M10002.npz): the long-double recursion against the
    float64 one on a chain where both are well conditioned, and mpmath at 50 digits against long double on a short chain;
    and the fixture itself: unitarity of the stored truth, float64-oracle error consistent with what is stored."""
    from oracle import rgf_oracle_hp as hp
    from transport_b200 import configs
    h, tnn, _ = configs.disordered_blocks(N=120, n=8, W=0.05, seed=1234)
    Es = np.linspace(-1.7, 1.7, 5).astype(complex)
    R0, T0 = rgf_oracle.transmission_closed(h, tnn, Es, np.eye(8))
    R1, T1 = hp.transmission_closed_ld(h, tnn, Es, np.eye(8))
    assert rt_err(T0, T1.astype(float)) < 1e-11 and rt_err(R0, R1.astype(float)) < 1e-11
    assert float(np.max(np.abs(R1.sum(-1) + T1.sum(-1) - 1))) < 1e-16
    hs, ts, _ = configs.disordered_blocks(N=20, n=8, W=0.05, seed=7)
    R2, T2, unit = hp.transmission_closed_mp(hs, ts, Es[1], np.eye(8), dps=50)
    R3, T3 = hp.transmission_closed_ld(hs, ts, Es[1:2], np.eye(8))
    assert unit < 1e-40 and rt_err(T3[0].astype(float), T2) < 1e-14 and rt_err(R3[0].astype(float), R2) < 1e-14
    g = load_golden("hp_cfg3_M10002.npz")
    assert np.max(np.abs(g["R_hp"].sum(-1) + g["T_hp"].sum(-1) - 1)) < 1e-14          # mpmath truth, rounded to float64
    assert len(g["mp_checked"]) == len(g["idx"])                                       # every energy went through mpmath
    assert np.max(g["kappa"]) > 1e9 and np.max(g["f64_err"] / (g["kappa"] * 2.2e-16)) > 100
