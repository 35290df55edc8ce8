"""GPU parity tests of the 64 < n_loc_dof <= 128 size class (512-thread instantiations of the large-block kernels):
run_wfm_gate.get_hblocks gives n = 2 (2s+1)^2 = 72 for two spin-5/2 impurities and 128 for spin-7/2, which the reference
handles through its dense inverse (wfm/__init__.py:80-99).  Same gates as test_gpu_parity.py: 1e-10 on every R/T channel
and on R + T = 1, against goldens of the live reference (oracle/make_golden_r2.py cicc_large_spins) and the numpy oracle."""
import numpy as np
import pytest

from conftest import load_golden, rt_err
from oracle import rgf_oracle, wfm_oracle
from transport_b200 import _lib, configs, green, leads, wfm
from transport_b200.sweep import transmission_sweep

pytestmark = pytest.mark.gpu

RTOL = 1e-10
UNIT = 1e-10


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on a GPU box")


def _random_chain(n, M, scalar, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = 0.2 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2 / np.sqrt(n / 32)
    h[0] = np.diag(np.zeros(n))
    h[-1] = np.diag(0.02 * rng.standard_normal(n))
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    if not scalar:
        tnn[1:-1] += 0.1 * (rng.standard_normal((M - 3, n, n)) + 1j * rng.standard_normal((M - 3, n, n))) / np.sqrt(n / 32)
    return h, tnn


@pytest.mark.parametrize("n", [50, 72, 98, 128])
def test_reference_goldens_two_large_spins(n):
    """the reference's own R/T for two spin-2 (n = 50), 5/2 (72), 3 (98) and 7/2 (128) impurities; the blocks come from
    configs.cicc_blocks, which the generating script holds bit-equal to run_wfm_gate.get_hblocks"""
    g = load_golden("cicc_builder_n%d.npz" % n)
    a = g["builder_args"]
    h, tnn, tnnn = configs.cicc_blocks(int(a[0]), a[1], a[2], a[3], a[4], a[5], int(a[6]), int(g["the_dist"]))
    assert h.shape[-1] == n and not np.any(tnnn)
    R, T, res = transmission_sweep(h, tnn, g["Es"], sources=g["sources"], want_resid=True)
    assert rt_err(T[0], g["T"]) < RTOL and rt_err(R[0], g["R"]) < RTOL
    assert np.max(np.abs(res)) < UNIT
    # the reference's call signature, one energy and source at a time
    Rs, Ts = wfm.kernel(h, tnn, tnnn, 1.0, g["Es"][1], "g_closed", g["sources"][1], False, False, all_debug=True)
    assert rt_err(Ts, g["T"][1, 1]) < RTOL and rt_err(Rs, g["R"][1, 1]) < RTOL


@pytest.mark.parametrize("n,scalar", [(72, True), (72, False), (80, True), (96, True), (104, False), (112, True), (120, False),
                                      (128, True), (128, False), (70, True), (99, False), (98, True), (50, True), (18, True), (121, True)])
def test_wide_block_sweep_and_caroli(n, scalar):
    """telescoped (scalar hopping) and plain recursion, inverse staging in shared memory (n <= 104) and in the global
    workspace, block edges that are not a multiple of 8 (telescoped sweep: work blocks padded to the tensor-core tile with
    decoupled channels; plain recursion: scalar products, column-by-column inverse)"""
    M = 5
    h, tnn = _random_chain(n, M, scalar, n)
    Es = np.array([-1.4, 0.3, 1.2], dtype=complex)
    srcs = np.eye(n)[[0, n // 2]]
    R, T, res, tot, car = transmission_sweep(h, tnn, Es, sources=srcs, want_resid=True, want_total=True, want_caroli=True)
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, srcs)
    assert rt_err(T[0], oT) < RTOL and rt_err(R[0], oR) < RTOL
    assert np.max(np.abs(res)) < UNIT
    SL, SR = rgf_oracle.closed_selfenergies(h, tnn, Es)
    _, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es, SL, SR)
    assert np.allclose(car[0], rgf_oracle.caroli_trace(GN0, SL, SR), rtol=1e-10)
    assert np.allclose(tot[0], T[0].sum(axis=(1, 2)), rtol=1e-12)


def test_wide_block_sweep_many_points_spin_sums_and_affine_family():
    """more points than resident CTAs (the persistent grid strides), an affine family of Hamiltonians, spin-resolved sums"""
    n, M = 72, 4
    h, tnn = _random_chain(n, M, True, 7)
    h1 = np.zeros_like(h)
    h1[1] = np.diag(np.linspace(-1, 1, n))
    Js = np.linspace(-0.3, 0.3, 5)
    Es = np.linspace(-1.8, 1.8, 64).astype(complex)
    srcs = np.eye(n)[[3]]
    T_only, spin = transmission_sweep(h, tnn, Es, h1=h1, params=Js, sources=srcs, want_rt=False, want_total=True, want_spin=4)
    R, T = transmission_sweep(h, tnn, Es, h1=h1, params=Js, sources=srcs)
    assert np.allclose(T_only, T.sum(axis=(2, 3)), rtol=1e-12)
    tn, tf = wfm_oracle.spin_flip_sums(T)
    rn, rf = wfm_oracle.spin_flip_sums(R)
    assert np.allclose(spin, np.stack([tn, tf, rn, rf], axis=-1), rtol=1e-11, atol=1e-300)
    for p in (0, 4):
        oR, oT = rgf_oracle.transmission_closed(h + Js[p] * h1, tnn, Es[[0, 31, 63]], srcs)
        assert rt_err(T[p][[0, 31, 63]], oT) < RTOL and rt_err(R[p][[0, 31, 63]], oR) < RTOL


@pytest.mark.parametrize("n", [72, 88, 96, 104, 112, 128, 98, 110])
def test_staged_inverse_wide(n):
    """blocked Gauss-Jordan with 512 threads (n % 8 == 0), staging in shared memory (n <= 104) or global memory, and the
    column-by-column inverse for the other block edges, against numpy"""
    from transport_b200._lib import as_c128, check, ptr
    rng = np.random.default_rng(n)
    count = 5
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


@pytest.mark.parametrize("n", [72, 112, 50])
def test_sancho_ribbon_leads_wide(n):
    """Sancho-Rubio lead tables of a ribbon wider than 64 channels against the numpy decimation and the fixed-point map;
    the sweep with the same tables against the oracle's Caroli trace; clean ribbon: integer channel count"""
    L = 4
    Es = np.array([-2.3, 0.15, 1.9], dtype=complex)
    h, tnn, _ = configs.ribbon_blocks(width=n, L=L, W=0.0, seed=3)
    eta = 1e-3
    SL, SR = leads.self_energies_batched(h, tnn, Es, ("sancho", eta, 1e-14))
    gL, _ = rgf_oracle.sancho_rubio(h[0], tnn[0], Es + 1j * eta, -1)
    gR, _ = rgf_oracle.sancho_rubio(h[-1], tnn[-1], Es + 1j * eta, 1)
    oSL = np.conj(tnn[0]).T @ gL @ tnn[0]
    oSR = tnn[-1] @ gR @ np.conj(tnn[-1]).T
    for dev_S, ora_S in ((SL, oSL), (SR, oSR)):
        # two correct decimations differ by round-off amplified by |G|^2 <= 1/eta^2 (measured up to 2e-10 at this eta)
        assert np.abs(dev_S - ora_S).max() / np.abs(ora_S).max() < 1e-9
    R, T, car = transmission_sweep(h, tnn, Es, "table", sigL=oSL, sigR=oSR, sources=np.eye(n)[:1], want_caroli=True)
    _, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es, oSL, oSR)
    assert np.allclose(car[0], rgf_oracle.caroli_trace(GN0, oSL, oSR), rtol=1e-10, atol=1e-12)
    # device-built tables inside the host sweep (TX_LEAD_SANCHO), small eta: channel count of the clean ribbon
    car2 = transmission_sweep(h, tnn, Es, ("sancho", 1e-8), sources=np.eye(n)[:1], want_rt=False, want_caroli=True)
    eps = np.linalg.eigvalsh(h[0].real)
    open_ch = np.array([np.sum(np.abs(E.real - eps) < 2.0) for E in Es])
    assert np.allclose(np.asarray(car2).reshape(-1), open_ch, atol=1e-4)


def test_padded_block_edge_with_affine_family_sources_and_spin_sums():
    """n = 50 (two spin-2 impurities' block edge) through the padded telescoped sweep: affine family, superposed sources,
    spin-resolved sums, unitarity; against the oracle and against the unpadded plain sweep (kmax = 1)"""
    n, M = 50, 6
    h, tnn = _random_chain(n, M, True, 11)
    h1 = np.zeros_like(h)
    h1[2] = np.diag(np.linspace(-1, 1, n))
    Js = np.array([-0.4, 0.0, 0.7])
    Es = np.array([-1.7, -0.2, 0.9, 1.6], dtype=complex)
    rng = np.random.default_rng(3)
    srcs = np.vstack([np.eye(n)[[0, 49]], rng.standard_normal((1, n))])
    R, T, res, tot, spin = transmission_sweep(h, tnn, Es, h1=h1, params=Js, sources=srcs, want_resid=True, want_total=True,
                                              want_spin=4)
    assert np.max(np.abs(res)) < UNIT
    for p in range(3):
        oR, oT = rgf_oracle.transmission_closed(h + Js[p] * h1, tnn, Es, srcs)
        assert rt_err(T[p], oT) < RTOL and rt_err(R[p], oR) < RTOL
    assert np.allclose(tot, T.sum(axis=(2, 3)), rtol=1e-12)
    assert np.allclose(spin[..., 0] + spin[..., 1], T.sum(axis=-1), rtol=1e-12)
    assert np.allclose(spin[..., 2] + spin[..., 3], R.sum(axis=-1), rtol=1e-12)
    lib = _lib.load()
    _lib.check(lib.tx_set_option(1, 1))          # one inverse per site: no padding, scalar products
    try:
        R1, T1 = transmission_sweep(h, tnn, Es, h1=h1, params=Js, sources=srcs)
    finally:
        _lib.check(lib.tx_set_option(1, 32))
    assert rt_err(T1, T) < RTOL and rt_err(R1, R) < RTOL


@pytest.mark.parametrize("n", [96, 104])
def test_tma_staged_operand_of_the_wide_products(n):
    """n = 96 ... 104: the run-time-n product stages its B operand in shared memory by TMA (tx_set_option(5, .)); same
    results as with fragments straight from L2, for the sweep and for the Sancho-Rubio lead tables"""
    M = 5
    h, tnn = _random_chain(n, M, True, n + 5)
    Es = np.array([-1.4, 0.3, 1.2], dtype=complex)
    srcs = np.eye(n)[[1, n - 1]]
    hr, tr, _ = configs.ribbon_blocks(width=n, L=4, W=0.5, seed=2)
    lib = _lib.load()
    out = []
    try:
        for mode in (1, 0):
            _lib.check(lib.tx_set_option(5, mode))
            R, T, car = transmission_sweep(h, tnn, Es, sources=srcs, want_caroli=True)
            SL, SR = leads.self_energies_batched(hr, tr, Es, ("sancho", 1e-4, 1e-14))
            out.append((R, T, car, SL, SR))
    finally:
        _lib.check(lib.tx_set_option(5, 1))
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, srcs)
    for R, T, car, SL, SR in out:
        assert rt_err(T[0], oT) < RTOL and rt_err(R[0], oR) < RTOL
    for a, b in zip(out[0], out[1]):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("n", [24, 50, 72])
def test_all_one_hot_sources_shortcut_equals_explicit_sources(n):
    """sources=None (every channel as a one-hot source: O(n) per source in the epilogue) against the same sources passed
    explicitly (general superposition path), all outputs"""
    h, tnn = _random_chain(n, 4, True, n + 9)
    Es = np.array([-1.1, 0.6], dtype=complex)
    a = transmission_sweep(h, tnn, Es, want_resid=True, want_total=True, want_spin=4)
    b = transmission_sweep(h, tnn, Es, sources=np.eye(n), want_resid=True, want_total=True, want_spin=4)
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.allclose(x, y, rtol=1e-13, atol=1e-15)
    assert np.max(np.abs(a[2])) < UNIT


@pytest.mark.parametrize("n", [24, 50, 64, 72, 128])
def test_diagonal_sites_as_row_scalings_in_the_large_block_sweep(n):
    """chains whose interior on-site blocks are mostly diagonal (spacer sites between two impurities, as
    run_wfm_gate.get_hblocks builds them): the telescoped large-block sweep turns those sites into row scalings (structure
    scan, tx_set_option(0, .)); against the oracle, and against the same sweep with the scan switched off; affine family
    whose affine term is diagonal on some sites and dense on others"""
    M = 9
    rng = np.random.default_rng(n + 17)
    h, tnn = _random_chain(n, M, True, n + 17)
    dense = (2, 6)
    for j in range(1, M - 1):
        if j not in dense:
            h[j] = np.diag(0.3 * rng.standard_normal(n))
    h1 = np.zeros_like(h)
    h1[2] = h[2] * 0.5                                   # dense affine term on a dense site
    h1[4] = np.diag(rng.standard_normal(n))              # diagonal affine term on a diagonal site
    Js = np.array([-0.5, 0.8])
    Es = np.array([-1.6, -0.3, 0.7, 1.5], dtype=complex)
    srcs = np.eye(n)[[0, n - 1]]
    lib = _lib.load()
    outs = {}
    try:
        for scan in (1, 0):
            _lib.check(lib.tx_set_option(0, scan))
            outs[scan] = transmission_sweep(h, tnn, Es, h1=h1, params=Js, sources=srcs, want_resid=True, want_caroli=True)
    finally:
        _lib.check(lib.tx_set_option(0, 1))
    R, T, res, car = outs[1]
    assert np.max(np.abs(res)) < UNIT
    for p in range(2):
        oR, oT = rgf_oracle.transmission_closed(h + Js[p] * h1, tnn, Es, srcs)
        assert rt_err(T[p], oT) < RTOL and rt_err(R[p], oR) < RTOL
    for a, b in zip(outs[1], outs[0]):
        assert np.allclose(a, b, rtol=1e-11, atol=1e-13)


def test_green_blocks_wide():
    """first block column and full Green's function at n = 72 (work blocks in shared memory) and n = 96 (global scratch)"""
    for n in (72, 96):
        M = 4
        h, tnn = _random_chain(n, M, False, n + 1)
        Es = np.array([-1.2, 0.3], dtype=complex)
        SL, SR = rgf_oracle.closed_selfenergies(h, tnn, Es)
        dSL, dSR = leads.self_energies_batched(h, tnn, Es, "g_closed")
        assert np.allclose(dSL, SL, rtol=1e-13, atol=1e-15) and np.allclose(dSR, SR, rtol=1e-13, atol=1e-15)
        col = green.green_column0(h, tnn, Es, SL, SR)
        assert np.allclose(col, rgf_oracle.green_column0(h, tnn, Es, SL, SR), rtol=1e-10, atol=1e-13)
        full = green.green_full(h, tnn, Es, SL, SR)
        want = np.array([wfm_oracle.Green(h, tnn, np.zeros((M - 2, n, n), complex), 1.0, E, "g_closed") for E in Es])
        assert np.allclose(full, want, rtol=1e-10, atol=1e-12)


def test_block_edge_beyond_the_limit_is_refused():
    n, M = 136, 3
    h = np.zeros((M, n, n), complex)
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    with pytest.raises(_lib.TransportError) as ei:
        transmission_sweep(h, tnn, np.array([0.1 + 0j]), sources=np.eye(n)[:1])
    assert ei.value.code == _lib.TX_ERR_UNSUPPORTED
