"""CPU tests: the C-ABI library loads, exports every declared symbol, refuses to compute without a
GPU (no CPU fallback), and the Python shim reproduces the reference's argument errors."""
import ctypes

import numpy as np
import pytest

from transport_b200 import _lib, configs, wfm
from oracle.wfm_oracle import spin_flip_sums
from transport_b200.sweep import transmission_sweep


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_lib._SIGNATURES) <= set(names)
    assert b"sm_100a" in lib.tx_version()


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.tx_device_count() > 0:
        pytest.skip("a GPU is present")
    h, tnn, tnnn = configs.barrier_blocks(0.4, 3)
    with pytest.raises(_lib.TransportError) as ei:
        transmission_sweep(h, tnn, np.array([-1.0 + 0j]))
    assert ei.value.code == _lib.TX_ERR_CUDA
    with pytest.raises(_lib.TransportError):
        wfm.g_closed(np.zeros((1, 1), complex), -np.ones((1, 1), complex), -1.0, 1)
    with pytest.raises(_lib.TransportError):
        wfm.Hmat(h, tnn, tnnn)


def test_argument_errors_match_reference():
    h, tnn, tnnn = configs.barrier_blocks(0.4, 3)
    src = np.array([1.0])
    with pytest.raises(TypeError):                                   # wfm/__init__.py:57
        wfm.kernel(list(h), tnn, tnnn, 1.0, -1.0, "g_closed", src, False, False)
    with pytest.raises(ValueError):                                  # :179
        wfm.kernel(h, tnn[:-1], tnnn, 1.0, -1.0, "g_closed", src, False, False, all_debug=False)
    with pytest.raises(ValueError):                                  # :180
        wfm.Hmat(h, tnn, tnnn[:-1])
    with pytest.raises(AssertionError):                              # :70 source channel must sit at mu=0
        hb = h.copy(); hb[0, 0, 0] = 0.3
        wfm.kernel(hb, tnn, tnnn, 1.0, -1.0, "g_closed", src, False, False)
    with pytest.raises(AssertionError):                              # :74
        wfm.kernel(h, tnn, tnnn, 1.0, -1.0, "g_closed", np.array([1.0, 0.0]), False, False)
    with pytest.raises(Exception, match="not supported"):            # :294
        wfm.SelfEnergies(h, tnn, tnnn, -1.0, "g_magic")
    with pytest.raises(ValueError):                                  # :324
        wfm.g_closed(np.zeros((2, 2)), np.zeros((3, 3)), 0.0, 1)
    with pytest.raises(ValueError):                                  # :325
        wfm.g_closed(np.zeros((2, 2)), np.eye(2), 0.0, 2)
    with pytest.raises(Exception, match="Not diagonal"):             # :327
        wfm.g_closed(np.ones((2, 2)), np.eye(2), 0.0, 1)
    with pytest.raises(ValueError, match="Rice-Mele"):               # :370
        wfm.g_RiceMele(np.zeros((3, 3)), np.zeros((3, 3)), 0.0, 1)
    with pytest.raises(ValueError, match="Rice-Mele"):               # :375
        wfm.g_RiceMele(np.zeros((2, 2)), np.ones((2, 2)), 0.0, 1)
    with pytest.raises(NotImplementedError):                         # :502
        wfm.g_iter(np.eye(3), np.eye(3), 0.0, 1, 1e-2, 1e-3)
    with pytest.raises(TypeError):                                   # :537
        wfm.g_ith(np.eye(2), np.eye(2), 0.0, 1, 1e-2, 1.0, np.eye(2))
    with pytest.raises(Exception, match="Not diagonal"):
        hb = np.tile(h, (1, 2, 2)); tb = np.tile(tnn, (1, 2, 2)); hb[0] = 0.1
        transmission_sweep(hb, tb, np.array([-1.0 + 0j]))


def test_c_abi_rejects_bad_arguments_before_touching_the_gpu():
    lib = _lib.load()
    z = np.zeros(16, dtype=np.complex128)
    out = np.zeros(16)
    flag = np.zeros(1, dtype=np.int32)
    rc = lib.tx_transmission_sweep_dev(_lib.ptr(z), 1, None, None, _lib.ptr(z), 1, 1, 1, 1, _lib.ptr(z), 1,
                                       0, None, None, 1, 1, None, None, 1, _lib.ptr(out), _lib.ptr(out), None, None,
                                       _lib.ptr(flag), None)
    assert rc == _lib.TX_ERR_ARG and b"bad sizes" in lib.tx_last_error()
    rc = lib.tx_transmission_sweep_dev(_lib.ptr(z), 1, None, None, _lib.ptr(z), 1, 20, 3, 1, _lib.ptr(z), 1,
                                       0, None, None, 1, 1, None, None, 20, _lib.ptr(out), _lib.ptr(out), None, None,
                                       _lib.ptr(flag), None)
    assert rc == _lib.TX_ERR_UNSUPPORTED
    rc = lib.tx_surface_gf(_lib.ptr(z), _lib.ptr(z), 3, _lib.ptr(z), 1, 1, _lib.TX_GF_RICEMELE, 0.0, 0.0, 5, 100, 1000,
                           _lib.ptr(z), None, None, None, None)
    assert rc == _lib.TX_ERR_ARG
    assert lib.tx_set_variant(77) == _lib.TX_ERR_ARG
    assert lib.tx_set_variant(11) == _lib.TX_ERR_ARG      # the wrong-result timing experiments are gone
    assert lib.tx_set_variant(0) == _lib.TX_OK
    assert lib.tx_set_variant(6) == _lib.TX_OK
    assert lib.tx_set_variant(13) == _lib.TX_OK


def test_ricemele_helpers_match_golden_parameters():
    diag = np.array([[0.2, -0.8], [-0.8, -0.2]])
    off = np.array([[0, 0], [-1.1, 0]])
    edges = wfm.bandedges_RiceMele(diag, off)
    assert np.all(np.diff(edges) > 0)
    ks = np.linspace(0, np.pi, 7)
    bands = wfm.dispersion_RiceMele(diag, off, ks)
    assert bands.shape == (2, 7)
    assert np.isclose(bands[1].max(), edges[3]) and np.isclose(bands[1].min(), edges[2])
    kk = wfm.inverted_RiceMele(diag, off, bands[1])
    assert np.allclose(np.cos(kk), np.cos(ks), atol=1e-12)
    assert "u = 0.20" in wfm.string_RiceMele(diag, off, tex=False)


def test_spin_flip_sums():
    T = np.arange(64, dtype=float).reshape(1, 1, 8, 8)
    keep, flip = spin_flip_sums(T)
    assert keep[0, 0, 0] == T[0, 0, 0, :4].sum() and flip[0, 0, 0] == T[0, 0, 0, 4:].sum()
    assert keep[0, 0, 5] == T[0, 0, 5, 4:].sum() and flip[0, 0, 5] == T[0, 0, 5, :4].sum()


def test_bardeen_host_assembly_matches_oracle_hsysmat():
    """host-side slicing of the region parameters into block-tridiagonal arrays (no device needed)"""
    from oracle import bardeen_oracle as bo
    from transport_b200 import bardeen
    rng = np.random.default_rng(3)
    n, NC = 2, 5
    mats = [np.diag(rng.normal(size=n)) for _ in range(6)]             # tinfty, tL, tR, Vinfty, VL, VR
    HC = np.zeros((NC, NC, n, n), dtype=complex)
    for j in range(NC):
        HC[j, j] = rng.normal(size=(n, n))
    for j in range(NC - 1):
        HC[j, j + 1] = rng.normal(size=(n, n)); HC[j + 1, j] = rng.normal(size=(n, n))
    H = bo.Hsysmat(*mats, 3, 4, 6, HC)
    dg, up, lo = bardeen.hsys_site_blocks(*mats, 3, 4, 6, HC)
    S = H.shape[0]
    i = np.arange(S)
    assert np.array_equal(H[i, i], dg) and np.array_equal(H[i[:-1], i[:-1] + 1], up) and np.array_equal(H[i[:-1] + 1, i[:-1]], lo)
    H[i, i] = 0; H[i[:-1], i[:-1] + 1] = 0; H[i[:-1] + 1, i[:-1]] = 0
    assert not np.any(H)
    HC[0, 2] = 1.0
    with pytest.raises(NotImplementedError):
        bardeen.hsys_site_blocks(*mats, 3, 4, 6, HC)
    with pytest.raises(ValueError):                                    # bardeen/__init__.py:705 NC odd
        bardeen.hsys_site_blocks(*mats, 3, 4, 6, HC[:4, :4])


def test_bardeen_argument_errors_match_reference():
    from transport_b200 import bardeen
    one = np.eye(1)
    HC = np.zeros((3, 3, 1, 1), dtype=complex)
    with pytest.raises(ValueError):                                    # :70 shape mismatch
        bardeen.kernel_well(one, one, one, one, one, one, one, one, 2, 3, 3, HC, HC[:1, :1], one, one)
    with pytest.raises(ValueError):                                    # :73 len(defines_Sz)
        bardeen.kernel_well(one, one, one, one, one, one, one, one, 2, 3, 3, HC, HC, np.eye(2), one)
    with pytest.raises(AssertionError):                                # :432 HC == HCprime
        bardeen.kernel_well_prime(one, one, one, one, one, one, one, one, 2, 3, 3, HC, HC, one)
    with pytest.raises(ValueError, match="spin diagonal"):             # :81
        HC2 = np.zeros((3, 3, 2, 2), dtype=complex)
        bardeen.kernel_well_prime(np.eye(2), np.ones((2, 2)), np.eye(2), np.eye(2), np.eye(2), np.eye(2), np.eye(2),
                                  np.eye(2), 2, 3, 3, HC2 + 1, HC2, np.eye(2))
    with pytest.raises(TypeError):                                     # :541
        bardeen.current([[0.0]], np.ones((1, 1, 1)), np.array([0.1]), one, 0.0, 0.1)
    with pytest.raises(_lib.TransportError):                           # no CPU fallback
        if _lib.device_count() > 0:
            raise _lib.TransportError(2, "gpu present")
        bardeen.current(np.array([[0.0]]), np.ones((1, 1, 1)), np.array([0.1]), one, 0.0, 0.1)


@pytest.mark.parametrize("n,M", [(1, 6), (2, 7), (3, 5), (2, 4)])
def test_superblock_pairing_reproduces_the_pentadiagonal_matrix(n, M):
    """host logic of the next-nearest-neighbour path: pairing sites into super-blocks of size 2n turns the reference's
    block-pentadiagonal dense matrix (wfm.Hmat, wfm/__init__.py:168-215, restated in oracle.wfm_oracle.Hmat) into a
    block-tridiagonal one with the same entries (an odd chain gets one decoupled dummy site)"""
    from oracle import wfm_oracle
    from transport_b200 import superblocks
    rng = np.random.default_rng(10 * n + M)
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = (A + np.conj(np.swapaxes(A, 1, 2))) / 2
    tnn = rng.standard_normal((M - 1, n, n)) + 1j * rng.standard_normal((M - 1, n, n))
    tnnn = rng.standard_normal((M - 2, n, n)) + 1j * rng.standard_normal((M - 2, n, n))
    want = wfm_oracle.Hmat(h, tnn, tnnn)
    H, T, Mp = superblocks.pair_into_superblocks(h, tnn, tnnn)
    K = H.shape[0]
    assert Mp == M + M % 2 and H.shape == (K, 2 * n, 2 * n) and T.shape == (K - 1, 2 * n, 2 * n)
    got = wfm_oracle.Hmat(H, T, np.zeros((max(K - 2, 0), 2 * n, 2 * n), complex))
    assert np.array_equal(got[:M * n, :M * n], want)
    if M % 2:                                            # the dummy site is decoupled
        assert not np.any(got[:M * n, M * n:]) and not np.any(got[M * n:, :M * n])
