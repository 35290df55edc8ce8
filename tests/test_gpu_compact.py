"""GPU parity tests of the compact outputs and the context API (run with -m gpu on a B200).

The spin-resolved sums the reference's drivers form on the host from wfm.kernel's per-channel arrays
(run_wfm_gate.py:404-405,473-477) are produced by the device epilogue; they must equal the same sums taken
over the reference's golden R/T, for every kernel family (n <= 4 register kernel, n = 8 telescoped / plain /
register-resident, 8 < n <= 16, CTA-per-point kernel for n > 16), with and without the per-channel arrays.
"""
import threading

import numpy as np
import pytest

from conftest import load_golden, rt_err
from oracle import rgf_oracle
from oracle.wfm_oracle import spin_flip_sums
from transport_b200 import _lib, configs
from transport_b200.sweep import Context, transmission_sweep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if _lib.device_count() < 1:
        pytest.fail("no CUDA device: the gpu-marked tests must run on a GPU box")


def _want_spin(R, T, n_up=None, sources=None, spin_n=4):
    tn, tf = spin_flip_sums(T, n_up, sources)
    if spin_n == 2:
        return np.stack([tn, tf], axis=-1)
    rn, rf = spin_flip_sums(R, n_up, sources)
    return np.stack([tn, tf, rn, rf], axis=-1)


def _close(got, want, tol=1e-10):
    scale = np.maximum(np.abs(want), 1e-13 * np.abs(want).max(axis=-1, keepdims=True))
    assert np.max(np.abs(got - want) / np.where(scale == 0, 1, scale)) < tol


@pytest.mark.parametrize("variant", [13, 6, 0])
def test_spin_sums_cfg2_golden_all_kernels(variant):
    """cfg2 golden (the reference's own R/T): spin sums with and without the per-channel arrays."""
    lib = _lib.load()
    g = load_golden("cfg2_cicc.npz")
    want4 = _want_spin(g["R"], g["T"])
    try:
        _lib.check(lib.tx_set_variant(variant))
        R, T, spin = transmission_sweep(g["h"], g["tnn"], g["Es"], want_spin=4)
        assert rt_err(T, g["T"]) < 1e-10 and rt_err(R, g["R"]) < 1e-10
        _close(spin, want4)
        # the device sums are the sums of the device's own channels, to round-off
        assert np.allclose(spin, _want_spin(R, T), rtol=1e-13, atol=1e-300)
        # compact: no R/T at all; T pair only
        tot, spin2 = transmission_sweep(g["h"], g["tnn"], g["Es"], want_rt=False, want_total=True, want_spin=2)
        _close(spin2, want4[..., :2])
        assert np.array_equal(spin2, spin[..., :2])
        assert np.allclose(tot, g["T"].sum(axis=(2, 3)), rtol=1e-10)
        only = transmission_sweep(g["h"], g["tnn"], g["Es"], want_rt=False, want_spin=4)
        assert np.array_equal(only, spin)
        # a different sector boundary and explicit (non one-hot) source vectors
        srcs = np.array([[1.0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0.6, 0.8, 0], [0.5, 0.5, 0, 0, 0, 0, 0.1, 0]])
        R3, T3, spin3 = transmission_sweep(g["h"], g["tnn"], g["Es"], sources=srcs, want_spin=4, n_elec_up=2)
        assert np.allclose(spin3, _want_spin(R3, T3, 2, srcs), rtol=1e-13, atol=1e-300)
    finally:
        _lib.check(lib.tx_set_variant(13))


@pytest.mark.parametrize("name", ["kondo_n2.npz", "general_hop_n4.npz", "general_hop_n6.npz", "general_hop_n12.npz",
                                  "general_hop_n16.npz", "cfg3_disorder_N30.npz", "cicc_spin1_n18.npz",
                                  "cicc_spin32_n32.npz"])
def test_spin_sums_other_block_sizes(name):
    g = load_golden(name)
    n = g["h"].shape[-1]
    srcs = g["sources"] if "sources" in g.files else np.eye(n)
    if name.startswith("cicc_spin1_"):
        srcs = np.eye(18)[[0, 7]]
    gR, gT = g["R"], g["T"]
    R, T, spin = transmission_sweep(g["h"], g["tnn"], g["Es"], sources=srcs, want_spin=4)
    assert rt_err(T[0], gT) < 1e-10 and rt_err(R[0], gR) < 1e-10
    _close(spin[0], _want_spin(gR, gT, None, srcs))
    spin_only = transmission_sweep(g["h"], g["tnn"], g["Es"], sources=srcs, want_rt=False, want_spin=4)
    assert np.array_equal(spin_only, spin)


def test_spin_sums_large_blocks_n64():
    n, M = 64, 4
    rng = np.random.default_rng(5)
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = 0.2 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2
    h[0] = 0
    h[-1] = np.diag(0.02 * rng.standard_normal(n))
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    Es = np.array([-1.1, 0.4], dtype=complex)
    srcs = np.eye(n)[[3, 40]]
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Es, srcs)
    spin = transmission_sweep(h, tnn, Es, sources=srcs, want_rt=False, want_spin=4, n_elec_up=24)
    _close(spin[0], _want_spin(oR, oT, 24, srcs))


def test_compact_full_size_cfg2_matches_full_outputs():
    """Size-independent property at BASELINE's full cfg2 size: the compact sweep (no R/T arrays) gives bit-identical
    sums to the full sweep's epilogue, and they add up to the per-point total and to unitarity."""
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    Js, Es = configs.cfg2_grid(1000, 1000)
    tot, spin = transmission_sweep(h0, tnn, Es, h1=h1, params=Js, want_rt=False, want_total=True, want_spin=4)
    assert spin.shape == (1000, 1000, 8, 4) and np.isfinite(spin).all()
    assert np.max(np.abs(spin.sum(-1) - 1.0)) < 1e-10                       # R + T = 1 per source
    assert np.allclose(tot, spin[..., :2].sum(axis=(2, 3)), rtol=1e-12)
    g = load_golden("cfg2_cicc.npz")
    want = _want_spin(g["R"], g["T"])
    for k, jidx in enumerate([0, 250, 500, 750, 999]):
        _close(spin[jidx, ::40], want[k])


def test_contexts_run_concurrently_from_threads():
    """Two host threads, one context each (own workspace, own stream): same bits as the default context."""
    g = load_golden("cfg2_cicc.npz")
    R0, T0 = transmission_sweep(g["h"], g["tnn"], g["Es"])
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    Js, Es = configs.cfg2_grid(64, 256)
    Ra, Ta = transmission_sweep(h0, tnn, Es, h1=h1, params=Js)
    out, errs = {}, []

    def work(key, reps):
        try:
            with Context(0) as cx:
                for _ in range(reps):
                    if key == "a":
                        out[key] = transmission_sweep(h0, tnn, Es, h1=h1, params=Js, context=cx)
                    else:
                        out[key] = transmission_sweep(g["h"], g["tnn"], g["Es"], context=cx)
        except Exception as exc:      # surfaced below
            errs.append(exc)

    ts = [threading.Thread(target=work, args=("a", 6)), threading.Thread(target=work, args=("b", 40))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert np.array_equal(out["a"][0], Ra) and np.array_equal(out["a"][1], Ta)
    assert np.array_equal(out["b"][0], R0) and np.array_equal(out["b"][1], T0)
    # per-context options do not leak into the default context
    with Context(0) as cx:
        cx.set_option(4, 6)
        R6, T6 = transmission_sweep(g["h"], g["tnn"], g["Es"], context=cx)
    assert rt_err(T6, T0) < 1e-11
    R1, T1 = transmission_sweep(g["h"], g["tnn"], g["Es"])
    assert np.array_equal(T1, T0)


def test_large_block_sweeps_with_device_built_leads_from_two_threads():
    """two host threads, one context and stream each, large blocks with Sancho-Rubio lead tables built on the device: the
    lead kernel's per-device workspace and each context's sweep workspace are shared resources whose reuse across streams
    is ordered by events on the device; results equal the serial ones bit for bit"""
    hA, tA, _ = configs.ribbon_blocks(width=64, L=5, W=1.0, seed=5)
    hB, tB, _ = configs.ribbon_blocks(width=72, L=4, W=0.5, seed=6)
    EsA = np.linspace(-2.5, 2.5, 40).astype(complex)
    EsB = np.linspace(-2.0, 2.0, 24).astype(complex)
    kwA = dict(converger=("sancho", 1e-6), sources=np.eye(64)[:2], want_caroli=True)
    kwB = dict(converger=("sancho", 1e-6), sources=np.eye(72)[:2], want_caroli=True)
    wantA = transmission_sweep(hA, tA, EsA, **kwA)
    wantB = transmission_sweep(hB, tB, EsB, **kwB)
    out, errs = {}, []

    def work(key, reps):
        try:
            with Context(0) as cx:
                for _ in range(reps):
                    out[key] = transmission_sweep(hA, tA, EsA, context=cx, **kwA) if key == "a" else \
                        transmission_sweep(hB, tB, EsB, context=cx, **kwB)
        except Exception as exc:      # surfaced below
            errs.append(exc)

    ts = [threading.Thread(target=work, args=("a", 5)), threading.Thread(target=work, args=("b", 5))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for got, want in ((out["a"], wantA), (out["b"], wantB)):
        for x, y in zip(got, want):
            assert np.array_equal(x, y)


def test_device_built_lead_tables_through_the_host_sweep():
    """TX_LEAD_SANCHO / TX_LEAD_SEPARABLE: the host sweep builds the self-energy tables on the device ahead of the sweep
    (they never travel to the host).  Same numbers as the two-call route (tables to the host and back), for a large
    block (n = 24, CTA-per-point kernel) and a small one (n = 8, register kernel in table mode)."""
    from transport_b200 import leads
    for n, M in ((24, 6), (8, 7)):
        h, tnn, _ = configs.ribbon_blocks(width=n, L=M - 2, W=0.8, seed=3)
        Es = np.linspace(-2.5, 2.5, 16).astype(complex)
        for conv in (("sancho", 1e-6), ("separable", 1e-6), "separable"):
            SL, SR = leads.self_energies_batched(h, tnn, Es, conv)
            srcs = np.eye(n)[:2]
            R0, T0, c0 = transmission_sweep(h, tnn, Es, "table", sigL=SL, sigR=SR, sources=srcs, want_caroli=True)
            R1, T1, c1 = transmission_sweep(h, tnn, Es, conv, sources=srcs, want_caroli=True)
            assert np.array_equal(c1, c0) and np.array_equal(T1, T0) and np.array_equal(R1, R0)
            only = transmission_sweep(h, tnn, Es, conv, sources=srcs, want_rt=False, want_caroli=True)
            assert np.array_equal(only, c0)
    # a decimation that cannot converge inside its iteration cap is reported, not returned
    h, tnn, _ = configs.ribbon_blocks(width=24, L=3, W=0.0, seed=3)
    with pytest.raises(Exception, match="did not converge"):
        transmission_sweep(h, tnn, np.array([0.3 + 0j]), ("sancho", 1e-12, 1e-14, 3), want_rt=False, want_caroli=True)


def test_context_on_a_caller_stream_and_device_entry():
    """tx_context_set_stream: the host entry issues on the caller's stream; the device entry on two different streams
    through one context (scratch reuse ordered by an event) gives the bits of the default path."""
    import ctypes
    torch = pytest.importorskip("torch")
    g = load_golden("cfg2_cicc.npz")
    R0, T0 = transmission_sweep(g["h"], g["tnn"], g["Es"])
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with Context(0, stream=s1.cuda_stream) as cx:
        R1, T1 = transmission_sweep(g["h"], g["tnn"], g["Es"], context=cx)
    assert np.array_equal(T1, T0) and np.array_equal(R1, R0)
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    Js, Es = configs.cfg2_grid(64, 64)
    want_R, want_T = transmission_sweep(h0, tnn, Es, h1=h1, params=Js)

    def dc(a):
        return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))).to(dev)
    d_h0, d_h1, d_t, d_E = dc(h0), dc(h1), dc(tnn), dc(Es)
    halves = []
    ctx = ctypes.c_void_p()
    _lib.check(lib.tx_context_create(0, ctypes.byref(ctx)))
    try:
        for st, sl in ((s1, slice(0, 32)), (s2, slice(32, 64))):
            dJ = torch.from_numpy(np.ascontiguousarray(Js[sl])).to(dev)
            R = torch.empty((32 * 64, 8, 8), dtype=torch.float64, device=dev)
            T = torch.empty_like(R)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            _lib.check(lib.tx_transmission_sweep_spin_dev(ctx, d_h0.data_ptr(), 1, d_h1.data_ptr(), dJ.data_ptr(), d_t.data_ptr(), 1,
                                                          8, 8, 32, d_E.data_ptr(), 64, _lib.TX_LEAD_CLOSED, None, None, 1,
                                                          _lib.TX_HOP_SCALAR, None, None, 8, R.data_ptr(), T.data_ptr(), None, None,
                                                          0, 0, None, flag.data_ptr(), st.cuda_stream))
            halves.append((R, T, flag, dJ))
        torch.cuda.synchronize()
    finally:
        lib.tx_context_destroy(ctx)
    got_T = np.concatenate([hh[1].cpu().numpy().reshape(32, 64, 8, 8) for hh in halves])
    got_R = np.concatenate([hh[0].cpu().numpy().reshape(32, 64, 8, 8) for hh in halves])
    assert np.array_equal(got_T, want_T) and np.array_equal(got_R, want_R)


def test_device_entry_under_cuda_graph_capture():
    """the stream-asynchronous device entry is capture-safe after one warm-up call (which sizes the context's scratch and
    sets the kernel attributes): a CUDA graph of the whole sweep (structure scan, affine expansion, both sweep
    instantiations) replays to the bits of the direct call, also after the parameters in the captured buffers change"""
    import ctypes
    torch = pytest.importorskip("torch")
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    Js, Es = configs.cfg2_grid(48, 64)

    def dc(a):
        return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))).to(dev)
    d_h0, d_h1, d_t, d_E = dc(h0), dc(h1), dc(tnn), dc(Es)
    dJ = torch.from_numpy(np.ascontiguousarray(Js)).to(dev)
    R = torch.zeros((48 * 64, 8, 8), dtype=torch.float64, device=dev)
    T = torch.zeros_like(R)
    spin = torch.zeros((48 * 64, 8, 2), dtype=torch.float64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ctx = ctypes.c_void_p()
    _lib.check(lib.tx_context_create(0, ctypes.byref(ctx)))
    s = torch.cuda.Stream()

    def launch():
        _lib.check(lib.tx_transmission_sweep_spin_dev(ctx, d_h0.data_ptr(), 1, d_h1.data_ptr(), dJ.data_ptr(), d_t.data_ptr(), 1,
                                                      8, 8, 48, d_E.data_ptr(), 64, _lib.TX_LEAD_CLOSED, None, None, 1,
                                                      _lib.TX_HOP_SCALAR, None, None, 8, R.data_ptr(), T.data_ptr(), None, None,
                                                      4, 2, spin.data_ptr(), flag.data_ptr(), s.cuda_stream))
    try:
        with torch.cuda.stream(s):
            launch()                                  # warm-up outside the capture
            s.synchronize()
            want_R, want_T = transmission_sweep(h0, tnn, Es, h1=h1, params=Js)
            assert np.array_equal(T.cpu().numpy().reshape(48, 64, 8, 8), want_T)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                launch()
            for Jscale in (1.0, 0.5):
                dJ.copy_(torch.from_numpy(np.ascontiguousarray(Js * Jscale)))
                R.zero_(); T.zero_(); spin.zero_()
                torch.cuda.synchronize()
                graph.replay()
                torch.cuda.synchronize()
                want_R, want_T, want_spin = transmission_sweep(h0, tnn, Es, h1=h1, params=Js * Jscale, want_spin=2, n_elec_up=4)
                assert np.array_equal(T.cpu().numpy().reshape(48, 64, 8, 8), want_T)
                assert np.array_equal(R.cpu().numpy().reshape(48, 64, 8, 8), want_R)
                assert np.array_equal(spin.cpu().numpy().reshape(48, 64, 8, 2), want_spin)
        assert int(flag.item()) == 0
    finally:
        torch.cuda.synchronize()
        lib.tx_context_destroy(ctx)
