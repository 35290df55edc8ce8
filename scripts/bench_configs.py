#!/usr/bin/env python
"""BASELINE configs 1, 3, 4 and 5 (config 2 is bench.py's headline) under the same measurement rules as bench.py.

    python scripts/bench_configs.py [cfg1] [cfg3] [cfg4] [cfg5] [--small] [--steps K]

bench.py imports `run(name, env, ...)` for its `--config` flag and for the `secondary` block of its default line.
Every config reports: device-resident throughput (CUDA events, >= 3 warm-ups, L2 flush between steps, max over
ranks), roofline of the dominant kernel, e2e through the public host-buffer API, the CPU baseline (reference-loop
port where the reference's algorithm can run, labelled otherwise), parity inside the run, and clocks sampled during
the timed region.  Multi-GPU: the energy (cfg1/3/4) or geometry (cfg5) axis is sharded, no collective.
"""
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import benchutil as bu  # noqa: E402
from transport_b200 import _lib, configs, shard  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

WORKLOADS = {
    "cfg1": "cfg1: rbbarrier spinless chain with rectangular barrier, n_loc_dof=1, M=13, 10^4 energies",
    "cfg2": "cfg2: rbmenez-style two spin-1/2 impurities, n_loc_dof=8, M=8, 1000 J x 1000 E, all 8 sources",
    "cfg3": "cfg3: disordered scattering region, n_loc_dof=8, N=10^4 sites (M=10002), 10^5 energies, all 8 sources",
    "cfg4": "cfg4: square-lattice ribbon, n_loc_dof=64, L=32 slices (M=34), 4096 energies, Sancho-Rubio leads eta=1e-8, Caroli trace",
    "cfg5": "cfg5: bardeen rbstep sweep, 256 step heights per pass, N_L=N_R=200 and 800 (S=451, 1651), 99 bias points",
}


def _slice(n_base, env, scaling):
    """This rank's contiguous slice of the unit axis: weak = a grid world x finer, strong = the base grid split."""
    total = n_base * env.world if scaling == "weak" else n_base
    lo, hi = shard.split_range(total, env.world, env.rank)
    return total, lo, hi


def _std(name, env, scaling, units_total, ms, extra):
    out = {"workload": WORKLOADS[name], "value": units_total / (ms * 1e-3), "unit": bu.UNIT, "ms_per_step": ms,
           "n_gpus": env.world, "scaling": scaling, "units_per_step_all_gpus": int(units_total)}
    out.update(extra)
    return out


# ---------------------------------------------------------------------------------------------- CPU legs
def _cpu_kernel_loops(args):
    """Reference-loop port (oracle.wfm_oracle.kernel_loops) over a list of energies x sources."""
    bu.one_thread_blas()
    from oracle import wfm_oracle
    h, tnn, tnnn, Es, srcs = args
    t0 = time.perf_counter()
    acc = 0.0
    for E in Es:
        for s in srcs:
            Rs, Ts = wfm_oracle.kernel_loops(h, tnn, tnnn, 1.0, E, "g_closed", s, False, False, all_debug=False)
            acc += Ts.sum()
    return time.perf_counter() - t0, acc


def _cpu_rgf(args):
    bu.one_thread_blas()
    from oracle import rgf_oracle
    h, tnn, Es, srcs = args
    t0 = time.perf_counter()
    rgf_oracle.transmission_closed(h, tnn, Es, srcs)
    return time.perf_counter() - t0, 0.0


def _cpu_ribbon(args):
    """numpy restatement of cfg4's pipeline for a few energies: Sancho-Rubio leads + recursive sweep + Caroli trace."""
    bu.one_thread_blas()
    from oracle import rgf_oracle
    h, tnn, Es, eta = args
    t0 = time.perf_counter()
    z = Es + 1j * eta
    gL, _ = rgf_oracle.sancho_rubio(h[0], tnn[0], z, -1)
    gR, _ = rgf_oracle.sancho_rubio(h[-1], tnn[-1], z, 1)
    SL = np.conj(tnn[0]).T @ gL @ tnn[0]
    SR = tnn[-1] @ gR @ np.conj(tnn[-1]).T
    _, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es, SL, SR)
    car = rgf_oracle.caroli_trace(GN0, SL, SR)
    return time.perf_counter() - t0, float(car.sum())


def _cpu_bardeen(args):
    bu.one_thread_blas()
    from oracle import bardeen_oracle as bo
    VRs, NLR = args
    one = np.eye(1)
    HC = _barrier_HC()
    t0 = time.perf_counter()
    for v in VRs:
        bo.kernel_well(one, one, one, 0.5 * one, 0.0 * one, (v + 0.5) * one, v * one, 0.5 * one, 20, NLR, NLR, HC, HC, one,
                       0.1 * one, interval=1e-2)
    return time.perf_counter() - t0, 0.0


def _cpu_dense_fit(args):
    """The reference's own algorithm (dense assembly + dense inverse, Python loops) timed where it can run: one call at
    M = 128 (its O((M n)^2) Python loops dominate there) and the LAPACK inverse alone at two sizes for the cubic term."""
    bu.one_thread_blas()
    from oracle import wfm_oracle
    M0, = args
    hq, tq, t3q = configs.disordered_blocks(N=M0 - 2, n=8, W=0.05, seed=1234)
    t0 = time.perf_counter()
    wfm_oracle.kernel_loops(hq, tq, t3q, 1.0, complex(-0.7), "g_closed", np.eye(8)[0], False, False, all_debug=False)
    t_call = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    t_inv = {}
    for N in (8 * M0, 2048):
        A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        t0 = time.perf_counter()
        np.linalg.inv(A)
        t_inv[N] = time.perf_counter() - t0
    return t_call, t_inv[8 * M0], t_inv[2048]


def _pool_rate(worker, chunks, units):
    """Runs `chunks` on one worker process each; returns (units/s aggregate, wall s, cores)."""
    cores = len(chunks)
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        pool.map(worker, chunks)
        wall = time.perf_counter() - t0
    return units / wall, wall, cores


# ---------------------------------------------------------------------------------------------- cfg1 / cfg3
def _run_small_blocks(name, env, h, tnn, tnnn, Es_fn, n_base, flops_per_point, steps, warmup, scaling, want_cpu, want_e2e,
                      parity_fn, cpu_fn):
    torch = env.torch
    lib = env.lib
    n, M = h.shape[-1], h.shape[0]
    total, lo, hi = _slice(n_base, env, scaling)
    Es = Es_fn(total)[lo:hi]
    nE = len(Es)
    d_h, d_t, d_E = env.dev_c(h), env.dev_c(tnn), env.dev_c(Es)
    R = torch.empty((nE, n, n), dtype=torch.float64, device=env.dev)
    T = torch.empty_like(R)
    res = torch.empty((nE, n), dtype=torch.float64, device=env.dev)
    flag = torch.zeros(1, dtype=torch.int32, device=env.dev)
    st = env.stream.cuda_stream

    def step():
        _lib.check(lib.tx_transmission_sweep_dev(d_h.data_ptr(), 1, None, None, d_t.data_ptr(), 1, n, M, 1,
                                                 d_E.data_ptr(), nE, _lib.TX_LEAD_CLOSED, None, None, 1,
                                                 _lib.TX_HOP_SCALAR, None, None, n, R.data_ptr(), T.data_ptr(),
                                                 res.data_ptr(), None, flag.data_ptr(), st))
    sampler = bu.ClockSampler(env)
    l0 = int(lib.tx_launch_counter())
    ms, wall = env.timed(step, steps, warmup, sampler)
    launches = (int(lib.tx_launch_counter()) - l0) * steps // (steps + max(warmup, 3))
    assert int(flag.item()) == 0, "singular block in the benchmark sweep"
    Th, Rh = T.cpu().numpy(), R.cpu().numpy()
    resid = np.abs(res.cpu().numpy())
    peak = env.fp64_peak()
    hbm, hbm_src = bu.hbm_peak()
    ach = flops_per_point * nE / (ms * 1e-3) / 1e12            # this rank's kernel: nE points in ms (max over ranks)
    bytes_pt = 24 + 16 * n * n
    roof = {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "flops_per_point": flops_per_point, "kernel_ms": ms,
            "peak_source": "tx_measure_fp64_peak (register-resident DFMA loop, measured in this run)",
            "hbm": {"achieved": bytes_pt * nE / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                    "frac": bytes_pt * nE / (ms * 1e-3) / 1e9 / hbm, "bytes_per_point": bytes_pt, "peak_source": hbm_src}}
    out = _std(name, env, scaling, total, ms, {"roofline": roof, "gpu_launches": launches, "clocks": sampler.summary(),
                                               "unitarity_max_abs": float(np.nanmax(resid)),
                                               "unitarity_median_abs": float(np.nanmedian(resid))})
    if env.rank == 0:
        out.update(parity_fn(Es, Rh, Th, lo))
    if name == "cfg1" and env.world == 1:
        # launch-latency bound (one ~12 us step): the same step captured into a CUDA graph and replayed back to back,
        # without the L2 flush between steps (so this is a launch-rate figure next to `value`, not a replacement)
        side = torch.cuda.Stream()
        side.wait_stream(env.stream)
        with torch.cuda.stream(side):
            def step_side():
                _lib.check(lib.tx_transmission_sweep_dev(d_h.data_ptr(), 1, None, None, d_t.data_ptr(), 1, n, M, 1,
                                                         d_E.data_ptr(), nE, _lib.TX_LEAD_CLOSED, None, None, 1,
                                                         _lib.TX_HOP_SCALAR, None, None, n, R.data_ptr(), T.data_ptr(),
                                                         res.data_ptr(), None, flag.data_ptr(), side.cuda_stream))
            step_side()
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                step_side()
            reps = 200
            for _ in range(10):
                graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(reps):
                graph.replay()
            e1.record(side)
            side.synchronize()
            g_ms = e0.elapsed_time(e1) / reps
            e0.record(side)
            for _ in range(reps):
                step_side()
            e1.record(side)
            side.synchronize()
            d_ms = e0.elapsed_time(e1) / reps
        assert np.array_equal(T.cpu().numpy(), Th)
        out["back_to_back"] = {"cuda_graph_replay_ms_per_sweep": g_ms, "direct_calls_ms_per_sweep": d_ms,
                               "cuda_graph_points_per_s": nE / (g_ms * 1e-3), "direct_points_per_s": nE / (d_ms * 1e-3),
                               "note": "200 sweeps issued back to back on one stream, no L2 flush in between (results stay "
                                       "L2-resident): the launch rate of the device entry, direct and as a replayed CUDA graph"}
    if want_e2e:
        from transport_b200.sweep import transmission_sweep
        Rp = torch.empty((1, nE, n, n), dtype=torch.float64, pin_memory=True).numpy()
        Tp = torch.empty((1, nE, n, n), dtype=torch.float64, pin_memory=True).numpy()
        ph, pt, pE = env.pin(h), env.pin(tnn), env.pin(Es)
        transmission_sweep(ph, pt, pE, out=(Rp, Tp))
        env.barrier()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            transmission_sweep(ph, pt, pE, out=(Rp, Tp))
            _ = float(Tp[0, 0, 0, 0])
        env.barrier()
        dt = env.max_over_ranks((time.perf_counter() - t0) / reps)
        out["e2e"] = {"value": total / dt, "unit": bu.UNIT, "h2d_bytes_per_step": int(ph.nbytes + pt.nbytes + pE.nbytes),
                      "d2h_bytes_per_step": int(Rp.nbytes + Tp.nbytes), "ms_per_step": dt * 1e3,
                      "api": "transport_b200.sweep.transmission_sweep -> tx_transmission_sweep_spin (host pointers)"}
        assert np.array_equal(Tp[0], Th), "e2e path and device-resident path disagree"
    if want_cpu and env.rank == 0 and env.world == 1:
        out["cpu_baseline"] = cpu_fn()
    return out


def run_cfg1(env, steps=20, warmup=3, scaling="weak", want_cpu=True, want_e2e=True, small=False):
    h, tnn, tnnn = configs.barrier_blocks(0.4, 11)
    n_base = 10_000
    g = np.load(os.path.join(GOLDEN, "cfg1_barrier.npz"))

    def parity(Es, Rh, Th, lo):
        if env.world > 1:
            return {}
        got_T, got_R = Th[::50, 0], Rh[::50, 0]
        err = max(bu.rel_err(got_T, g["T"][:, 0]), bu.rel_err(got_R, g["R"][:, 0]))
        return {"parity_max_rel_err": err, "parity_against": "tests/golden/cfg1_barrier.npz (unmodified reference, every 50th energy)"}

    def cpu():
        cores = os.cpu_count() or 1
        Es = configs.barrier_energies(n_base)
        chunks = [(h, tnn, tnnn, Es[i::cores], [np.array([1.0])]) for i in range(cores)]
        rate, wall, cores = _pool_rate(_cpu_kernel_loops, chunks, n_base)
        return {"value": rate, "unit": bu.UNIT, "cores": cores, "kind": "port",
                "sample": "all %d energies x 1 source, reference-loop port (dense assembly + inverse), %.1f s wall" % (n_base, wall)}

    return _run_small_blocks("cfg1", env, h, tnn, tnnn, lambda k: configs.barrier_energies(k), n_base, 8 * 1 * 13 * 2, steps, warmup,
                             scaling, want_cpu, want_e2e, parity, cpu)


def run_cfg3(env, steps=3, warmup=3, scaling="weak", want_cpu=True, want_e2e=True, small=False):
    N = 1000 if small else 10_000
    n_base = 10_000 if small else 100_000
    h, tnn, tnnn = configs.disordered_blocks(N=N, n=8, W=0.05, seed=1234)
    hp_path = os.path.join(GOLDEN, "hp_cfg3_M10002.npz")

    def parity(Es, Rh, Th, lo):
        from oracle import rgf_oracle
        out = {}
        if not small and env.world == 1 and os.path.exists(hp_path):
            g = np.load(hp_path)
            idx = g["idx"]
            eT = np.abs(Th[idx] - g["T_hp"]) / np.maximum(np.abs(g["T_hp"]), 1e-13 * np.abs(g["T_hp"]).max(-1, keepdims=True))
            eR = np.abs(Rh[idx] - g["R_hp"]) / np.maximum(np.abs(g["R_hp"]), 1e-13 * np.abs(g["R_hp"]).max(-1, keepdims=True))
            dev = np.maximum(eT.max(axis=(1, 2)), eR.max(axis=(1, 2)))                 # per energy
            f64 = g["f64_err"]                                                        # numpy float64 recursion vs truth
            kap = g["kappa"]
            out.update({"parity_max_rel_err": float(dev.max()),
                        "parity_against": "tests/golden/hp_cfg3_M10002.npz: long-double recursion (eps 1.1e-19) validated "
                                          "by mpmath at 50 digits; %d energies incl. the worst-unitarity ones" % len(idx),
                        "float64_numpy_oracle_max_rel_err_vs_extended_precision": float(f64.max()),
                        "device_over_float64_oracle_error_ratio_median": float(np.median(dev / np.maximum(f64, 1e-300))),
                        "device_over_float64_oracle_error_ratio_max": float(np.max(dev / np.maximum(f64, 1e-300))),
                        "kappa_max": float(kap.max()),
                        "device_err_over_kappa_eps_max": float(np.max(dev / (kap * 2.2e-16))),
                        "float64_oracle_err_over_kappa_eps_max": float(np.max(f64 / (kap * 2.2e-16)))})
        else:
            idx = np.arange(0, len(Es), max(1, len(Es) // 8))
            oR, oT = rgf_oracle.transmission_closed(h, tnn, Es[idx], np.eye(8))
            out.update({"parity_max_rel_err": max(bu.rel_err(Th[idx], oT), bu.rel_err(Rh[idx], oR)),
                        "parity_against": "oracle/rgf_oracle.py (float64 numpy recursion) on %d energies" % len(idx)})
        return out

    def cpu():
        cores = os.cpu_count() or 1
        # (a) the numpy restatement of the recursive algorithm at full length, one energy batch per core
        per = 2
        Es = np.linspace(-1.9, 1.9, cores * per).astype(complex)
        chunks = [(h, tnn, Es[i::cores], np.eye(8)) for i in range(cores)]
        rgf_rate, rgf_wall, _ = _pool_rate(_cpu_rgf, chunks, cores * per)
        # (b) the reference's own algorithm (dense assembly + dense inverse, Python loops) at M in {64, 128, 256},
        #     one source, fitted t(M) = a M^2 + b M^3 and extrapolated to M = 10 002; the
        #     reference itself cannot run there: 102 GB dense matrix
        #     (timed in a worker process like every other CPU leg: BLAS in the parent is not fork-safe once pools ran)
        M0 = 128
        with mp.get_context("fork").Pool(1) as pool:
            t_call, t_inv_small, t_inv_2048 = pool.map(_cpu_dense_fit, [(M0,)])[0]
        a = max(t_call - t_inv_small, 0.0) / M0 ** 2                               # Python loops: O((M n)^2)
        b = t_inv_2048 / 2048.0 ** 3 * 8 ** 3                                      # LAPACK inverse: O((M n)^3), per M^3
        t_full = (a * (N + 2) ** 2 + b * (N + 2) ** 3) * 8                         # x 8 sources per point
        Ms, ts = [M0], [t_call]
        return {"value": cores / t_full, "unit": bu.UNIT, "cores": cores, "kind": "port",
                "sample": "EXTRAPOLATED: reference-loop port timed at M=%s (%s s per call: Python loops a M^2) + LAPACK inverse of a "
                          "2048^2 block (%.2f s: b M^3) -> %.3g s per point (8 sources) at M=%d, x %d cores; the dense path cannot "
                          "run at this size (102 GB)" % (Ms, ["%.2f" % t for t in ts], t_inv_2048, t_full, N + 2, cores),
                "rgf_numpy_restatement": {"value": rgf_rate, "unit": bu.UNIT, "cores": cores,
                                          "sample": "%d energies x 8 sources at full length, %.1f s wall" % (cores * per, rgf_wall)}}

    return _run_small_blocks("cfg3", env, h, tnn, tnnn, lambda k: np.linspace(-1.9, 1.9, k).astype(complex), n_base,
                             8 * 8 ** 3 * (N + 2) * 2, steps, warmup, scaling, want_cpu, want_e2e, parity, cpu)


# ---------------------------------------------------------------------------------------------- cfg4
def run_cfg4(env, steps=2, warmup=3, scaling="weak", want_cpu=True, want_e2e=True, small=False, width=None, n_energies=None,
             L=32):
    """`width`, `n_energies`, `L`: the same workload at another ribbon width / grid (scripts/bench_wide_blocks.py)."""
    torch = env.torch
    lib = env.lib
    n = width or (32 if small else 64)
    n_base = n_energies or (256 if small else 4096)
    h, tnn, _ = configs.ribbon_blocks(width=n, L=L, W=1.0, seed=1234)
    total, lo, hi = _slice(n_base, env, scaling)
    Es = np.linspace(-3.5, 3.5, total).astype(complex)[lo:hi]
    nE = len(Es)
    eta = 1e-8
    M = h.shape[0]
    d_h, d_t, d_E = env.dev_c(h), env.dev_c(tnn), env.dev_c(Es)
    sigL = torch.empty((nE, n, n, 2), dtype=torch.float64, device=env.dev)
    sigR = torch.empty_like(sigL)
    nit = torch.zeros((2, nE), dtype=torch.int32, device=env.dev)
    ok = torch.zeros((2, nE), dtype=torch.int32, device=env.dev)
    conv = torch.zeros((2, nE), dtype=torch.float64, device=env.dev)
    car = torch.empty(nE, dtype=torch.float64, device=env.dev)
    flag = torch.zeros(1, dtype=torch.int32, device=env.dev)
    wse = int(lib.tx_sweep_workspace_elems(n, nE))
    work = torch.empty((max(wse, 1), 2), dtype=torch.float64, device=env.dev)
    st = env.stream.cuda_stream
    nn16 = n * n * 16

    if os.environ.get("TX_BENCH_HERM") is not None:           # A/B of the Hermitian-hopping decimation (scripts/ab_cfg4_leads.py)
        _lib.check(lib.tx_set_option(6, int(os.environ["TX_BENCH_HERM"])))
    # both leads are the same clean ribbon: one Sancho-Rubio decimation carries both surfaces (tx_surface_gf_pair_dev);
    # TX_BENCH_TWO_DECIMATIONS=1 times the two independent decimations instead (A/B)
    same_leads = (np.array_equal(h[0], h[-1]) and np.array_equal(tnn[0], tnn[-1])
                  and not os.environ.get("TX_BENCH_TWO_DECIMATIONS"))

    def leads_step():
        if same_leads:
            _lib.check(lib.tx_surface_gf_pair_dev(d_h.data_ptr(), d_t.data_ptr(), n, d_E.data_ptr(), nE, eta, 1e-14, 200,
                                                  None, sigL.data_ptr(), None, sigR.data_ptr(), nit[0].data_ptr(),
                                                  conv[0].data_ptr(), ok[0].data_ptr(), flag.data_ptr(), st))
            return
        _lib.check(lib.tx_surface_gf_dev(d_h.data_ptr(), d_t.data_ptr(), n, d_E.data_ptr(), nE, -1, _lib.TX_GF_SANCHO,
                                         eta, 1e-14, 0, 0, 200, None, sigL.data_ptr(), nit[0].data_ptr(),
                                         conv[0].data_ptr(), ok[0].data_ptr(), flag.data_ptr(), st))
        _lib.check(lib.tx_surface_gf_dev(d_h.data_ptr() + (M - 1) * nn16, d_t.data_ptr() + (M - 2) * nn16, n,
                                         d_E.data_ptr(), nE, 1, _lib.TX_GF_SANCHO, eta, 1e-14, 0, 0, 200, None,
                                         sigR.data_ptr(), nit[1].data_ptr(), conv[1].data_ptr(), ok[1].data_ptr(),
                                         flag.data_ptr(), st))

    def sweep_step():
        _lib.check(lib.tx_transmission_sweep_large_dev(d_h.data_ptr(), 1, None, None, d_t.data_ptr(), 1, n, M, 1,
                                                       d_E.data_ptr(), nE, sigL.data_ptr(), sigR.data_ptr(), 1,
                                                       _lib.TX_HOP_SCALAR, None, n, None, None, None, None,
                                                       car.data_ptr(), work.data_ptr() if wse else None,
                                                       flag.data_ptr(), st))

    def step():
        leads_step()
        sweep_step()

    sampler = bu.ClockSampler(env)
    l0 = int(lib.tx_launch_counter())
    ms, wall = env.timed(step, steps, warmup, sampler)
    launches = (int(lib.tx_launch_counter()) - l0) * steps // (steps + max(warmup, 3))
    ms_leads, _ = env.timed(leads_step, max(1, steps // 2), 3)
    ms_sweep, _ = env.timed(sweep_step, max(1, steps // 2), 3)
    if same_leads:
        nit[1], ok[1], conv[1] = nit[0], ok[0], conv[0]
    assert bool(ok.all().item()), "Sancho-Rubio did not converge everywhere"
    assert int(flag.item()) == 0
    n_it = float(nit.double().mean().item())
    carh = car.cpu().numpy()
    flops_sweep = 8 * n ** 3 * M * 2
    flops_leads = 2 * n_it * (8 + 6 * 8) * n ** 3
    flops = flops_sweep + flops_leads
    herm_hop = bool(np.array_equal(tnn[0], np.conj(tnn[0]).T) and os.environ.get("TX_BENCH_HERM", "1") != "0")
    flops_exec = flops_sweep + (1 if same_leads else 2) * n_it * (8 + (2 if herm_hop else 6) * 8) * n ** 3
    peak = env.fp64_peak()
    ach = flops * nE / (ms * 1e-3) / 1e12
    roof = {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "flops_per_point": flops, "kernel_ms": ms, "sancho_iterations_mean": n_it,
            "lead_decimations_per_energy": 1 if same_leads else 2,
            "hermitian_hopping_decimation": herm_hop,
            "executed": {"flops_per_point": flops_exec, "frac": flops_exec * nE / (ms * 1e-3) / 1e12 / peak,
                         "note": ("identical leads: one Sancho-Rubio decimation yields both surface blocks (the other one is "
                                  "bulk + h00 - surface); " if same_leads else "two independent decimations; ") +
                                 ("Hermitian hopping block: alpha = beta throughout, 1 inverse + 2 products per iteration "
                                  "instead of 1 + 6" if herm_hop else "general hopping: 1 inverse + 6 products per iteration")},
            "kernels": {"surface_gf_sancho (both leads)": {"ms": ms_leads, "flops_per_point": flops_leads,
                                                           "frac": flops_leads * nE / (ms_leads * 1e-3) / 1e12 / peak},
                        "rgf_sweep_generic": {"ms": ms_sweep, "flops_per_point": flops_sweep,
                                              "frac": flops_sweep * nE / (ms_sweep * 1e-3) / 1e12 / peak}},
            "peak_source": "tx_measure_fp64_peak (register-resident DFMA loop, measured in this run; the DMMA path shares the pipe)"}
    out = _std("cfg4", env, scaling, total, ms, {"roofline": roof, "gpu_launches": launches, "clocks": sampler.summary()})
    if env.rank == 0:
        from oracle import rgf_oracle
        idx = np.arange(0, nE, max(1, nE // 6))
        SL = torch.view_as_complex(sigL[idx]).cpu().numpy()
        SR = torch.view_as_complex(sigR[idx]).cpu().numpy()
        _, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es[idx], SL, SR)
        want = rgf_oracle.caroli_trace(GN0, SL, SR)
        out["parity_max_rel_err"] = float(np.max(np.abs(carh[idx] - want) / np.maximum(np.abs(want), 1e-12)))
        out["parity_against"] = ("oracle/rgf_oracle.py Caroli trace on the device's own Sancho-Rubio tables, %d energies "
                                 "(parity of multi-channel leads is unpinned by the reference: g_iter refuses n>2)" % len(idx))
        hp_path = os.path.join(GOLDEN, "hp_cfg4_leads.npz")
        if not small and width is None and n_energies is None and env.world == 1 and os.path.exists(hp_path):
            g = np.load(hp_path)
            li = g["idx"]
            SLd = torch.view_as_complex(sigL[li]).cpu().numpy()
            den = np.abs(g["sigL_hp"]).max(axis=(1, 2), keepdims=True)
            out["leads_max_err_vs_extended_precision"] = float(np.max(np.abs(SLd - g["sigL_hp"]) / den))
            out["leads_float64_numpy_err_vs_extended_precision"] = float(g["f64_err"].max())
    if want_e2e:
        from transport_b200.sweep import transmission_sweep
        ph, pt, pE = env.pin(h), env.pin(tnn), env.pin(Es)
        car2 = transmission_sweep(ph, pt, pE, converger=("sancho", eta), want_rt=False, want_caroli=True)
        env.barrier()
        reps = 2
        t0 = time.perf_counter()
        for _ in range(reps):
            car2 = transmission_sweep(ph, pt, pE, converger=("sancho", eta), want_rt=False, want_caroli=True)
            _ = float(car2[0, 0])
        env.barrier()
        dt = env.max_over_ranks((time.perf_counter() - t0) / reps)
        assert np.array_equal(car2[0], carh) if same_leads else np.allclose(car2[0], carh, rtol=1e-6), "e2e path and device-resident path disagree"
        out["e2e"] = {"value": total / dt, "unit": bu.UNIT, "h2d_bytes_per_step": int(ph.nbytes + pt.nbytes + pE.nbytes),
                      "d2h_bytes_per_step": int(car2.nbytes), "ms_per_step": dt * 1e3,
                      "api": "transport_b200.sweep.transmission_sweep(converger=('sancho', eta), want_caroli) -> "
                             "tx_transmission_sweep_spin (host pointers; self-energy tables stay on the device)"}
    if want_cpu and env.rank == 0 and env.world == 1:
        cores = os.cpu_count() or 1
        per = 2
        Ec = np.linspace(-3.5, 3.5, cores * per).astype(complex)
        chunks = [(h, tnn, Ec[i::cores], eta) for i in range(cores)]
        rate, wall_c, cores = _pool_rate(_cpu_ribbon, chunks, cores * per)
        out["cpu_baseline"] = {"value": rate, "unit": bu.UNIT, "cores": cores, "kind": "port",
                               "sample": "numpy restatement (Sancho-Rubio + recursive sweep + Caroli trace; the reference refuses "
                                         "n>2 iterative leads, wfm/__init__.py:502) on %d energies, %.1f s wall" % (cores * per, wall_c)}
    return out


# ---------------------------------------------------------------------------------------------- cfg5
def _barrier_HC():
    HC = np.zeros((11, 11, 1, 1), dtype=complex)
    for j in range(11):
        HC[j, j] = 0.4
    for j in range(10):
        HC[j, j + 1] = HC[j + 1, j] = -1.0
    return HC


def run_cfg5(env, steps=5, warmup=3, scaling="weak", want_cpu=True, want_e2e=True, small=False):
    """rbstep.py-style sweep of the right-well step height (rbstep.py:43-104): P geometries per device pass — bound
    states of both wells, window-averaged matrix elements, current over the bias grid of rbbarrier.py:447-450.
    The unit of `value` is geometries/s; the two sizes N_LR = 200 and 800 are reported side by side and `value` is
    the S = 451 one (rbstep.py's own size)."""
    torch = env.torch
    lib = env.lib
    from transport_b200 import bardeen
    one = np.eye(1)
    sizes = (200,) if small else (200, 800)
    P_base = 64 if small else 256
    per_size = []
    for NLR in sizes:
        total, lo, hi = _slice(P_base, env, scaling)
        VRs = np.linspace(-0.05, 0.05, total)[lo:hi]
        P = len(VRs)
        Vinf, VL = 0.5 * one, 0.0 * one
        HC = _barrier_HC()
        interval = 1e-2
        t0 = time.perf_counter()
        A = bardeen.step_sweep_arrays(one, one, one, Vinf, VL, VRs, 20, NLR, NLR, HC)
        host_assembly_s = time.perf_counter() - t0
        S = A["dL"].shape[-1]
        cut = np.full((P, 1), 0.1 - 2.0)
        r = bardeen.wells_batched(A["dL"], A["eL"], A["dR"], A["eR"], cut, cut + interval, cut, A["h0"], A["hU"], A["hL"],
                                  interval, sign_fix=True)
        ms_states = int(max(r["nL"].max(), r["nR_states"].max()))
        D = {k: env.dev_f(v) for k, v in A.items()}
        dcut, dcut2 = env.dev_f(cut), env.dev_f(cut + interval)
        EL = torch.zeros((P, 1, ms_states), dtype=torch.float64, device=env.dev)
        ER = torch.zeros_like(EL)
        Mavg = torch.zeros((P, 1, ms_states, 1), dtype=torch.float64, device=env.dev)
        nL = torch.zeros((P, 1), dtype=torch.int32, device=env.dev)
        nRs, nRb = torch.zeros_like(nL), torch.zeros_like(nL)
        ws = torch.empty(int(lib.tx_bardeen_workspace_bytes(P, 1, S, ms_states)), dtype=torch.uint8, device=env.dev)
        Vb = torch.linspace(-0.05, 0.05, 99, dtype=torch.float64, device=env.dev)
        Iab = torch.zeros((P, 1, 1, 99), dtype=torch.float64, device=env.dev)
        M2 = torch.zeros_like(Mavg)
        st = env.stream.cuda_stream

        def step():
            _lib.check(lib.tx_bardeen_wells_dev(D["dL"].data_ptr(), D["eL"].data_ptr(), D["dR"].data_ptr(), D["eR"].data_ptr(),
                                                dcut.data_ptr(), dcut2.data_ptr(), dcut.data_ptr(), D["h0"].data_ptr(),
                                                D["hU"].data_ptr(), D["hL"].data_ptr(), P, 1, S, ms_states, interval, 1,
                                                EL.data_ptr(), nL.data_ptr(), ER.data_ptr(), nRs.data_ptr(), nRb.data_ptr(),
                                                Mavg.data_ptr(), ws.data_ptr(), st))
            torch.mul(Mavg, Mavg, out=M2)
            _lib.check(lib.tx_bardeen_current_dev(EL.data_ptr(), M2.data_ptr(), P, 1, ms_states, Vb.data_ptr(), 99, -1.95, 0.01,
                                                  Iab.data_ptr(), st))
        sampler = bu.ClockSampler(env)
        l0 = int(lib.tx_launch_counter())
        ms, wall = env.timed(step, steps, warmup, sampler)
        launches = (int(lib.tx_launch_counter()) - l0) * steps // (steps + max(warmup, 3))
        assert np.array_equal(nL.cpu().numpy(), r["nL"]) and np.allclose(Mavg.cpu().numpy(), r["Mavg"], rtol=1e-12, atol=0)
        states = int(r["nL"].sum() + r["nR_states"].sum())
        # executed work of the eigenvalue stage: states x Sturm counts x S sites x 3 FP64 instructions
        item = {"N_LR": NLR, "S": S, "geometries_per_gpu": P, "bound_states": states, "max_states": ms_states,
                "ms_per_step": ms, "geometries_per_s": total / (ms * 1e-3), "gpu_launches": launches,
                "clocks": sampler.summary(), "host_assembly_s_python": host_assembly_s,
                "roofline": {"bound": "fp64 (latency: chains of S dependent FMAs on shared-memory resident data)",
                             "achieved": states * 53 * S * 3 * 2 / (ms * 1e-3) / 1e12, "peak": env.fp64_peak(), "unit": "TFLOP/s",
                             "frac": states * 53 * S * 3 * 2 / (ms * 1e-3) / 1e12 / env.fp64_peak(), "traffic": None,
                             "note": "algorithmic work = 53 bisection counts x S sites x 3 FP64 instructions per eigenvalue "
                                     "(DESIGN.md §3); the kernels are parallelism-bound, not pipe-bound"}}
        if env.rank == 0:
            from oracle import bardeen_oracle as bo
            idx = [0, P // 2, P - 1] if NLR == 200 else [P // 3]
            err, cpu_s = 0.0, 0.0
            for p in idx:
                VR = VRs[p] * one
                t0 = time.perf_counter()
                E0, M0 = bo.kernel_well(one, one, one, Vinf, VL, VR + Vinf, VR, Vinf, 20, NLR, NLR, HC, HC, one, 0.1 * one,
                                        interval=interval)
                cpu_s += time.perf_counter() - t0
                mu, nu = int(r["nL"][p, 0]), int(r["nR_below"][p, 0])
                Msq = r["Mavg"][p, 0, :mu, 0] ** 2
                Msq[nu:] = 0.0
                big = M0[0, :, 0] > 1e-30
                err = max(err, float(np.max(np.abs(r["EL"][p, 0, :mu] - E0[0].real))),
                          float(np.max(np.abs(Msq[big] - M0[0, big, 0]) / M0[0, big, 0])) if big.any() else 0.0)
            item["parity_max_err"] = err
            item["parity_against"] = "oracle/bardeen_oracle.kernel_well (dense eigh as the reference; pinned to reference goldens) on %d geometries: E absolute, |M|^2 relative" % len(idx)
            item["cpu_oracle_s_per_geometry_1core_blas_default"] = cpu_s / len(idx)
        if want_e2e:
            reps = 5
            bardeen.step_sweep(one, one, one, Vinf, VL, VRs, 20, NLR, NLR, HC, E_cutoff=0.1, interval=interval,
                               Vb=np.linspace(-0.05, 0.05, 99), muR=-1.95, kBT=0.01)      # warm-up: workspace + buffer capacity
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                rr = bardeen.step_sweep(one, one, one, Vinf, VL, VRs, 20, NLR, NLR, HC, E_cutoff=0.1, interval=interval,
                                        Vb=np.linspace(-0.05, 0.05, 99), muR=-1.95, kBT=0.01)
                _ = float(rr["I"][0, 0, 0, 0])
            env.barrier()
            dt = env.max_over_ranks((time.perf_counter() - t0) / reps)
            assert np.allclose(rr["I"], Iab.cpu().numpy(), rtol=1e-12, atol=1e-300)
            item["e2e"] = {"value": total / dt, "unit": "geometries/s", "ms_per_step": dt * 1e3,
                           "h2d_bytes_per_step": int(rr["h2d_bytes"]), "d2h_bytes_per_step": int(rr["d2h_bytes"]),
                           "api": "transport_b200.bardeen.step_sweep (geometry parameters in, states / matrix elements / current out)"}
        if want_cpu and env.rank == 0 and env.world == 1:
            cores = min(os.cpu_count() or 1, 8 if NLR == 800 else 32)
            chunks = [([VRs[(i * 7) % P]], NLR) for i in range(cores)]
            rate, wall_c, cores = _pool_rate(_cpu_bardeen, chunks, cores)
            item["cpu_baseline"] = {"value": rate, "unit": "geometries/s", "cores": cores, "kind": "port",
                                    "sample": "%d geometries (one per core), oracle kernel_well = the reference's dense-eigh algorithm, %.1f s wall" % (cores, wall_c)}
        per_size.append(item)
    head = per_size[0]
    out = {"workload": WORKLOADS["cfg5"], "value": head["geometries_per_s"], "unit": "geometries/s", "ms_per_step": head["ms_per_step"],
           "n_gpus": env.world, "scaling": scaling, "units_per_step_all_gpus": int(P_base * (env.world if scaling == "weak" else 1)),
           "roofline": head["roofline"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "sizes": per_size}
    for k in ("e2e", "cpu_baseline", "parity_max_err", "parity_against"):
        if k in head:
            out[k if k != "parity_max_err" else "parity_max_rel_err"] = head[k]
    return out


RUNNERS = {"cfg1": run_cfg1, "cfg3": run_cfg3, "cfg4": run_cfg4, "cfg5": run_cfg5}


def run(name, env, **kw):
    return RUNNERS[name](env, **kw)


if __name__ == "__main__":
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("TX_BENCH_WATCHDOG", "1200")), exit=True)
    small = "--small" in sys.argv
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg1", "cfg3", "cfg4", "cfg5"]
    env = bu.Env()
    env.init_dist()
    for w in which:
        res = run(w, env, small=small)
        if env.rank == 0:
            print(json.dumps(res), flush=True)
