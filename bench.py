#!/usr/bin/env python
"""Benchmark of the T(E) hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config cfg1|cfg2|cfg3|cfg4|cfg5] [--scaling weak|strong] [--no-secondary] [--no-cpu]

Headline (default, BASELINE.json configs[1]): two spin-1/2 impurities, n_loc_dof = 8, M = 8 blocks,
H(J) = H0 + J*H1, J = linspace(-1,-0.001,1000) x E = logspace(-4,0,1000) - 2 -> 10^6 (energy, parameter) points PER GPU
(weak scaling: rank r takes its own 1000-value J slice of a 1000*N-value grid; `--scaling strong` shards ONE 10^6-point
grid), all 8 one-hot sources per point.  A step = one pass of the sweep over the rank's points with inputs resident in
HBM; on N > 1 GPUs the per-point totals are gathered INSIDE the sweep: the K3 epilogue stores them into every rank's
result buffer over NVLink (CUDA-IPC peer memory), and a device-side flag barrier closes the step — no collective
follows the kernel (`--gather nccl` runs the round-1 NCCL all-gather instead).

One JSON line is printed by rank 0: metric value, e2e through the public host-buffer API (full per-channel R/T as the
reference returns them) plus the compact spin-resolved path, roofline of the dominant kernel, the CPU baseline (the
oracle's loop-for-loop port of the reference, all host cores), clocks sampled during the timed region, and (N = 1) a
`secondary` block with BASELINE configs 1, 3, 4 and 5 under the same rules (scripts/bench_configs.py).  `--config cfgK`
makes any of them the headline line instead, at any N.
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)

import benchutil as bu  # noqa: E402

N_J, N_E, NLOC, MBLK = 1000, 1000, 8, 8
FLOPS_PER_POINT = 8 * NLOC ** 3 * MBLK * 2          # 8 n^3 M (1+g), g=1 scalar hopping  
# What the telescoped kernel really executes per point on this geometry: 2 dense recurrence
# sites + the g_0 product (3 x 8 n^3 on DMMA), 1 inverse (8 n^3), 5 diagonal-site column scalings and the
# chunk start (6 x 8 n^2), prologue/epilogue O(n^2).  The algorithmic figure above counts one inverse and
# one product per site, which the telescoping removes; both fractions are reported.
EXECUTED_FLOPS_PER_POINT = 4 * 8 * NLOC ** 3 + 6 * 8 * NLOC ** 2 + 40 * NLOC ** 2
BYTES_PER_POINT = 24 + 16 * NLOC * NLOC             # unique bytes: E, J, R and T for 8 sources
METRIC, UNIT = bu.METRIC, bu.UNIT


def config_dict(name, world, scaling):
    """The `config` object of the JSON line: identical keys and values in both arms (ours / reference)."""
    from bench_configs import WORKLOADS
    return {"workload": WORKLOADS[name], "config": name, "scaling": scaling, "parallelism": "dp%d" % world,
            "dtype": "complex128 arithmetic, float64 outputs",
            "l2": "GPU arm: 256 MB L2 flush between timed steps, outside the event pair (CPU arm: not applicable)"}


# ------------------------------------------------------------------------------------------ CPU arm (cfg2)
def _cpu_worker(args):
    """Times the oracle's loop-for-loop port of wfm.kernel on a slice of (J, E) points."""
    bu.one_thread_blas()
    from oracle import wfm_oracle
    from transport_b200 import configs
    pts, = args
    t0 = time.perf_counter()
    acc = 0.0
    for (J, E) in pts:
        h, tnn, tnnn = configs.cicc_blocks(1, 1.0, J, 0.0, 0.0, 0.0, 3, 2)
        for a in range(NLOC):
            src = np.zeros(NLOC)
            src[a] = 1.0
            Rs, Ts = wfm_oracle.kernel_loops(h, tnn, tnnn, 1.0, E, "g_closed", src, False, False, all_debug=False)
            acc += Ts.sum()
    return time.perf_counter() - t0, acc


def cpu_sample_points(n_points, offset=0):
    from transport_b200 import configs
    Js, Es = configs.cfg2_grid(N_J, N_E)
    rng = np.random.default_rng(1234 + offset)
    ji = rng.integers(0, N_J, n_points)
    ei = rng.integers(0, N_E, n_points)
    return [(float(Js[j]), complex(Es[e])) for j, e in zip(ji, ei)]


def run_cpu_sample(pool, cores, pts_per_core, offset=0):
    """One bounded sample: every core evaluates `pts_per_core` points x 8 sources.  Returns pts/s."""
    pts = cpu_sample_points(cores * pts_per_core, offset)
    chunks = [(pts[i::cores],) for i in range(cores)]
    t0 = time.perf_counter()
    pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    return len(pts) / wall, wall


def reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port: Python loops + dense inverse,
    transport/wfm/__init__.py:29-215; oracle/bardeen_oracle.py for cfg5) on all host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import bench_configs as bc
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    unit = UNIT
    if args.config == "cfg2":
        pts_per_core = 12
        with mp.get_context("fork").Pool(cores) as pool:
            for w in range(args.warmup):
                run_cpu_sample(pool, cores, 2, offset=100 + w)
            t_tot, n_tot = 0.0, 0
            for k in range(args.steps):
                rate, wall = run_cpu_sample(pool, cores, pts_per_core, offset=k)
                t_tot += wall
                n_tot += cores * pts_per_core
        value = n_tot / t_tot
        ms = 1e3 * t_tot / args.steps
        sample = "%d random (J,E) points of the cfg2 grid x 8 sources per step, %d steps" % (cores * pts_per_core, args.steps)
        kind = "port"
    else:
        # bounded samples of the other configs, sized by scripts/bench_configs.py's own CPU legs
        from transport_b200 import configs
        steps = max(1, min(args.steps, 3))
        t_tot, n_tot, sample = 0.0, 0, ""
        kind = "port"
        for k in range(steps):
            if args.config == "cfg1":
                h, tnn, tnnn = configs.barrier_blocks(0.4, 11)
                Es = configs.barrier_energies(10_000)[k::steps]
                chunks = [(h, tnn, tnnn, Es[i::cores], [np.array([1.0])]) for i in range(cores)]
                rate, wall, _ = bc._pool_rate(bc._cpu_kernel_loops, chunks, len(Es))
                n_tot += len(Es)
                sample = "every %d-th of the 10^4 energies per step, reference-loop port" % steps
            elif args.config == "cfg3":
                h, tnn, _ = configs.disordered_blocks(N=10_000, n=8, W=0.05, seed=1234)
                Es = np.linspace(-1.9, 1.9, cores).astype(complex)
                chunks = [(h, tnn, Es[i::cores], np.eye(8)) for i in range(cores)]
                rate, wall, _ = bc._pool_rate(bc._cpu_rgf, chunks, cores)
                n_tot += cores
                sample = ("%d energies x 8 sources at full length per step with the numpy restatement of the RECURSIVE algorithm "
                          "(oracle/rgf_oracle.py): the reference's dense path cannot run M = 10 002 (102 GB)" % cores)
            elif args.config == "cfg4":
                h, tnn, _ = configs.ribbon_blocks(width=64, L=32, W=1.0, seed=1234)
                Es = np.linspace(-3.5, 3.5, 2 * cores).astype(complex)
                chunks = [(h, tnn, Es[i::cores], 1e-8) for i in range(cores)]
                rate, wall, _ = bc._pool_rate(bc._cpu_ribbon, chunks, 2 * cores)
                n_tot += 2 * cores
                sample = "%d energies per step, numpy restatement (Sancho-Rubio + recursive sweep; the reference refuses n>2 leads)" % (2 * cores)
            else:
                unit = "geometries/s"
                VRs = np.linspace(-0.05, 0.05, 256)
                use = min(cores, 32)
                chunks = [([VRs[(7 * i + k) % 256]], 200) for i in range(use)]
                rate, wall, _ = bc._pool_rate(bc._cpu_bardeen, chunks, use)
                n_tot += use
                sample = "%d geometries per step (N_LR = 200), oracle kernel_well = the reference's dense-eigh algorithm" % use
            t_tot += wall
        value = n_tot / t_tot
        ms = 1e3 * t_tot / steps
    line = {
        "impl": "reference", "metric": METRIC if unit == UNIT else "bardeen geometries/sec", "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "complex128 (f64)",
        "data": "synthetic", "config": config_dict(args.config, world, args.scaling),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ peer gather
class PeerGather:
    """Per-point totals gathered inside the sweep: every rank's [world * n_pts] result buffer is cudaMalloc'ed by the
    library, exported with CUDA IPC and mapped by the peers; the sweep epilogue stores into all of them."""

    def __init__(self, env, n_pts):
        from transport_b200 import _lib
        import torch
        self.env, self.lib, self._lib = env, env.lib, _lib
        dist = env.dist
        W = env.world
        self.n_pts = n_pts
        self.buf = ctypes.c_void_p()
        self.flags = ctypes.c_void_p()
        _lib.check(self.lib.tx_peer_alloc(8 * n_pts * W, ctypes.byref(self.buf)))
        _lib.check(self.lib.tx_peer_alloc(64, ctypes.byref(self.flags)))
        hb, hf = (ctypes.c_ubyte * 64)(), (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.tx_peer_export(self.buf, hb))
        _lib.check(self.lib.tx_peer_export(self.flags, hf))
        mine = torch.tensor(list(bytes(hb)) + list(bytes(hf)), dtype=torch.uint8, device=env.dev)
        allh = torch.empty(W * 128, dtype=torch.uint8, device=env.dev)
        dist.all_gather_into_tensor(allh, mine)
        allh = allh.cpu().numpy().reshape(W, 128)
        self.bufs, self.flagp, self.opened = [], [], []
        for r in range(W):
            if r == env.rank:
                self.bufs.append(self.buf.value)
                self.flagp.append(self.flags.value)
                continue
            pb, pf = ctypes.c_void_p(), ctypes.c_void_p()
            _lib.check(self.lib.tx_peer_open((ctypes.c_ubyte * 64)(*allh[r, :64].tolist()), ctypes.byref(pb)))
            _lib.check(self.lib.tx_peer_open((ctypes.c_ubyte * 64)(*allh[r, 64:].tolist()), ctypes.byref(pf)))
            self.bufs.append(pb.value)
            self.flagp.append(pf.value)
            self.opened += [pb, pf]
        self.ctx = ctypes.c_void_p()
        _lib.check(self.lib.tx_context_create(env.local_rank, ctypes.byref(self.ctx)))
        arr = (ctypes.c_void_p * W)(*self.bufs)
        _lib.check(self.lib.tx_context_set_gather(self.ctx, arr, W, env.rank * n_pts))
        self.flag_arr = (ctypes.c_void_p * W)(*self.flagp)
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=env.dev)
        self.epoch = 0
        dist.barrier()

    def barrier_dev(self, stream):
        self.epoch += 1
        self._lib.check(self.lib.tx_peer_barrier_dev(self.flag_arr, self.env.world, self.env.rank, self.epoch,
                                                     self.timed_out.data_ptr(), stream))

    def gathered(self):
        """The rank's gathered array as a torch tensor view (copied out)."""
        torch = self.env.torch
        out = torch.empty(self.n_pts * self.env.world, dtype=torch.float64, device=self.env.dev)
        # wrap the raw device pointer with the CUDA array interface
        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": (self.n_pts * self.env.world,), "typestr": "<f8",
                                        "data": (self.buf.value, False), "version": 2}
        out.copy_(torch.as_tensor(raw, device=self.env.dev))
        return out

    def close(self):
        self.env.torch.cuda.synchronize()
        self.env.dist.barrier()
        for p in self.opened:
            self.lib.tx_peer_close(p)
        self.env.dist.barrier()
        self.lib.tx_context_destroy(self.ctx)
        self.lib.tx_peer_free(self.buf)
        self.lib.tx_peer_free(self.flags)


# ------------------------------------------------------------------------------------------ GPU arm, cfg2
def run_cfg2(env, args, scaling):
    import torch
    from transport_b200 import _lib, configs, shard
    from transport_b200.sweep import transmission_sweep
    lib, dev, world, rank = env.lib, env.dev, env.world, env.rank
    dist = env.dist

    # ---- inputs: this rank's J slice, resident in HBM
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    n_j_total = N_J * world if scaling == "weak" else N_J
    Js_all = np.linspace(-1.0, -0.001, n_j_total)
    (j0, j1), _ = shard.plan(n_j_total, N_E, world, rank)      # parameter axis sharded, energies whole
    Js = np.ascontiguousarray(Js_all[j0:j1])
    n_j = len(Js)
    n_j_max = -(-n_j_total // world)
    _, Es = configs.cfg2_grid(N_J, N_E)
    n_pts = n_j * N_E
    n_pts_total = n_j_total * N_E

    d_h0, d_h1, d_tnn, d_E = env.dev_c(h0), env.dev_c(h1), env.dev_c(tnn), env.dev_c(Es)
    d_J = env.dev_f(Js)
    d_R = torch.empty((n_pts, NLOC, NLOC), dtype=torch.float64, device=dev)
    d_T = torch.empty_like(d_R)
    d_tot = torch.empty(n_pts, dtype=torch.float64, device=dev)
    d_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = env.stream
    gather_mode = "none"
    pg = None
    d_gather = None
    if world > 1:
        gather_mode = args.gather
        if gather_mode == "peer":
            try:
                pg = PeerGather(env, n_j_max * N_E)
            except Exception as exc:      # both gathers run on the device; the choice is recorded in the line
                sys.stderr.write("peer-memory gather unavailable (%s): using the NCCL all-gather\n" % exc)
                gather_mode = "nccl"
        if gather_mode == "nccl":
            d_gather = torch.empty(n_j_max * N_E * world, dtype=torch.float64, device=dev)
            d_pad = torch.zeros(n_j_max * N_E, dtype=torch.float64, device=dev)

    def sweep(ctx=None, R=d_R, T=d_T):
        rc = lib.tx_transmission_sweep_spin_dev(
            ctx, d_h0.data_ptr(), 1, d_h1.data_ptr(), d_J.data_ptr(), d_tnn.data_ptr(), 1, NLOC, MBLK, n_j,
            d_E.data_ptr(), N_E, _lib.TX_LEAD_CLOSED, None, None, 1, _lib.TX_HOP_SCALAR,
            None, None, NLOC, R.data_ptr(), T.data_ptr(), None, d_tot.data_ptr(), 0, 0, None,
            d_flag.data_ptr(), stream.cuda_stream)
        _lib.check(rc)

    def step():
        if pg is not None:
            sweep(pg.ctx)
            pg.barrier_dev(stream.cuda_stream)
        else:
            sweep()
            if gather_mode == "nccl":
                d_pad[:n_pts] = d_tot
                dist.all_gather_into_tensor(d_gather, d_pad)

    sampler = bu.ClockSampler(env)
    l0 = int(lib.tx_launch_counter())
    ms_per_step, t_wall = env.timed(step, args.steps, args.warmup, sampler if rank == 0 else None)
    launches = (int(lib.tx_launch_counter()) - l0) * args.steps // (args.steps + max(args.warmup, 3))
    assert int(d_flag.item()) == 0, "singular block in the benchmark sweep"
    clocks = sampler.summary() if rank == 0 else None
    value = n_pts_total / (ms_per_step * 1e-3)

    gather_ok = None
    if world > 1:
        # the gathered array must equal an NCCL all-gather of the ranks' totals, bit for bit
        ref_pad = torch.zeros(n_j_max * N_E, dtype=torch.float64, device=dev)
        ref_pad[:n_pts] = d_tot
        ref = torch.empty(n_j_max * N_E * world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(ref, ref_pad)
        if pg is not None:
            assert int(pg.timed_out.item()) == 0, "peer barrier timed out"
            got = pg.gathered()
            sizes = [shard.plan(n_j_total, N_E, world, r)[0] for r in range(world)]
            gather_ok = all(torch.equal(got[r * n_j_max * N_E:r * n_j_max * N_E + (b - a) * N_E],
                                        ref[r * n_j_max * N_E:r * n_j_max * N_E + (b - a) * N_E]) for r, (a, b) in enumerate(sizes))
        else:
            gather_ok = bool(torch.equal(d_gather, ref))
        assert gather_ok, "gathered totals differ from the NCCL all-gather of the same data"

    # ---- kernel-only time (the sweep alone, same stream) for the roofline
    k_ms, _ = env.timed(sweep, args.steps, 3)

    # ---- parity spot-check of the benchmarked arrays against the golden vectors of the reference
    parity = None
    if rank == 0 and world == 1:
        g = np.load(os.path.join(ROOT, "tests", "golden", "cfg2_cicc.npz"))
        Tn = d_T.view(n_j, N_E, NLOC, NLOC)
        Rn = d_R.view(n_j, N_E, NLOC, NLOC)
        worst = 0.0
        for k, jidx in enumerate([0, 250, 500, 750, 999]):
            worst = max(worst, bu.rel_err(Tn[jidx, ::40].cpu().numpy(), g["T"][k]), bu.rel_err(Rn[jidx, ::40].cpu().numpy(), g["R"][k]))
        parity = worst

    # ---- e2e through the public host-buffer API (pinned host memory, copies inside the timed region)
    def time_e2e(fn, reps):
        fn()                                                                     # warm-up (workspace allocation)
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        env.barrier()
        return env.max_over_ranks((time.perf_counter() - t0) / reps)

    n_e2e = max(2, min(args.steps, 5))
    pin = env.pin
    ph0, ph1, ptn, pE, pJ = pin(h0), pin(h1), pin(tnn), pin(Es), pin(Js)
    h2d = int(ph0.nbytes + ph1.nbytes + ptn.nbytes + pE.nbytes + pJ.nbytes)
    R_host = torch.empty((n_j, N_E, NLOC, NLOC), dtype=torch.float64, pin_memory=True).numpy()
    T_host = torch.empty((n_j, N_E, NLOC, NLOC), dtype=torch.float64, pin_memory=True).numpy()

    def full():
        transmission_sweep(ph0, ptn, pE, h1=ph1, params=pJ, out=(R_host, T_host))
        return float(T_host[0, 0, 0, 0])                                         # host read of the result
    dt = time_e2e(full, n_e2e)
    pcie = env.pcie_d2h_gbs()
    e2e = {"value": n_pts_total / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": int(R_host.nbytes + T_host.nbytes), "ms_per_step": dt * 1e3,
           "outputs": "full per-channel R and T [J, E, source, channel] — what wfm.kernel returns per call",
           "pcie_d2h_peak_gbs_measured": pcie, "pcie_frac": (R_host.nbytes + T_host.nbytes) / dt / 1e9 / pcie,
           "api": "transport_b200.sweep.transmission_sweep -> tx_transmission_sweep_spin (host pointers)"}
    del R_host, T_host
    compact = {}
    for spin_n in (2, 4):
        sp = torch.empty((n_j, N_E, NLOC, spin_n), dtype=torch.float64, pin_memory=True).numpy()
        tt = torch.empty((n_j, N_E), dtype=torch.float64, pin_memory=True).numpy()

        def comp(sp=sp, tt=tt, spin_n=spin_n):
            tot, s = transmission_sweep(ph0, ptn, pE, h1=ph1, params=pJ, want_rt=False, want_total=True, want_spin=spin_n,
                                        spin_out=sp, total_out=tt)
            return float(s[0, 0, 0, 0]) + float(tot[0, 0])
        dtc = time_e2e(comp, n_e2e)
        compact[spin_n] = {"value": n_pts_total / dtc, "unit": UNIT, "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": int(sp.nbytes + n_pts * 8), "ms_per_step": dtc * 1e3,
                           "outputs": ("spin-resolved sums per source (T no-flip, T flip%s) + per-point total, reduced in the device "
                                       "epilogue (run_wfm_gate.py:404-405,473-477 reduce the same on the host)"
                                       % (", R no-flip, R flip" if spin_n == 4 else "")),
                           "pcie_frac": (sp.nbytes + n_pts * 8) / dtc / 1e9 / pcie}
        del sp

    # ---- CPU baseline: the oracle's loop-for-loop port on all host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        with mp.get_context("fork").Pool(cores) as pool:
            run_cpu_sample(pool, cores, 1, offset=99)
            rate, wall = run_cpu_sample(pool, cores, 500, offset=0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d random (J,E) points of the same grid x 8 sources, %.1f s wall" % (cores * 500, wall)}

    # totals only: T(E) = Tr[Gamma_R G Gamma_L G^dag] per point, the array north_star's gather is about (8 B per point)
    tt = torch.empty((n_j, N_E), dtype=torch.float64, pin_memory=True).numpy()

    def totals(tt=tt):
        tot = transmission_sweep(ph0, ptn, pE, h1=ph1, params=pJ, want_rt=False, want_total=True, total_out=tt)
        return float(tot[0, 0])
    dtt = time_e2e(totals, n_e2e)
    e2e_total = {"value": n_pts_total / dtt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(tt.nbytes),
                 "ms_per_step": dtt * 1e3, "outputs": "T(E) = Tr[Gamma_R G Gamma_L G^dag] per (J, E) point only",
                 "pcie_frac": tt.nbytes / dtt / 1e9 / pcie}
    del tt

    if pg is not None:
        pg.close()
    fp64_peak = env.fp64_peak()
    hbm_peak, hbm_src = bu.hbm_peak()
    traffic, ncu_note = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the sweep kernel from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get("bytes_per_launch")
        ncu_note = {k: tj.get(k) for k in ("kernel", "binding", "fp64_plus_dmma_shared_pipe_pct_of_peak", "dmma_subpipe_pct_of_peak",
                                           "fp64_pipe_pct_of_peak", "lsu_data_pipe_pct_of_peak", "issue_active_pct_of_peak", "source")}
    except Exception:
        pass
    ach_tf = FLOPS_PER_POINT * n_pts / (k_ms * 1e-3) / 1e12
    ach_gbs = BYTES_PER_POINT * n_pts / (k_ms * 1e-3) / 1e9
    exe_tf = EXECUTED_FLOPS_PER_POINT * n_pts / (k_ms * 1e-3) / 1e12
    return {
        "value": value, "ms_per_step": ms_per_step, "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
        "e2e_compact": compact[2], "e2e_compact4": compact[4], "e2e_total": e2e_total, "cpu_baseline": cpu,
        "parity_max_rel_err": parity,
        "parity_against": "tests/golden/cfg2_cicc.npz (unmodified reference, 5 J x 25 E x 8 sources)" if parity is not None else None,
        "gather": {"mode": {"peer": "fused: K3 epilogue stores the per-point totals into every rank's buffer over NVLink (CUDA IPC peer "
                                    "memory) + device-side flag barrier; no collective",
                            "nccl": "NCCL all_gather_into_tensor after the sweep", "none": "single GPU"}[gather_mode],
                   "verified_equal_to_nccl_all_gather": gather_ok, "bytes_per_rank": n_pts * 8},
        "points_per_gpu": n_pts, "wall_s_timed_region": t_wall,
        "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                     "traffic": traffic,
                     "traffic_source": "profiles/traffic.json (ncu --set full, dram bytes read+written per launch)",
                     "ncu": ncu_note,
                     "peak_source": "tx_measure_fp64_peak: register-resident DFMA loop, measured in this run "
                                    "(MEASURED_PEAKS.json carries no FP64 figure)",
                     "flops_per_point": FLOPS_PER_POINT, "kernel_ms": k_ms,
                     "executed": {"flops_per_point": EXECUTED_FLOPS_PER_POINT, "achieved": exe_tf, "frac": exe_tf / fp64_peak,
                                  "note": "frac above is ALGORITHMIC flops (SURVEY 8d: one inverse + one product per site) over the "
                                          "measured DFMA peak; the telescoped sweep executes fewer, so the algorithmic fraction can exceed 1"},
                     "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                             "bytes_per_point": BYTES_PER_POINT, "peak_source": hbm_src}},
    }


def gpu_arm(args):
    import bench_configs as bc
    env = bu.Env()
    # NCCL (and torchrun banners) may write to stdout; the contract is ONE JSON line there, so everything
    # else goes to stderr and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    env.init_dist()
    rank, world = env.rank, env.world
    name = args.config
    extra = {}
    if name == "cfg2":
        res = run_cfg2(env, args, args.scaling)
        if world > 1 and args.scaling == "weak" and not args.no_strong:
            # the same grid as ONE 10^6-point job sharded over the ranks (north_star's own phrasing of config 2)
            s_args = argparse.Namespace(**vars(args))
            s_args.no_cpu = True
            s = run_cfg2(env, s_args, "strong")
            extra["strong_scaling"] = {
                "value": s["value"], "unit": UNIT, "ms_per_step": s["ms_per_step"], "points_total": N_J * N_E,
                "points_per_gpu": s["points_per_gpu"], "kernel_ms": s["roofline"]["kernel_ms"],
                "e2e": s["e2e"], "e2e_compact": s["e2e_compact"], "e2e_total": s["e2e_total"],
                "limiter": "fixed per-step cost (structure scan, affine expansion, the idle instantiation's launch, the flag "
                           "barrier) = ms_per_step - kernel_ms = %.3f ms against %.3f ms of sweep" %
                           (s["ms_per_step"] - s["roofline"]["kernel_ms"], s["roofline"]["kernel_ms"])}
    else:
        res = bc.run(name, env, steps=args.steps if args.steps_given else {"cfg1": 50, "cfg3": 5, "cfg4": 3, "cfg5": 10}[name],
                     warmup=args.warmup, scaling=args.scaling, want_cpu=not args.no_cpu)
    secondary = None
    if name == "cfg2" and world == 1 and not args.no_secondary:
        secondary = {}
        for k, st in (("cfg1", 20), ("cfg3", 3), ("cfg4", 2), ("cfg5", 5)):
            try:
                secondary[k] = bc.run(k, env, steps=st, warmup=3, scaling="weak", want_cpu=not args.no_cpu)
            except Exception as exc:    # a secondary config must not take the headline line down with it
                secondary[k] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if rank == 0:
        unit = res.get("unit", UNIT)
        line = {
            "metric": METRIC if unit == UNIT else "bardeen geometries/sec", "value": res["value"], "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
            "config": config_dict(name, world, args.scaling),
            "run": {"l2": "256 MB flush between timed steps (outside the event pair); every step also streams its own result arrays",
                    "host_affinity": env.affinity, "timing": "CUDA events on the launching stream, max over ranks"},
            "e2e": res.get("e2e"), "gpu_launches": res.get("gpu_launches"), "roofline": res.get("roofline"),
            "cpu_baseline": res.get("cpu_baseline"), "clocks": res.get("clocks"),
        }
        for k, v in res.items():
            if k not in line and k not in ("workload", "unit", "n_gpus", "scaling"):
                line[k] = v
        line.update(extra)
        if secondary is not None:
            line["secondary"] = secondary
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if env.dist is not None:
        env.dist.destroy_process_group()


def main():
    import faulthandler
    # a hung leg dumps every thread's stack to stderr and ends the process instead of burning the box's time limit
    faulthandler.dump_traceback_later(int(os.environ.get("TX_BENCH_WATCHDOG", "1200")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary block (configs 1, 3, 4, 5)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-run at N > 1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 200 if args.impl == "ours" else 5
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
