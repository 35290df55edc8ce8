#!/usr/bin/env python
"""Benchmark of the T(E) hot path on the BASELINE.json headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--variant V]

Workload (BASELINE.json configs[1]): two spin-1/2 impurities, n_loc_dof = 8,
M = 8 blocks, H(J) = H0 + J*H1, J = linspace(-1,-0.001,1000) x E = logspace(-4,0,1000) - 2 -> 10^6
(energy, parameter) points PER GPU (weak scaling: rank r takes its own 1000-value J slice of a
1000*N-value grid), all 8 one-hot sources per point.  A step = one pass of the sweep over the
rank's 10^6 points with inputs resident in HBM, plus (N>1) the NCCL all-gather of the compact
per-point transmission totals.

One JSON line is printed by rank 0 (see the task contract): metric value, e2e through the public
host-buffer API, roofline of the dominant kernel (FP64 FMA pipe; HBM figures alongside), the CPU
baseline (the oracle's loop-for-loop port of the reference, all host cores) and clocks.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_J, N_E, NLOC, MBLK = 1000, 1000, 8, 8
FLOPS_PER_POINT = 8 * NLOC ** 3 * MBLK * 2          # 8 n^3 M (1+g), g=1 scalar hopping  
# What the telescoped kernel really executes per point on this geometry: 2 dense recurrence
# sites + the g_0 product (3 x 8 n^3 on DMMA), 1 inverse (8 n^3), 5 diagonal-site column scalings and the
# chunk start (6 x 8 n^2), prologue/epilogue O(n^2).  The algorithmic figure above counts one inverse and
# one product per site, which the telescoping removes; both fractions are reported.
EXECUTED_FLOPS_PER_POINT = 4 * 8 * NLOC ** 3 + 6 * 8 * NLOC ** 2 + 40 * NLOC ** 2
BYTES_PER_POINT = 24 + 16 * NLOC * NLOC             # unique bytes: E, J, R and T for 8 sources
METRIC = "T(E) points/sec (energy x param)"
UNIT = "points/s"
WORKLOAD = "cfg2: rbmenez-style two spin-1/2 impurities, n_loc_dof=8, M=8, 1000 J x 1000 E, all 8 sources"


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """Times the oracle's loop-for-loop port of wfm.kernel on a slice of (J, E) points."""
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
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
    """`--impl reference`: the reference's CPU algorithm (oracle port, Python loops + dense inverse,
    transport/wfm/__init__.py:29-215) on all host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
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
    sample = "%d random (J,E) points of the cfg2 grid x 8 sources per step, %d steps" % (cores * pts_per_core, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [x.strip() for x in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import torch
    from transport_b200 import _lib, configs
    from transport_b200.sweep import transmission_sweep

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    try:   # run on the CPUs next to this GPU, so that pinned host buffers are allocated on its NUMA node
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        numa = "cpus %d" % len(os.sched_getaffinity(0))
    except Exception as exc:   # affinity is an optimisation of the e2e leg only
        numa = "unbound (%s)" % type(exc).__name__
    lib = _lib.load()
    _lib.check(lib.tx_set_device(local_rank))
    # NCCL (and torchrun banners) may write to stdout; the contract is ONE JSON line there, so everything
    # else goes to stderr and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(lib.tx_set_variant(args.variant))

    # ---- inputs: this rank's J slice of the 1000*world grid, resident in HBM
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, 3, 2)
    from transport_b200 import shard
    Js_all = np.linspace(-1.0, -0.001, N_J * world)
    (j0, j1), _ = shard.plan(N_J * world, N_E, world, rank)      # parameter axis sharded, energies whole
    Js = np.ascontiguousarray(Js_all[j0:j1])
    assert len(Js) == N_J
    _, Es = configs.cfg2_grid(N_J, N_E)
    n_pts = N_J * N_E

    def dev_c(a):
        return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))).to(dev)

    d_h0, d_h1, d_tnn, d_E = dev_c(h0), dev_c(h1), dev_c(tnn), dev_c(Es)
    d_J = torch.from_numpy(Js).to(dev)
    d_R = torch.empty((n_pts, NLOC, NLOC), dtype=torch.float64, device=dev)
    d_T = torch.empty_like(d_R)
    d_tot = torch.empty(n_pts, dtype=torch.float64, device=dev)
    d_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    d_gather = torch.empty(n_pts * world, dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    stream = torch.cuda.current_stream()

    def step():
        rc = lib.tx_transmission_sweep_dev(
            d_h0.data_ptr(), 1, d_h1.data_ptr(), d_J.data_ptr(), d_tnn.data_ptr(), 1, NLOC, MBLK, N_J,
            d_E.data_ptr(), N_E, _lib.TX_LEAD_CLOSED, None, None, 1, _lib.TX_HOP_SCALAR,
            None, None, NLOC, d_R.data_ptr(), d_T.data_ptr(), None, d_tot.data_ptr(),
            d_flag.data_ptr(), stream.cuda_stream)
        _lib.check(rc)
        if world > 1:
            dist.all_gather_into_tensor(d_gather, d_tot)

    # ---- FP64 peak (roofline denominator; MEASURED_PEAKS.json has no FP64 entry)
    import ctypes
    pk, pkms = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.tx_measure_fp64_peak(ctypes.byref(pk), ctypes.byref(pkms)))
    fp64_peak = pk.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    assert int(d_flag.item()) == 0, "singular block in the benchmark sweep"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = int(lib.tx_launch_counter())
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = int(lib.tx_launch_counter()) - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total = float(sum(ms_steps))
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n_pts * world / (ms_per_step * 1e-3)

    # ---- kernel-only time (the sweep kernel alone, same stream) for the roofline
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        kev[k][0].record(stream)
        rc = lib.tx_transmission_sweep_dev(
            d_h0.data_ptr(), 1, d_h1.data_ptr(), d_J.data_ptr(), d_tnn.data_ptr(), 1, NLOC, MBLK, N_J,
            d_E.data_ptr(), N_E, _lib.TX_LEAD_CLOSED, None, None, 1, _lib.TX_HOP_SCALAR,
            None, None, NLOC, d_R.data_ptr(), d_T.data_ptr(), None, d_tot.data_ptr(),
            d_flag.data_ptr(), stream.cuda_stream)
        kev[k][1].record(stream)
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))

    # ---- parity spot-check of the benchmarked arrays against the golden vectors of the reference
    parity = None
    if rank == 0 and world == 1:
        g = np.load(os.path.join(ROOT, "tests", "golden", "cfg2_cicc.npz"))
        Tn = d_T.view(N_J, N_E, NLOC, NLOC)
        worst = 0.0
        for k, jidx in enumerate([0, 250, 500, 750, 999]):
            got = Tn[jidx, ::40].cpu().numpy()
            want = g["T"][k]
            scale = np.maximum(np.abs(want), 1e-13 * np.abs(want).max(axis=-1, keepdims=True))
            worst = max(worst, float(np.max(np.abs(got - want) / np.where(scale == 0, 1, scale))))
        parity = worst

    # ---- e2e through the public host-buffer API (pinned host memory, copies inside the timed region)
    e2e = None
    if rank == 0 or world > 1:
        R_host = torch.empty((N_J, N_E, NLOC, NLOC), dtype=torch.float64, pin_memory=True)
        T_host = torch.empty_like(R_host).pin_memory()
        Rn, Tn_ = R_host.numpy(), T_host.numpy()
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
        ph0, ph1, ptn, pE, pJ = pin(h0), pin(h1), pin(tnn), pin(Es), pin(Js)
        transmission_sweep(ph0, ptn, pE, h1=ph1, params=pJ, out=(Rn, Tn_))      # warm-up (workspace alloc)
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            transmission_sweep(ph0, ptn, pE, h1=ph1, params=pJ, out=(Rn, Tn_))
            _ = float(Tn_[0, 0, 0, 0])                                           # host read of the result
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n_pts * world / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(ph0.nbytes + ph1.nbytes + ptn.nbytes + pE.nbytes + pJ.nbytes),
               "d2h_bytes_per_step": int(Rn.nbytes + Tn_.nbytes), "ms_per_step": dt * 1e3,
               "api": "transport_b200.sweep.transmission_sweep -> tx_transmission_sweep (host pointers)"}

    # ---- CPU baseline: the oracle's loop-for-loop port on all host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        with mp.get_context("fork").Pool(cores) as pool:
            run_cpu_sample(pool, cores, 1, offset=99)
            rate, wall = run_cpu_sample(pool, cores, 500, offset=0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d random (J,E) points of the same grid x 8 sources, %.1f s wall" % (cores * 500, wall)}

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, ncu_note = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of the sweep kernel from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj.get("bytes_per_launch")
            ncu_note = {"kernel": tj.get("kernel"),
                        "binding": "no single unit saturated: latency bound at 8 warps/SM (255 registers, 113 KB smem per CTA)",
                        "fp64_plus_dmma_pipe_pct_of_peak": tj.get("fp64_plus_dmma_shared_pipe_pct_of_peak"),
                        "dmma_subpipe_pct_of_peak": tj.get("dmma_subpipe_pct_of_peak"),
                        "fp64_pipe_pct_of_peak": tj.get("fp64_pipe_pct_of_peak"),
                        "lsu_data_pipe_pct_of_peak": tj.get("lsu_data_pipe_pct_of_peak"),
                        "issue_active_pct_of_peak": tj.get("issue_active_pct_of_peak")}
        except Exception:
            pass
        achieved_tf = FLOPS_PER_POINT * n_pts / (k_ms * 1e-3) / 1e12
        achieved_gbs = BYTES_PER_POINT * n_pts / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_gpu": n_pts, "l2": "256 MB flush between timed steps; "
                       "each step also writes 2.05 GB of R/T", "variant": args.variant,
                       "parallelism": "J axis sharded, dp%d" % world, "host_affinity": numa},
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak, "traffic": traffic,
                         "traffic_source": "profiles/traffic.json (ncu --set full, dram bytes read+written per launch)",
                         "ncu": ncu_note,
                         "peak_source": "tx_measure_fp64_peak: register-resident DFMA loop, measured in this run "
                                        "(MEASURED_PEAKS.json carries no FP64 figure)",
                         "flops_per_point": FLOPS_PER_POINT, "kernel_ms": k_ms,
                         "executed": {"flops_per_point": EXECUTED_FLOPS_PER_POINT,
                                      "achieved": EXECUTED_FLOPS_PER_POINT * n_pts / (k_ms * 1e-3) / 1e12,
                                      "frac": EXECUTED_FLOPS_PER_POINT * n_pts / (k_ms * 1e-3) / 1e12 / fp64_peak,
                                      "note": "frac above is ALGORITHMIC flops (SURVEY 8d: one inverse + one product per "
                                              "site) over the measured DFMA peak; the telescoped sweep executes fewer, "
                                              "so the algorithmic fraction can exceed 1"},
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": achieved_gbs / hbm_peak, "bytes_per_point": BYTES_PER_POINT,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}},
            "cpu_baseline": cpu, "clocks": clocks, "parity_max_rel_err_vs_reference_golden": parity,
            "wall_s_timed_region": t_wall,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=13)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
