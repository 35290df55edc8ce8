#!/usr/bin/env python
"""Secondary measurements: BASELINE configs 1, 3, 4 and 5 on one GPU (config 2 is bench.py's headline).

    python scripts/bench_configs.py [cfg1] [cfg3] [cfg4] [cfg5] [--small]

For each config: device-resident sweep timed with CUDA events (3 warm-ups, L2 flush between steps),
parity of a subsample against the numpy oracle, roofline fraction (FP64), one JSON line each.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rgf_oracle  # noqa: E402  (checker only)
from transport_b200 import _lib, configs, leads  # noqa: E402

lib = _lib.load()
if os.environ.get("TX_VARIANT"):
    _lib.check(lib.tx_set_variant(int(os.environ["TX_VARIANT"])))
if os.environ.get("TX_KMAX"):
    _lib.check(lib.tx_set_option(1, int(os.environ["TX_KMAX"])))
if os.environ.get("TX_GEXP"):
    _lib.check(lib.tx_set_option(2, int(os.environ["TX_GEXP"])))
if os.environ.get("TX_NO_STRUCTURE"):
    _lib.check(lib.tx_set_option(0, 0))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def dev_c(a):
    return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))).to(dev)


def fp64_peak():
    v, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.tx_measure_fp64_peak(ctypes.byref(v), ctypes.byref(ms)))
    return v.value


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms))


def rel_err(got, want):
    scale = np.maximum(np.abs(want), 1e-13 * np.abs(want).max(axis=-1, keepdims=True))
    return float(np.max(np.abs(got - want) / np.where(scale == 0, 1, scale)))


def run_small(name, h, tnn, Es, n_src_all, flops_per_point, check_idx, steps):
    """Closed-form leads, scalar hopping, n <= 16: tx_transmission_sweep_dev."""
    n, M, nE = h.shape[-1], h.shape[0], len(Es)
    d_h, d_t, d_E = dev_c(h), dev_c(tnn), dev_c(Es)
    R = torch.empty((nE, n, n), dtype=torch.float64, device=dev)
    T = torch.empty_like(R)
    res = torch.empty((nE, n), dtype=torch.float64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(lib.tx_transmission_sweep_dev(d_h.data_ptr(), 1, None, None, d_t.data_ptr(), 1, n, M, 1,
                                                 d_E.data_ptr(), nE, _lib.TX_LEAD_CLOSED, None, None, 1,
                                                 _lib.TX_HOP_SCALAR, None, None, n, R.data_ptr(), T.data_ptr(),
                                                 res.data_ptr(), None, flag.data_ptr(), st))
    ms = timed(step, steps=steps)
    assert int(flag.item()) == 0
    Th = T.cpu().numpy()
    Rh = R.cpu().numpy()
    resid = float(np.nanmax(np.abs(res.cpu().numpy())))
    t0 = time.perf_counter()
    oR, oT = rgf_oracle.transmission_closed(h, tnn, Es[check_idx], np.eye(n))
    cpu_s = time.perf_counter() - t0
    err = max(rel_err(Th[check_idx], oT), rel_err(Rh[check_idx], oR))
    # conditioning floor: how much the ORACLE's own answer moves when E is perturbed by ~1 ulp
    pR, pT = rgf_oracle.transmission_closed(h, tnn, Es[check_idx] * (1 + 2.2e-16), np.eye(n))
    floor = max(rel_err(pT, oT), rel_err(pR, oR))
    peak = fp64_peak()
    ach = flops_per_point * nE / (ms * 1e-3) / 1e12
    return {"config": name, "n": n, "M": M, "points": nE, "sources": n, "ms_per_sweep": ms,
            "points_per_s": nE / (ms * 1e-3), "flops_per_point": flops_per_point,
            "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak},
            "parity_max_rel_err_vs_oracle": err, "oracle_change_under_1ulp_energy_perturbation": floor,
            "checked_points": int(len(check_idx)),
            "unitarity_max_abs": resid,
            "cpu_oracle_rgf_numpy_points_per_s_1core": len(check_idx) / cpu_s}


def cfg1():
    h, tnn, _ = configs.barrier_blocks(0.4, 11)
    Es = configs.barrier_energies(10_000)
    return run_small("cfg1 rbbarrier n=1 M=13, 1e4 energies", h, tnn, Es, 1, 8 * 1 * 13 * 2, np.arange(0, 10_000, 97), 10)


def cfg3(small=False):
    N = 1000 if small else 10_000
    nE = 10_000 if small else 100_000
    h, tnn, _ = configs.disordered_blocks(N=N, n=8, W=0.05, seed=1234)
    Es = np.linspace(-1.9, 1.9, nE).astype(complex)
    idx = np.arange(0, nE, nE // 8)
    return run_small("cfg3 disordered chain n=8 N=%d, %d energies" % (N, nE), h, tnn, Es, 8,
                     8 * 8 ** 3 * (N + 2) * 2, idx, 3)


def cfg4(small=False):
    n = 32 if small else 64
    nE = 256 if small else 4096
    h, tnn, _ = configs.ribbon_blocks(width=n, L=32, W=1.0, seed=1234)
    Es = np.linspace(-3.5, 3.5, nE).astype(complex)
    eta = 1e-8
    M = h.shape[0]
    d_h, d_t, d_E = dev_c(h), dev_c(tnn), dev_c(Es)
    sigL = torch.empty((nE, n, n, 2), dtype=torch.float64, device=dev)
    sigR = torch.empty_like(sigL)
    nit = torch.zeros((2, nE), dtype=torch.int32, device=dev)
    ok = torch.zeros((2, nE), dtype=torch.int32, device=dev)
    conv = torch.zeros((2, nE), dtype=torch.float64, device=dev)
    car = torch.empty(nE, dtype=torch.float64, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    wse = int(lib.tx_sweep_workspace_elems(n, nE))
    work = torch.empty((max(wse, 1), 2), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    nn16 = n * n * 16

    def leads_step():
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
    def leads_separable_step():   # the same leads by the separable closed form (h01 = -I): eta -> 0 limit
        _lib.check(lib.tx_surface_gf_dev(d_h.data_ptr(), d_t.data_ptr(), n, d_E.data_ptr(), nE, -1, _lib.TX_GF_SEPARABLE,
                                         0.0, 0.0, 0, 0, 0, None, sigL.data_ptr(), None, None, None, flag.data_ptr(), st))
        _lib.check(lib.tx_surface_gf_dev(d_h.data_ptr() + (M - 1) * nn16, d_t.data_ptr() + (M - 2) * nn16, n,
                                         d_E.data_ptr(), nE, 1, _lib.TX_GF_SEPARABLE, 0.0, 0.0, 0, 0, 0, None,
                                         sigR.data_ptr(), None, None, None, flag.data_ptr(), st))
    ms_leads_sep = timed(leads_separable_step, steps=2, warmup=1)
    sweep_step()
    torch.cuda.synchronize()
    car_sep = car.cpu().numpy().copy()
    ms_leads = timed(leads_step, steps=2, warmup=1)
    ms_sweep = timed(sweep_step, steps=2, warmup=1)
    assert bool(ok.all().item()), "Sancho-Rubio did not converge everywhere"
    n_it = float(nit.double().mean().item())
    carh = car.cpu().numpy()
    idx = np.arange(0, nE, max(1, nE // 6))
    SL = torch.view_as_complex(sigL[idx]).cpu().numpy()
    SR = torch.view_as_complex(sigR[idx]).cpu().numpy()
    _, GN0 = rgf_oracle.rgf_blocks(h, tnn, Es[idx], SL, SR)        # same tables on both sides
    want = rgf_oracle.caroli_trace(GN0, SL, SR)
    err = float(np.max(np.abs(carh[idx] - want) / np.maximum(np.abs(want), 1e-12)))
    flops = 8 * n ** 3 * M * 2 + 2 * n_it * (8 + 6 * 8) * n ** 3
    peak = fp64_peak()
    ms = ms_leads + ms_sweep
    ach = flops * nE / (ms * 1e-3) / 1e12
    return {"config": "cfg4 ribbon n=%d L=32, %d energies, Sancho-Rubio eta=1e-8" % (n, nE), "n": n, "M": M,
            "points": nE, "ms_leads": ms_leads, "ms_sweep": ms_sweep, "points_per_s": nE / (ms * 1e-3),
            "sancho_iterations_mean": n_it, "flops_per_point": flops,
            "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak},
            "parity_caroli_max_rel_err_vs_oracle_same_tables": err, "checked_points": int(len(idx)),
            "separable_leads": {"ms_leads": ms_leads_sep, "points_per_s": nE / ((ms_leads_sep + ms_sweep) * 1e-3),
                                "median_rel_diff_of_T_vs_sancho_eta_1e-8": float(np.median(np.abs(car_sep - carh) / np.maximum(np.abs(carh), 1e-12))),
                                "note": "TX_GF_SEPARABLE: h01 = -I commutes with h00, g = U diag(g_closed) U^dag (eta -> 0)"},
            "kernel": "CTA-per-point generic kernel; DMMA block products, shared-memory staged Gauss-Jordan inverse"}


def cfg5(small=False):
    """rbstep.py-style sweep of the right-well step height (rbstep.py:43-104): P geometries in one device pass —
    bound states of both wells, window-averaged matrix elements, current over the bias grid of rbbarrier.py:447-450."""
    from oracle import bardeen_oracle as bo                      # checker + CPU baseline only
    from transport_b200 import bardeen
    out = []
    one = np.eye(1)
    sizes = (200,) if small else (200, 800)
    if os.environ.get("TX_CFG5_NLR"):
        sizes = (int(os.environ["TX_CFG5_NLR"]),)
    if os.environ.get("TX_EIG_MODE"):
        _lib.check(lib.tx_set_option(3, int(os.environ["TX_EIG_MODE"])))
    for NLR in sizes:
        P = 64 if small else 256
        VRs = np.linspace(-0.05, 0.05, P)
        Vinf, VL = 0.5 * one, 0.0 * one
        HC = np.zeros((11, 11, 1, 1), dtype=complex)
        for j in range(11):
            HC[j, j] = 0.4
        for j in range(10):
            HC[j, j + 1] = HC[j + 1, j] = -1.0
        arrs = {k: [] for k in ("dL", "eL", "dR", "eR", "h0", "hU", "hL")}
        t0 = time.time()
        for v in VRs:
            VR = v * one
            bL = bardeen.hsys_site_blocks(one, one, one, Vinf, VL, Vinf, 20, NLR, NLR, HC)
            bR = bardeen.hsys_site_blocks(one, one, one, Vinf, VR + Vinf, VR, 20, NLR, NLR, HC)
            bS = bardeen.hsys_site_blocks(one, one, one, Vinf, VL, VR, 20, NLR, NLR, HC)
            a, b = bardeen._channel_tridiagonals(bL); arrs["dL"].append(a); arrs["eL"].append(b)
            a, b = bardeen._channel_tridiagonals(bR); arrs["dR"].append(a); arrs["eR"].append(b)
            hd = [np.real(x - y) for x, y in zip(bS, bL)]
            arrs["h0"].append(hd[0]); arrs["hU"].append(hd[1]); arrs["hL"].append(hd[2])
        host_assembly_s = time.time() - t0
        A = {k: np.ascontiguousarray(np.array(v), dtype=np.float64) for k, v in arrs.items()}
        S = A["dL"].shape[-1]
        interval = 1e-2
        cut = np.full((P, 1), 0.1 - 2.0)
        # e2e through the host entry (H2D + counts + second pass sized from the counts + D2H)
        t0 = time.time()
        r = bardeen.wells_batched(A["dL"], A["eL"], A["dR"], A["eR"], cut, cut + interval, cut, A["h0"], A["hU"], A["hL"],
                                  interval, sign_fix=True)
        e2e_first_s = time.time() - t0
        t0 = time.time()
        r = bardeen.wells_batched(A["dL"], A["eL"], A["dR"], A["eR"], cut, cut + interval, cut, A["h0"], A["hU"], A["hL"],
                                  interval, sign_fix=True)
        e2e_s = time.time() - t0
        ms = int(max(r["nL"].max(), r["nR_states"].max()))
        # device resident
        D = {k: torch.from_numpy(v).to(dev) for k, v in A.items()}
        dcut = torch.from_numpy(cut).to(dev); dcut2 = torch.from_numpy(cut + interval).to(dev)
        EL = torch.zeros((P, 1, ms), dtype=torch.float64, device=dev); ER = torch.zeros_like(EL)
        Mavg = torch.zeros((P, 1, ms, 1), dtype=torch.float64, device=dev)
        nL = torch.zeros((P, 1), dtype=torch.int32, device=dev); nRs = torch.zeros_like(nL); nRb = torch.zeros_like(nL)
        ws = torch.empty(int(lib.tx_bardeen_workspace_bytes(P, 1, S, ms)), dtype=torch.uint8, device=dev)
        Vb = torch.linspace(-0.05, 0.05, 99, dtype=torch.float64, device=dev)
        Iab = torch.zeros((P, 1, 1, 99), dtype=torch.float64, device=dev)
        M2 = torch.zeros_like(Mavg)
        st = torch.cuda.current_stream().cuda_stream

        def wells_step():
            _lib.check(lib.tx_bardeen_wells_dev(D["dL"].data_ptr(), D["eL"].data_ptr(), D["dR"].data_ptr(), D["eR"].data_ptr(),
                                                dcut.data_ptr(), dcut2.data_ptr(), dcut.data_ptr(), D["h0"].data_ptr(),
                                                D["hU"].data_ptr(), D["hL"].data_ptr(), P, 1, S, ms, interval, 1,
                                                EL.data_ptr(), nL.data_ptr(), ER.data_ptr(), nRs.data_ptr(), nRb.data_ptr(),
                                                Mavg.data_ptr(), ws.data_ptr(), st))

        def step():
            wells_step()
            torch.mul(Mavg, Mavg, out=M2)
            _lib.check(lib.tx_bardeen_current_dev(EL.data_ptr(), M2.data_ptr(), P, 1, ms, Vb.data_ptr(), 99, -1.95, 0.01,
                                                  Iab.data_ptr(), st))
        def eigenvalues_only_step():
            for dd, ee, cc, nn, EE in ((D["dL"], D["eL"], dcut, nL, EL), (D["dR"], D["eR"], dcut2, nRs, ER)):
                _lib.check(lib.tx_tridiag_eigh_below_dev(dd.data_ptr(), ee.data_ptr(), S, P, cc.data_ptr(), nn.data_ptr(), ms,
                                                         EE.data_ptr(), None, None, st))
        ms_step = timed(step, steps=5, warmup=3)
        ms_evals = timed(eigenvalues_only_step, steps=5, warmup=2)
        ms_wells = timed(wells_step, steps=5, warmup=3)
        assert np.array_equal(nL.cpu().numpy(), r["nL"]) and np.allclose(Mavg.cpu().numpy(), r["Mavg"], rtol=1e-12, atol=0)
        # parity + CPU baseline: the oracle's kernel_well (dense LAPACK eigh x4, as the reference) on a few geometries
        idx = [0, P // 2, P - 1] if NLR == 200 else [P // 3]
        err, cpu_s = 0.0, 0.0
        for p in idx:
            VR = VRs[p] * one
            t0 = time.time()
            E0, M0 = bo.kernel_well(one, one, one, Vinf, VL, VR + Vinf, VR, Vinf, 20, NLR, NLR, HC, HC, one, 0.1 * one,
                                    interval=interval)
            cpu_s += time.time() - t0
            mu, nu = int(r["nL"][p, 0]), int(r["nR_below"][p, 0])
            M = r["Mavg"][p, 0, :mu, 0] ** 2
            M[nu:] = 0.0
            big = M0[0, :, 0] > 1e-30
            err = max(err, float(np.max(np.abs(r["EL"][p, 0, :mu] - E0[0].real))),
                      float(np.max(np.abs(M[big] - M0[0, big, 0]) / M0[0, big, 0])) if big.any() else 0.0)
        states = int(r["nL"].sum() + r["nR_states"].sum())
        out.append({"config": "cfg5 bardeen rbstep sweep: %d step heights, N_L=N_R=%d (S=%d), interval 1e-2, 99 bias points" % (P, NLR, S),
                    "geometries": P, "S": S, "bound_states_total": states, "max_states": ms,
                    "ms_per_sweep_device_resident": ms_step, "ms_wells_only": ms_wells, "ms_eigenvalues_only": ms_evals,
                    "geometries_per_s": P / (ms_step * 1e-3),
                    "e2e_host_call_s": e2e_s, "e2e_first_call_s": e2e_first_s, "e2e_geometries_per_s": P / e2e_s,
                    "host_assembly_s": host_assembly_s,
                    "cpu_oracle_dense_eigh_s_per_geometry_1core": cpu_s / len(idx),
                    "speedup_device_resident_vs_1core": (cpu_s / len(idx)) / (ms_step * 1e-3 / P),
                    "parity_max_err_vs_oracle(E abs, |M|^2 rel)": err, "checked_geometries": len(idx),
                    "kernel": "Sturm multisection (warp per eigenvalue) + twisted factorisation + windowed dot products"})
    return out


if __name__ == "__main__":
    small = "--small" in sys.argv
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg1", "cfg3", "cfg4"]
    for w in which:
        out = {"cfg1": cfg1, "cfg3": lambda: cfg3(small), "cfg4": lambda: cfg4(small), "cfg5": lambda: cfg5(small)}[w]()
        for line in (out if isinstance(out, list) else [out]):
            print(json.dumps(line), flush=True)
