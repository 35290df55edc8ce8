"""Shared measurement helpers of bench.py and scripts/bench_configs.py (timing rules of the task contract):
CUDA-event timing on the launching stream after warm-up, L2 flush between timed iterations, clocks and throttle
reasons sampled DURING the timed region, max over ranks."""
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "points/s"
METRIC = "T(E) points/sec (energy x param)"


class Env:
    """One process per GPU: rank / world from torchrun's environment, device, library handle, L2 flush buffer."""

    def __init__(self):
        import torch
        from transport_b200 import _lib
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench needs a CUDA device (there is no CPU fallback for the product path)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.lib = _lib.load()
        _lib.check(self.lib.tx_set_device(self.local_rank))
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # 256 MB > 126 MB L2
        self.stream = torch.cuda.current_stream()
        self.dist = None
        self._peak = None
        self._pcie = None
        self.nvml_handle = None
        self.affinity = self._bind_host()

    def _bind_host(self):
        """Run on the CPUs next to this GPU so that pinned host buffers land on its NUMA node."""
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(self.torch.cuda.get_device_properties(self.local_rank).uuid)
                self.nvml_handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.nvml_handle = pynvml.nvmlDeviceGetHandleByIndex(self.local_rank)
            pynvml.nvmlDeviceSetCpuAffinity(self.nvml_handle)
            return "cpus %d" % len(os.sched_getaffinity(0))
        except Exception as exc:   # affinity is an optimisation of the e2e leg only
            return "unbound (%s)" % type(exc).__name__

    def init_dist(self):
        if self.world > 1 and self.dist is None:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def dev_c(self, a):
        torch = self.torch
        return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))).to(self.dev)

    def dev_f(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.dev)

    def pin(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    def fp64_peak(self):
        """Register-resident DFMA loop (tx_measure_fp64_peak): MEASURED_PEAKS.json carries no FP64 figure."""
        if self._peak is None:
            from transport_b200 import _lib
            v, ms = ctypes.c_double(), ctypes.c_double()
            _lib.check(self.lib.tx_measure_fp64_peak(ctypes.byref(v), ctypes.byref(ms)))
            self._peak = v.value
        return self._peak

    def pcie_d2h_gbs(self):
        """Measured pinned device->host copy rate (256 MB, best of 3): the roofline of the full-output e2e path."""
        if self._pcie is None:
            torch = self.torch
            nbytes = 256 << 20
            src = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            best = 0.0
            for _ in range(4):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                dst.copy_(src, non_blocking=True)
                b.record()
                torch.cuda.synchronize()
                best = max(best, nbytes / (a.elapsed_time(b) * 1e-3) / 1e9)
            self._pcie = best
        return self._pcie

    def timed(self, step, steps, warmup, sampler=None):
        """`warmup` (>= 3) untimed steps, then `steps` timed ones: CUDA events on the launching stream around each
        step, a 256 MB L2 flush between steps (outside the event pair), barrier + synchronize on both sides.
        Returns (mean ms per step as max over ranks of the summed step times, wall seconds)."""
        torch = self.torch
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        if sampler is not None:
            sampler.start()
        self.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            self.flush.zero_()
            ev[k][0].record(self.stream)
            step()
            ev[k][1].record(self.stream)
        self.barrier()
        wall = time.perf_counter() - t0
        if sampler is not None:
            sampler.stop()
        total = float(sum(a.elapsed_time(b) for a, b in ev))
        return self.max_over_ranks(total) / steps, wall


class ClockSampler:
    """SM clock and throttle reasons polled through NVML from a thread while the timed region runs (nvidia-smi takes
    longer to start than a short timed region lasts)."""

    def __init__(self, env, period_s=0.004):
        self.env = env
        self.period = period_s
        self.rows = []
        self._stop = threading.Event()
        self.thread = None

    def _poll(self):
        import pynvml
        h = self.env.nvml_handle
        while not self._stop.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                util = pynvml.nvmlDeviceGetUtilizationRates(h).gpu
                self.rows.append((sm, reasons, util))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.env.nvml_handle is None:
            return
        self._stop.clear()
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            self.thread = None

    def summary(self):
        if self.env.nvml_handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        import pynvml
        try:
            smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.env.nvml_handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            smax = None
        names = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        seen = set()
        for _, reasons, _ in self.rows:
            for name, bit in names.items():
                if reasons & bit:
                    seen.add(name)
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "samples": len(sm),
                "reasons": sorted(seen)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def hbm_peak():
    """(GB/s, source): the driver-written measured copy bandwidth, else the profiling recipe's fallback."""
    p = peaks()
    if "hbm_gbs" in p:
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def rel_err(got, want):
    """Worst elementwise relative error of per-channel probabilities; entries that vanish by symmetry are measured
    against the largest channel of the same row (tests/conftest.py rt_err)."""
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    scale = np.maximum(np.abs(want), 1e-13 * np.abs(want).max(axis=-1, keepdims=True))
    return float(np.max(np.abs(got - want) / np.where(scale == 0, 1, scale)))


def one_thread_blas():
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
