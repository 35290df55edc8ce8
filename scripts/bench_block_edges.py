"""Device time of the large-block sweep (closed-form lead tables, scalar hopping, telescoped) per block edge n: the
reference's two-impurity chains have n = 2 (2s+1)^2 = 8, 18, 32, 50, 72, 98, 128 for s = 1/2 ... 7/2, so half of the spin
values give a block edge that is not a multiple of the 8 x 8 tensor-core tile.

    python scripts/bench_block_edges.py [n ...] > profiles/r2_bench_block_edges.jsonl
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import benchutil as bu  # noqa: E402
from transport_b200 import _lib, leads  # noqa: E402


def chain(n, M, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, n, n)) + 1j * rng.standard_normal((M, n, n))
    h = 0.2 * (A + np.conj(np.swapaxes(A, 1, 2))) / 2 / np.sqrt(n / 8)
    h[0] = 0
    h[-1] = 0
    tnn = -np.tile(np.eye(n, dtype=complex), (M - 1, 1, 1))
    return h, tnn


def main(edges):
    env = bu.Env()
    torch, lib = env.torch, env.lib
    M, nE = 8, 1184
    Es = np.linspace(-1.9, 1.9, nE).astype(complex)
    # TX_BENCH_DIAG_SITES=1: only sites 2 and 5 keep a dense block (two impurities with spacer sites, as get_hblocks
    # builds them); the others are diagonal and become row scalings in the telescoped sweep
    spacer = bool(os.environ.get("TX_BENCH_DIAG_SITES"))
    if os.environ.get("TX_BENCH_SCAN") is not None:      # A/B: structure scan (diagonal sites as row scalings) on / off
        _lib.check(lib.tx_set_option(0, int(os.environ["TX_BENCH_SCAN"])))
    for n in edges:
        h, tnn = chain(n, M, n)
        if spacer:
            for j in range(1, M - 1):
                if j not in (2, 5):
                    h[j] = np.diag(np.real(np.diagonal(h[j]))) * 3
        SL, SR = leads.self_energies_batched(h, tnn, Es, "g_closed")
        d_h, d_t, d_E = env.dev_c(h), env.dev_c(tnn), env.dev_c(Es)
        sigL, sigR = env.dev_c(SL), env.dev_c(SR)
        tot = torch.empty(nE, dtype=torch.float64, device=env.dev)
        flag = torch.zeros(1, dtype=torch.int32, device=env.dev)
        wse = int(lib.tx_sweep_workspace_elems(n, nE))
        work = torch.empty((max(wse, 1), 2), dtype=torch.float64, device=env.dev)
        st = env.stream.cuda_stream

        def step():
            _lib.check(lib.tx_transmission_sweep_large_dev(d_h.data_ptr(), 1, None, None, d_t.data_ptr(), 1, n, M, 1,
                                                           d_E.data_ptr(), nE, sigL.data_ptr(), sigR.data_ptr(), 1,
                                                           _lib.TX_HOP_SCALAR, None, n, None, None, None, None,
                                                           tot.data_ptr(), work.data_ptr() if wse else None, flag.data_ptr(), st))
        ms, _ = env.timed(step, 3, 3)
        assert int(flag.item()) == 0
        from oracle import rgf_oracle
        idx = [0, nE // 3, nE - 1]
        _, oT = rgf_oracle.transmission_closed(h, tnn, Es[idx], np.eye(n))
        want = oT.sum(axis=(1, 2))
        got = tot.cpu().numpy()[idx]
        flops = 8 * n ** 3 * M * 2
        print(json.dumps({"n": n, "M": M, "nE": nE, "diagonal_spacer_sites": spacer, "structure_scan": os.environ.get("TX_BENCH_SCAN", "1"), "ms": ms, "points_per_s": nE / (ms * 1e-3),
                          "fp64_frac_algorithmic": flops * nE / (ms * 1e-3) / 1e12 / env.fp64_peak(),
                          "caroli_trace_max_rel_err_vs_oracle_total_T": float(np.max(np.abs(got - want) / np.abs(want)))}), flush=True)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [18, 24, 32, 50, 56, 72, 98, 104, 128])
