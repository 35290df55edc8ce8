"""The cfg4 workload (disordered square-lattice ribbon, Sancho-Rubio lead tables, telescoped sweep, Caroli trace) at ribbon
widths on both sides of the n = 64 boundary of the large-block kernels: 64 runs the 256-thread instantiations (two CTAs per
SM, TMA-staged operands), 72 ... 128 the 512-thread ones (one CTA per SM; inverse staging in shared memory up to n = 104,
in the L2-resident workspace beyond).  One JSON line per width: ms per pass, points/s, FP64 fraction per kernel, parity of
the Caroli trace against oracle/rgf_oracle on the device's own tables.

    python scripts/bench_wide_blocks.py [widths ...] > profiles/r2_bench_wide_blocks.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import benchutil as bu  # noqa: E402
import bench_configs as bc  # noqa: E402


def main(widths):
    env = bu.Env()
    for n in widths:
        out = bc.run_cfg4(env, steps=2, warmup=3, want_cpu=False, want_e2e=False, width=n, n_energies=1184, L=16)
        out["workload"] = "ribbon width %d, L = 16, 1184 energies, Sancho-Rubio leads (eta = 1e-8) + telescoped sweep + Caroli trace" % n
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [64, 72, 96, 104, 112, 128])
