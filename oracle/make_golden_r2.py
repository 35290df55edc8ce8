"""
TEST INFRASTRUCTURE.  Round-2 additions to tests/golden/: outputs of the UNMODIFIED reference (imported from
/root/reference through oracle/ref_loader.py).  Runs only in the build container.

    python -m oracle.make_golden_r2 [name ...]
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=SyntaxWarning)

from oracle.make_golden import sweep  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402
from transport_b200 import configs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cicc_spin32(ref):
    """Two spin-3/2 impurities: n_loc_dof = 2 (2s+1)^2 = 32, the large-block kernel's size class, straight from
    run_wfm_gate.get_hblocks (run_wfm_gate.py:80-143)."""
    wfm, gate = ref["wfm"], ref["run_wfm_gate"]
    h, tnn, tnnn = gate.get_hblocks(3, 1.0, -0.3, 0.05, 0.0, 0.0, 1, the_dist=1)
    hb, tb, _ = configs.cicc_blocks(3, 1.0, -0.3, 0.05, 0.0, 0.0, 1, 1)
    assert np.allclose(h, hb, atol=0, rtol=0) and np.array_equal(tnn, tb), "cicc builder deviates from reference"
    assert h.shape[-1] == 32
    Es = np.array([-1.95, -1.4, -0.6], dtype=complex)
    srcs = np.eye(32)[[0, 13, 31]]
    R, T = sweep(wfm, h, tnn, tnnn, 1.0, Es, "g_closed", list(srcs), all_debug=True)
    np.savez(os.path.join(OUT, "cicc_spin32_n32.npz"), h=h, tnn=tnn, tnnn=tnnn, Es=Es, sources=srcs, R=R, T=T)


MAKERS = {"cicc_spin32": cicc_spin32}


def main(names):
    ref = load_reference(with_bardeen=True, with_scripts=True)
    for name in names or sorted(MAKERS):
        MAKERS[name](ref)
        print("wrote", name)


if __name__ == "__main__":
    main(sys.argv[1:])
