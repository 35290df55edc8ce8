"""Per-site and per-point cost of the n=8 sweep: time vs chain length M at 1e6 points (cfg2 geometry, longer barrier)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transport_b200 import _lib, configs
lib = _lib.load(); dev = torch.device("cuda", 0)
def dev_c(a): return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))).to(dev)
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 8
_lib.check(lib.tx_set_variant(variant))
if os.environ.get('TX_KMAX'): _lib.check(lib.tx_set_option(1, int(os.environ['TX_KMAX'])))
if os.environ.get('TX_GEXP'): _lib.check(lib.tx_set_option(2, int(os.environ['TX_GEXP'])))
Js, Es = configs.cfg2_grid(1000, 1000)
out = {}
for NB in (3, 11, 27):
    h0, h1, tnn, _ = configs.cicc_affine(1, 1.0, 0.0, 0.0, 0.0, NB, 2)
    M = h0.shape[0]
    d = [dev_c(x) for x in (h0, h1, tnn, Es)]
    dJ = torch.from_numpy(Js).to(dev)
    R = torch.empty((1000000, 8, 8), dtype=torch.float64, device=dev); T = torch.empty_like(R)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def step():
        _lib.check(lib.tx_transmission_sweep_dev(d[0].data_ptr(), 1, d[1].data_ptr(), dJ.data_ptr(), d[2].data_ptr(), 1, 8, M, 1000,
                   d[3].data_ptr(), 1000, 0, None, None, 1, 1, None, None, 8, R.data_ptr(), T.data_ptr(), None, None, flag.data_ptr(), st))
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): step()
    b.record(); torch.cuda.synchronize()
    out[M] = a.elapsed_time(b) / 5
Ms = sorted(out)
slope = (out[Ms[-1]] - out[Ms[0]]) / (Ms[-1] - Ms[0])
print(json.dumps({"variant": variant, "ms_by_M": out, "ms_per_site": slope, "ms_per_point_overhead": out[Ms[0]] - slope * Ms[0],
                  "frac_of_37TF_at_large_M": 2 * 8 * 512 * 1e6 / (slope * 1e-3) / 37e12}))
