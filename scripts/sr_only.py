"""Profiling helper: two Sancho-Rubio launches (n = 64 ribbon lead, 592 energies = two waves of resident CTAs).
    ncu --set full --import-source on -k regex:surface_gf_kernel -s 1 -c 1 -o prof python scripts/sr_only.py"""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from transport_b200 import _lib, configs
lib = _lib.load()
dev = torch.device("cuda", 0)
n, nE = 64, 592
h, tnn, _ = configs.ribbon_blocks(width=n, L=32, W=1.0, seed=1234)
Es = np.linspace(-3.5, 3.5, nE).astype(complex)
c = lambda a: torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a, dtype=np.complex128))).to(dev)
d_h, d_t, d_E = c(h), c(tnn), c(Es)
sig = torch.empty((nE, n, n, 2), dtype=torch.float64, device=dev)
nit = torch.zeros(nE, dtype=torch.int32, device=dev); ok = torch.zeros_like(nit)
conv = torch.zeros(nE, dtype=torch.float64, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(lib.tx_surface_gf_dev(d_h.data_ptr(), d_t.data_ptr(), n, d_E.data_ptr(), nE, -1, _lib.TX_GF_SANCHO, 1e-8, 1e-14, 0, 0, 200,
                                     None, sig.data_ptr(), nit.data_ptr(), conv.data_ptr(), ok.data_ptr(), flag.data_ptr(), st))
torch.cuda.synchronize()
print("ok", int(ok.sum()), float(nit.double().mean()))
