import os, sys, json, numpy as np, torch
sys.path.insert(0, '/root/repo')
from transport_b200 import _lib, configs
lib=_lib.load(); dev=torch.device('cuda',0)
def dev_c(a): return torch.view_as_real(torch.from_numpy(np.ascontiguousarray(a,dtype=np.complex128))).to(dev)
for n in (4,8,16):
  for hop,(noise) in (("scalar",0.0),("general",0.05)):
    h,tnn,_=configs.disordered_blocks(N=62,n=n,W=0.05,seed=1,hop_noise=noise)
    M=h.shape[0]; nE=200000 if n<=8 else 50000
    Es=np.linspace(-1.9,1.9,nE).astype(complex)
    d=[dev_c(x) for x in (h,tnn,Es)]
    R=torch.empty((nE,n,n),dtype=torch.float64,device=dev); T=torch.empty_like(R); flag=torch.zeros(1,dtype=torch.int32,device=dev)
    st=torch.cuda.current_stream().cuda_stream
    mode=_lib.TX_HOP_SCALAR if hop=="scalar" else _lib.TX_HOP_GENERAL
    def step(): _lib.check(lib.tx_transmission_sweep_dev(d[0].data_ptr(),1,None,None,d[1].data_ptr(),1,n,M,1,d[2].data_ptr(),nE,0,None,None,1,mode,None,None,n,R.data_ptr(),T.data_ptr(),None,None,flag.data_ptr(),st))
    for _ in range(2): step()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); 
    for _ in range(3): step()
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/3
    g=1 if hop=="scalar" else 4
    fl=8*n**3*M*(1+g)*nE
    print(json.dumps({"n":n,"hop":hop,"ms":ms,"pts_per_s":nE/ms*1e3,"tflops":fl/ms/1e9,"frac":fl/ms/1e9/37.0}))
