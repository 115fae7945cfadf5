import sys
import os; R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
from connecting_the_dots_b200 import _lib
from bench_ops import timeit
dev=torch.device("cuda",0)
for (B,H,W) in ((8,480,640),(64,480,640),(1,4096,4096),(1,4096,640),(40,96,640)):
    g=torch.Generator(device=dev).manual_seed(1)
    es=torch.randn(B,1,H,W,device=dev,generator=g); ta=torch.randn(B,1,H,W,device=dev,generator=g); go=torch.rand(B,1,H,W,device=dev,generator=g)
    o1=torch.empty_like(es); o2=torch.empty_like(es)
    def f(i,st):
        _lib.call("ctd_photometric_fwd_bwd_f32", es.data_ptr(), ta.data_ptr(), go.data_ptr(), o1.data_ptr(), o2.data_ptr(), B,1,H,W,9,3,0.5,st)
    ms=timeit(f,10,warmup=2,nrep=3)[0]
    px=B*H*W
    tiles=B*((H+15)//16)*((W+63)//64)
    border=B*(2*((W+63)//64)+2*((H+15)//16)-4)
    print("%dx%dx%d: %.1f us, %.2f cycles per warp-tap, border tiles %.0f%%, waves %.2f" % (B,H,W,ms*1000, ms*1e-3*1.965e9*592/(px*81/32), 100.0*border/tiles, tiles/444.0))
