import sys
import os; R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tools"))
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
from bench_ops import timeit
B,H,W=8,480,640
dev=torch.device("cuda",0)
base=synth.make_batch(B,H,W)
NS=5
sets=[]
for s in range(NS):
    d={k: torch.from_numpy(np.ascontiguousarray(np.roll(base[k],5*s,axis=2))).to(dev) for k in ("es","ta","go","std")}
    d["o1"]=torch.empty(B,1,H,W,device=dev); d["o2"]=torch.empty(B,1,H,W,device=dev); sets.append(d)
sums=torch.zeros(2,device=dev)
def f(i,st):
    d=sets[i%NS]
    _lib.call("ctd_photometric_fwd_bwd_f32", d["es"].data_ptr(), d["ta"].data_ptr(), d["go"].data_ptr(), d["o1"].data_ptr(), d["o2"].data_ptr(), B,1,H,W,9,3,0.5,st)
print("fused census, no masked sums: %.1f us" % (timeit(f,30)[0]*1000))
def g(i,st):
    d=sets[i%NS]
    _lib.call("ctd_photometric_fwd_bwd_masked_f32", d["es"].data_ptr(), d["ta"].data_ptr(), d["go"].data_ptr(), d["std"].data_ptr(), d["o1"].data_ptr(), d["o2"].data_ptr(), sums.data_ptr(), B,1,H,W,9,3,0.5,st)
print("fused census, masked sums:    %.1f us" % (timeit(g,30)[0]*1000))
