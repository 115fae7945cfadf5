"""XCorrVol block 5 at the bench size with different capacities of the fix-up's hit list (default: 1/32 of the outputs, at
most 4 M entries -> overflows at block 5, where 2.5 % of the outputs are ill-conditioned, and the fallback kernel runs)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
from bench_ops import timeit
B, H, W, D = 8, 480, 640, 128
dev = torch.device("cuda", 0)
d = synth.make_batch(B, H, W)
a = torch.from_numpy(d["ta"][:, 0].copy()).to(dev); b = torch.from_numpy(d["pat_lcn"][:, 0].copy()).to(dev)
outs = [torch.empty(B, D, H, W, device=dev) for _ in range(2)]
res = {}
ref = None
for bs in (5, 9):
    for cap in (-1, 6 << 20, 10 << 20, 16 << 20):
        _lib.set_option("xcorr_hitcap", cap)
        def f(i, st):
            _lib.call("ctd_xcorrvol_f32", a.data_ptr(), b.data_ptr(), outs[i % 2].data_ptr(), B, 1, H, W, D, bs, st)
        res["bs%d_cap%d" % (bs, cap)] = round(timeit(f, 6)[0] * 1e3, 1)
        torch.cuda.synchronize()
        if cap == -1:
            ref = outs[0].clone()
        else:
            res["bs%d_cap%d_maxdiff" % (bs, cap)] = float((outs[0] - ref).abs().max())
_lib.set_option("xcorr_hitcap", -1)
print(json.dumps(res, indent=1))
