import sys, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np, torch
import connecting_the_dots_b200 as ctd
from connecting_the_dots_b200 import _lib, synth
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from bench_ops import timeit
B,H,W=8,480,640
dev=torch.device("cuda",0)
gd = synth.make_depth_pairs(B, H, W, seed=0)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ray = ctd.torchext.projection_rays(gd["Ki"], H, W).to(dev)
g = {k: cu(v) for k, v in gd.items()}
NS=5
deps = [(cu(np.roll(gd["depth0"], 5 * s, axis=3)), cu(np.roll(gd["depth1"], 5 * s, axis=3))) for s in range(NS)]
g0s = [torch.empty(B, 1, H, W, device=dev) for _ in range(NS)]
g1s = [torch.zeros(B, 1, H, W, device=dev) for _ in range(NS)]
sums = torch.zeros(2, 2, device=dev)
for name, gb, ga in (("both", True, True), ("no_scatter", False, True), ("no_direct", True, False), ("value_only", False, False)):
    def f(i, st):
        d0, d1 = deps[i % NS]
        _lib.call("ctd_depth_similarity_f32", d0.data_ptr(), d1.data_ptr(), ray.data_ptr(), g["K"].data_ptr(), g["R0"].data_ptr(),
                  g["t0"].data_ptr(), g["R1"].data_ptr(), g["t1"].data_ptr(), g0s[i % NS].data_ptr() if ga else 0, g1s[i % NS].data_ptr() if gb else 0, sums[0].data_ptr(),
                  B, H, W, 0.1, 1e-6, 0, st)
    print(name, "one direction us:", round(timeit(f, 20)[0]*1000,1))
