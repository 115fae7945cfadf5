"""What the non-walk phases of the pair-symmetric census kernel cost: the fused masked call (census_mse and census_sad) with
parts switched off through the debug option census_sym_dbg (results are wrong in these runs; timing only).
python tools/experiments/census_sym_phases.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
from bench_ops import timeit
B, H, W = 8, 480, 640
dev = torch.device("cuda", 0)
base = synth.make_batch(B, H, W)
NS = 5
sets = []
for s in range(NS):
    d = {k: torch.from_numpy(np.ascontiguousarray(np.roll(base[k], 5 * s, axis=2))).to(dev) for k in ("es", "ta", "go", "std")}
    d["o1"], d["o2"], d["sums"] = torch.empty(B, 1, H, W, device=dev), torch.empty(B, 1, H, W, device=dev), torch.zeros(2, device=dev)
    sets.append(d)
_lib.set_option("census_sym", 1)
res = {}
for ty in (2, 3):
    def fused(i, st):
        d = sets[i % NS]
        _lib.call("ctd_photometric_fwd_bwd_masked_f32", d["es"].data_ptr(), d["ta"].data_ptr(), d["go"].data_ptr(), d["std"].data_ptr(),
                  d["o1"].data_ptr(), d["o2"].data_ptr(), d["sums"].data_ptr(), B, 1, H, W, 9, ty, 0.5, st)
    def fwd(i, st):
        d = sets[i % NS]
        _lib.call("ctd_photometric_fwd_f32", d["es"].data_ptr(), d["ta"].data_ptr(), d["o1"].data_ptr(), B, 1, H, W, 9, ty, 0.5, st)
    for name, dbg in (("all", 0), ("no_band", 4), ("no_band_no_loads", 5), ("no_band_no_stores", 6), ("no_band_no_sums", 12), ("walk_only", 15)):
        _lib.set_option("census_sym_dbg", dbg)
        res["t%d_fused_%s" % (ty, name)] = round(timeit(fused, 8)[0] * 1e3, 1)
        if ty == 3:
            res["t3_fwd_%s" % name] = round(timeit(fwd, 8)[0] * 1e3, 1)
    _lib.set_option("census_sym_dbg", 0)
_lib.set_option("census_sym_noguard", 1)
_lib.set_option("census_sym_dbg", 15)
res["t3_fused_walk_only_noguard"] = round(timeit(fused, 8)[0] * 1e3, 1)
print(json.dumps(res, indent=1))
