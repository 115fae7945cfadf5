import sys, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
H, W = 480, 640
for B in (8, 64):
    base = synth.make_batch(min(B, 8), H, W)
    rep = B // min(B, 8)
    d = {k: torch.from_numpy(np.ascontiguousarray(np.tile(base[k], (rep, 1, 1, 1)))).cuda() for k in ("es", "ta", "go")}
    o = torch.empty(B, 1, H, W, device="cuda"); g = torch.empty(B, 1, H, W, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for pairs in (0, 1):
        _lib.set_option("census_pairs", pairs)
        for name, fn in (("fwd", lambda: _lib.call("ctd_photometric_fwd_f32", d["es"].data_ptr(), d["ta"].data_ptr(), o.data_ptr(), B, 1, H, W, 9, 3, 0.5, st)),):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            print("B", B, "pairs", pairs, name, "%.1f us" % (e0.elapsed_time(e1) * 100))
