"""Per-warp timeline of the pair-symmetric census kernel (debug hook ctd_debug_census_timeline): where the time of a
launch goes -- staging / walk / barrier waits / carry / exact pass / write-out, band CTAs, tail.
python tools/experiments/census_sym_timeline.py [batch] [type]"""
import ctypes, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 480, 640
ty = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
base = synth.make_batch(min(B, 8), H, W)
d = {k: torch.from_numpy(np.ascontiguousarray(base[k])).to(dev) for k in ("es", "ta", "go", "std")}
o1, o2, sums = torch.empty(B, 1, H, W, device=dev), torch.empty(B, 1, H, W, device=dev), torch.zeros(2, device=dev)
st = torch.cuda.current_stream().cuda_stream
_lib.set_option("census_sym", 1)
L = _lib.lib()
L.ctd_debug_census_timeline.argtypes = [ctypes.c_void_p]
L.ctd_debug_census_timeline.restype = None
def run():
    _lib.call("ctd_photometric_fwd_bwd_masked_f32", d["es"].data_ptr(), d["ta"].data_ptr(), d["go"].data_ptr(), d["std"].data_ptr(),
              o1.data_ptr(), o2.data_ptr(), sums.data_ptr(), B, 1, H, W, 9, ty, 0.5, st)
for _ in range(3):
    run()
torch.cuda.synchronize()
buf = torch.zeros(20000 * 128, dtype=torch.int64, device=dev)
L.ctd_debug_census_timeline(ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
L.ctd_debug_census_timeline(None)
t = buf.cpu().numpy().reshape(-1, 16, 8).astype(np.float64)
n = int((t[:, 0, 0] > 0).sum())
t = t[:n]
t0 = t[:, 0, 0].min()
is_tile = t[:, 0, 2] > 0
main, band = t[is_tile], t[~is_tile]
nw = int((main[0, :, 0] > 0).sum())
main = main[:, :nw]
us = lambda a: a / 1e3
st_ = lambda v: {"mean": round(float(us(v).mean()), 2), "p10": round(float(np.percentile(us(v), 10)), 2), "p90": round(float(np.percentile(us(v), 90)), 2), "max": round(float(us(v).max()), 2)}
out = {"ctas": n, "tile_ctas": int(is_tile.sum()), "band_ctas": int((~is_tile).sum()), "warps_per_cta": nw,
       "kernel_us": round(float(us(t[:, :, 4].max() - t0)), 2)}
# slots: 0 start, 1 staged (after barrier), 5 walk end (before barrier 1), 2 after barrier 1, 6 carry end (before barrier 2), 3 after barrier 2 + exact, 4 end
out["per_warp_us"] = {
    "stage (start -> after staging barrier)": st_(main[:, :, 1] - main[:, :, 0]),
    "walk (own)": st_(main[:, :, 5] - main[:, :, 1]),
    "wait at barrier 1": st_(main[:, :, 2] - main[:, :, 5]),
    "carry (own)": st_(main[:, :, 6] - main[:, :, 2]),
    "barrier 2 + exact pass": st_(main[:, :, 3] - main[:, :, 6]),
    "write-out + finish": st_(main[:, :, 4] - main[:, :, 3]),
    "total": st_(main[:, :, 4] - main[:, :, 0])}
out["walk_us_by_warp"] = [round(float(us(main[:, w, 5] - main[:, w, 1]).mean()), 2) for w in range(nw)]
if len(band):
    out["band_cta_us"] = {"mean": round(float(us(band[:, 0, 4] - band[:, 0, 0]).mean()), 2), "first_start": round(float(us(band[:, 0, 0].min() - t0)), 2),
                          "last_end": round(float(us(band[:, 0, 4].max() - t0)), 2)}
out["tile_last_end_us"] = round(float(us(main[:, :, 4].max() - t0)), 2)
out["tile_start_us_percentiles"] = [round(float(us(np.percentile(main[:, 0, 0], p) - t0)), 2) for p in (0, 25, 50, 51, 75, 100)]
print(json.dumps(out, indent=1))
