import sys, time, ctypes
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
B, H, W = 8, 480, 640
base = synth.make_batch(B, H, W)
h = {k: torch.from_numpy(np.ascontiguousarray(base[k])).pin_memory() for k in ("im", "es", "ta", "go")}
for k in ("lcn", "std", "o", "g"): h[k] = torch.empty(B, 1, H, W).pin_memory()
P = lambda t: ctypes.c_void_p(t.data_ptr())
def lcn(): _lib.call("ctd_host_lcn_f32", P(h["im"]), P(h["lcn"]), P(h["std"]), B, H, W, 5, 0.05)
def ph(ty): _lib.call("ctd_host_photometric_fwd_bwd_f32", P(h["es"]), P(h["ta"]), P(h["go"]), P(h["o"]), P(h["g"]), B, 1, H, W, 9, ty, 0.5)
for nch in (1, 2, 4, 8):
    _lib.set_option("host_chunks", nch)
    for name, f in (("lcn", lcn), ("sad", lambda: ph(1)), ("census_sad", lambda: ph(3))):
        for _ in range(3): f()
        t = time.perf_counter()
        for _ in range(10): f()
        dt = (time.perf_counter() - t) / 10
        print("chunks", nch, name, "%.3f ms" % (dt * 1e3))
