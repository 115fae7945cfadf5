import sys, time, ctypes
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
B, H, W = 8, 480, 640
base = synth.make_batch(B, H, W)
h = {k: torch.from_numpy(np.ascontiguousarray(base[k])).pin_memory() for k in ("im", "es", "ta", "go")}
for k in ("lcn", "std", "o", "g", "o2", "g2"): h[k] = torch.empty(B, 1, H, W).pin_memory()
P = lambda t: ctypes.c_void_p(t.data_ptr())
def step(defer):
    if defer: _lib.call("ctd_host_begin_batch")
    _lib.call("ctd_host_lcn_f32", P(h["im"]), P(h["lcn"]), P(h["std"]), B, H, W, 5, 0.05)
    _lib.call("ctd_host_photometric_fwd_bwd_f32", P(h["es"]), P(h["ta"]), P(h["go"]), P(h["o"]), P(h["g"]), B, 1, H, W, 9, 1, 0.5)
    _lib.call("ctd_host_photometric_fwd_bwd_f32", P(h["es"]), P(h["ta"]), P(h["go"]), P(h["o2"]), P(h["g2"]), B, 1, H, W, 9, 3, 0.5)
    if defer: _lib.call("ctd_host_end_batch")
for defer in (0, 1):
    for nch in (1, 2, 4, 8):
        _lib.set_option("host_chunks", nch)
        for _ in range(3): step(defer)
        t = time.perf_counter()
        for _ in range(20): step(defer)
        dt = (time.perf_counter() - t) / 20
        print("deferred", defer, "chunks", nch, "%.3f ms" % (dt * 1e3))
