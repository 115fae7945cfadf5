"""The bench's end-to-end step (host buffers through the C ABI) with different chunk counts inside the batch and
different call orders, issued the ordinary way and replayed as a cached CUDA graph: [ms per step, of which host time in the
calls before ctd_host_end_batch].  python tools/experiments/e2e_chunks.py"""
import os, sys, time, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
B, H, W = 8, 480, 640
torch.cuda.set_device(0)
base = synth.make_batch(B, H, W)
NS = 4
sets = []
for s in range(NS):
    h = {k: torch.from_numpy(np.ascontiguousarray(np.roll(base[k], 3 * s, axis=2))).pin_memory() for k in ("im", "es", "ta", "go")}
    for k in ("lcn", "std", "gi_sad", "gi_cs"):
        h[k] = torch.empty(B, 1, H, W).pin_memory()
    h["sums"] = torch.zeros(2, 2).pin_memory()
    sets.append(h)
P = lambda t: ctypes.c_void_p(t.data_ptr())
enq = [0.0]
def step(k, order):
    h = sets[k % NS]
    t_in = time.perf_counter()
    _lib.call("ctd_host_begin_batch")
    _lib.call("ctd_host_lcn_f32", P(h["im"]), P(h["lcn"]), P(h["std"]), B, H, W, 5, 0.05)
    for ty, gi, off in (order):
        _lib.call("ctd_host_photometric_fwd_bwd_masked_f32", P(h["es"]), P(h["ta"]), P(h["go"]), P(h["std"]), None, P(h[gi]),
                  ctypes.c_void_p(h["sums"].data_ptr() + off), B, 1, H, W, 9, ty, 0.5)
    enq[0] += time.perf_counter() - t_in   # host time in the calls, before the wait in ctd_host_end_batch
    _lib.call("ctd_host_end_batch")
res = {}
for name, order in (("sad_then_census", ((1, "gi_sad", 0), (3, "gi_cs", 8))), ("census_then_sad", ((3, "gi_cs", 8), (1, "gi_sad", 0)))):
    for graphs, nch in ((0, 1), (0, 2), (0, 4), (1, 1), (1, 2), (1, 4)):
        _lib.lib().ctd_host_release()  # forget cached batches
        _lib.set_option("host_graphs", graphs)
        _lib.set_option("host_chunks", 8)
        _lib.set_option("host_chunks_batch", nch)
        _lib.set_option("host_chunks_graph", nch)
        for k in range(3 * NS):
            step(k, order)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 30
        enq[0] = 0.0
        for k in range(n):
            step(k, order)
        res["%s_%s_chunks%d" % (name, "graph" if graphs else "eager", nch)] = [round((time.perf_counter() - t0) / n * 1e3, 4), round(enq[0] / n * 1e3, 4)]
print(json.dumps(res, indent=1))
