#!/usr/bin/env python
"""Minimal launch sequence for ncu: one warm-up round and one profiled round of the bench chain's
kernels at the bench size (batch 8, 480x640).  `ncu -s <launches per round> -c <launches per round>`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from connecting_the_dots_b200 import _lib, synth  # noqa: E402

H, W, B = 480, 640, 8
extra = sys.argv[1:]
dev = torch.device("cuda", 0)
base = synth.make_batch(B, H, W)
d = {k: torch.from_numpy(base[k]).to(dev) for k in ("im", "es", "ta", "go", "std", "pat_lcn")}
o = [torch.empty(B, 1, H, W, device=dev) for _ in range(6)]
ws = torch.zeros(int(_lib.lib().ctd_masked_sums_workspace_bytes()), dtype=torch.uint8, device=dev)
sums = torch.zeros(2, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = {k: v.data_ptr() for k, v in d.items()}
vol = torch.empty(2, 128, H, W, device=dev) if "xcorrvol" in extra else None
for rnd in range(2):
    _lib.call("ctd_lcn_f32", p["im"], o[0].data_ptr(), o[1].data_ptr(), B, H, W, 5, 0.05, st)
    _lib.call("ctd_photometric_fwd_f32", p["es"], p["ta"], o[2].data_ptr(), B, 1, H, W, 9, 1, 0.5, st)
    _lib.call("ctd_photometric_bwd_f32", p["es"], p["ta"], p["go"], o[3].data_ptr(), B, 1, H, W, 9, 1, 0.5, st)
    _lib.call("ctd_photometric_fwd_f32", p["es"], p["ta"], o[4].data_ptr(), B, 1, H, W, 9, 3, 0.5, st)
    _lib.call("ctd_photometric_bwd_f32", p["es"], p["ta"], p["go"], o[5].data_ptr(), B, 1, H, W, 9, 3, 0.5, st)
    _lib.call("ctd_masked_sums_f32", o[4].data_ptr(), p["std"], B * H * W, sums.data_ptr(), ws.data_ptr(), st)
    if vol is not None:
        _lib.call("ctd_xcorrvol_f32", p["ta"], p["pat_lcn"], vol.data_ptr(), 2, 1, H, W, 128, 9, st)
    torch.cuda.synchronize()
print("ok", float(sums[0] / sums[1]))
