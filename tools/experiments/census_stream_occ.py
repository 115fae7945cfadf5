"""Prints the occupancy the driver grants census_stream_kernel (option census_stream = 2) for one fused call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from connecting_the_dots_b200 import _lib, synth
B, H, W = 8, 480, 640
dev = torch.device("cuda", 0)
base = synth.make_batch(B, H, W)
d = {k: torch.from_numpy(np.ascontiguousarray(base[k])).to(dev) for k in ("es", "ta", "go", "std")}
o1, o2, sums = torch.empty(B, 1, H, W, device=dev), torch.empty(B, 1, H, W, device=dev), torch.zeros(2, device=dev)
_lib.set_option("census_sym", 0)
_lib.set_option("census_stream", 2)
_lib.call("ctd_photometric_fwd_bwd_masked_f32", d["es"].data_ptr(), d["ta"].data_ptr(), d["go"].data_ptr(), d["std"].data_ptr(),
          o1.data_ptr(), o2.data_ptr(), sums.data_ptr(), B, 1, H, W, 9, 3, 0.5, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(sums.tolist())
