"""ncu target: a few launches of the fused masked census_sad call (pair-symmetric kernel) at batch 8, 480x640."""
import os, sys
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
for a in sys.argv[3:]:  # option or option=value (e.g. census_sym=0: the gather kernel)
    _lib.set_option(a.split("=")[0], int(a.split("=")[1]) if "=" in a else 1)
for _ in range(3):
    _lib.call("ctd_photometric_fwd_bwd_masked_f32", d["es"].data_ptr(), d["ta"].data_ptr(), d["go"].data_ptr(), d["std"].data_ptr(),
              o1.data_ptr(), o2.data_ptr(), sums.data_ptr(), B, 1, H, W, 9, ty, 0.5, st)
torch.cuda.synchronize()
print(sums.tolist())
