import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np, torch
import connecting_the_dots_b200 as ctd
from connecting_the_dots_b200 import synth
tx = ctd.torchext
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d = synth.make_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 2, 480, 640)
a, b = cu(d["ta"]), cu(d["pat_lcn"])
for it in range(2):
    o = tx.xcorrvol(a, b, 128, int(sys.argv[2]) if len(sys.argv) > 2 else 9)
torch.cuda.synchronize()
