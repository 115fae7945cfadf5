import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import connecting_the_dots_b200 as ctd
from connecting_the_dots_b200 import synth
tx = ctd.torchext
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
d = synth.make_batch(2, 480, 640)
a, b = cu(d["ta"]), cu(d["pat_lcn"])
for bs in (9, 5):
    for it in range(2):
        o = tx.xcorrvol(a, b, 128, bs)
torch.cuda.synchronize()
# conditioning of the bench data, numpy: ratio var / S2 over 9x9 windows
x = d["ta"][0, 0].astype(np.float64)
from numpy.lib.stride_tricks import sliding_window_view as swv
for name, im in (("ta", d["ta"][0, 0]), ("pat_lcn", d["pat_lcn"][0, 0])):
    w = swv(im.astype(np.float64), (9, 9))
    s1 = w.sum((2, 3)); s2 = (w * w).sum((2, 3))
    var = s2 - s1 * s1 / 81
    print(name, "flag frac", float((var < 0.1 * s2).mean()), "min ratio", float((var / s2).min()), "mean", float(im.mean()), "std", float(im.std()))
