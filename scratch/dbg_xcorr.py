import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
import oracle
import connecting_the_dots_b200 as ctd
from connecting_the_dots_b200 import _lib, synth
tx = ctd.torchext
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
rng = np.random.RandomState(7)
H, W, D = 40, 132, 21
z = rng.randn(1, 1, H, W).astype(np.float32)
z[0, 0, :, 40:90] = 0.0
z2 = np.roll(z, 2, axis=3)
for bs in (5, 9, 3, 7):
    got = tx.xcorrvol(cu(z), cu(z2), D, bs).cpu().numpy()[0]
    ref = oracle.xcorrvol(z[0], z2[0], D, bs)
    err = np.abs(got - ref)
    idx = np.unravel_index(np.argmax(err), err.shape)
    print("bs", bs, "maxerr", err.max(), "at", idx, "got", got[idx], "ref", ref[idx], "n>1e-5:", (err > 1e-5).sum())
    bad = np.argwhere(err > 1e-5)
    if len(bad):
        print(" d range", bad[:, 0].min(), bad[:, 0].max(), "h", bad[:, 1].min(), bad[:, 1].max(), "w", bad[:, 2].min(), bad[:, 2].max())
# timing
d = synth.make_batch(2, 480, 640)
a, b = cu(d["ta"]), cu(d["pat_lcn"])
for bs in (9, 5):
    for it in range(3):
        torch.cuda.synchronize(); t = time.time()
        o = tx.xcorrvol(a, b, 128, bs)
        torch.cuda.synchronize(); print("bs", bs, "B2 D128 ms", (time.time() - t) * 1e3)
