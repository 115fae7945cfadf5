import sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np, torch
from scipy.ndimage import uniform_filter
import connecting_the_dots_b200 as ctd
from connecting_the_dots_b200 import synth, _lib
tx = ctd.torchext
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
H, W, D, BS = 480, 640, 128, (int(sys.argv[1]) if len(sys.argv) > 1 else 9)
R = BS // 2
for n in (0, 3):
    d = synth.make_pair(n, H, W)
    a = d["ta"].astype(np.float64); b = d["pat_lcn"].astype(np.float64)
    _lib.set_option("xcorr_nofix", 1)
    fast = tx.xcorrvol(cu(d["ta"][None]), cu(d["pat_lcn"][None]), D, BS).cpu().numpy()
    _lib.set_option("xcorr_nofix", 0)
    full = tx.xcorrvol(cu(d["ta"][None]), cu(d["pat_lcn"][None]), D, BS).cpu().numpy()
    ap = np.pad(a, R, mode="edge")
    bp = np.pad(np.pad(b, ((0, 0), (D - 1, 0)), mode="edge"), R, mode="edge")   # columns u = -(D-1) ..
    def box(x):  # BS x BS sums, valid
        c = np.cumsum(np.cumsum(np.pad(x, ((1, 0), (1, 0))), 0), 1)
        return c[BS:, BS:] - c[:-BS, BS:] - c[BS:, :-BS] + c[:-BS, :-BS]
    N = BS * BS
    sa, saa = box(ap), box(ap * ap)
    sb, sbb = box(bp), box(bp * bp)          # [H, W + D - 1], index u + D - 1
    va = np.maximum(saa - sa * sa / N, 0); vb = np.maximum(sbb - sb * sb / N, 0)
    ra = va / np.maximum(saa, 1e-300); rb = vb / np.maximum(sbb, 1e-300)
    La = np.clip(np.floor(-4 * np.log2(np.maximum(ra, 1e-300))), 0, 255); Lb = np.clip(np.floor(-4 * np.log2(np.maximum(rb, 1e-300))), 0, 255)
    bins = np.zeros(80); cnt = np.zeros(80); emax_full = 0
    for dd in range(D):
        bsh = bp[:, D - 1 - dd: D - 1 - dd + W + 2 * R]
        sab = box(ap * bsh)
        sl = slice(D - 1 - dd, D - 1 - dd + W)
        dot = sab - sa * sb[:, sl] / N
        ref = dot / (np.sqrt(va * vb[:, sl]) + 1e-8)
        err = np.abs(fast[dd] - ref)
        Ls = (La + Lb[:, sl]).astype(int).clip(0, 79)
        np.maximum.at(bins, Ls.ravel(), err.ravel()); np.add.at(cnt, Ls.ravel(), 1)
        emax_full = max(emax_full, np.abs(full[dd] - ref).max())
    print("image", n, "max err with fix-up:", emax_full)
    for L in range(0, 80, 4):
        print("  L0+L1 in [%d,%d): count %9d  max fast-path err %.2e" % (L, L + 4, cnt[L:L + 4].sum(), bins[L:L + 4].max()))
