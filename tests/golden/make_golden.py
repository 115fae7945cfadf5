"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

Sources of truth:
  * oracle/_ref/ctd_ref_ext_cpu.so -- /root/reference/torchext/ext/ext_cpu.cpp compiled as-is by
    oracle/build_ref.py (photometric fwd/bwd, xcorrvol, proj_nn, nn, crosscheck);
  * /root/reference/model/networks.py:507-533 -- the LCN class, exec'd from the reference file at
    run time on CPU torch with a stub for its TimedModule base (networks.py:10-23), because
    importing model.networks wholesale needs matplotlib (networks.py:4).
Inputs are seeded; every .npz stores inputs and the reference's outputs.  /root/reference is not on
the GPU box, so tests read only the committed .npz files.

    python tests/golden/make_golden.py
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

REF_NETWORKS = "/root/reference/model/networks.py"


def reference_lcn_class():
    src = open(REF_NETWORKS).read()
    tree = ast.parse(src)
    node = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LCN"][0]
    cls_src = ast.get_source_segment(src, node)

    class TimedModule(torch.nn.Module):  # stub of networks.py:10-23 without the device syncs/timer
        def __init__(self, mod_name):
            super().__init__()
            self.mod_name = mod_name

        def forward(self, *a, **k):
            return self.tforward(*a, **k)

    ns = {"torch": torch, "TimedModule": TimedModule}
    exec(compile(cls_src, REF_NETWORKS, "exec"), ns)
    return ns["LCN"]


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def main():
    build_ref.build()
    ref = build_ref.load_ref()
    assert ref is not None, "reference extension not built"
    torch.set_num_threads(1)

    # ---- photometric loss: 4 types x C in {1,2} x bs in {2,3,9}, fp32 (+ one fp64 set) ----
    rng = np.random.RandomState(1234)
    out = {}
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for C in (1, 2):
            B, H, W = 2, 12, 14
            es = rng.randn(B, C, H, W).astype(dt)
            ta = (es + 0.5 * rng.randn(B, C, H, W)).astype(dt)
            es[0, 0, 3, 4] = ta[0, 0, 3, 4]          # sign(0) = 0 in the sad gradient
            go = rng.rand(B, 1, H, W).astype(dt)
            out[f"{tag}_C{C}_es"], out[f"{tag}_C{C}_ta"], out[f"{tag}_C{C}_go"] = es, ta, go
            for bs in (2, 3, 9) if tag == "f32" else (9,):
                for ty in range(4):
                    eps = 0.5 if ty == 3 else 0.1
                    f = ref.photometric_loss_forward(t(es), t(ta), bs, ty, eps).numpy()
                    g = ref.photometric_loss_backward(t(es), t(ta), t(go), bs, ty, eps).numpy()
                    out[f"{tag}_C{C}_bs{bs}_t{ty}_fwd"] = f
                    out[f"{tag}_C{C}_bs{bs}_t{ty}_bwd"] = g
    np.savez_compressed(os.path.join(HERE, "photometric.npz"), **out)

    # ---- xcorrvol ----
    rng = np.random.RandomState(77)
    out = {}
    for C, H, W, D, bs in ((1, 10, 24, 6, 9), (2, 9, 20, 5, 3), (1, 8, 16, 20, 5), (1, 7, 12, 3, 2)):
        a = rng.rand(C, H, W).astype(np.float32)
        b = np.roll(a, 2, axis=2) + 0.1 * rng.randn(C, H, W).astype(np.float32)
        key = f"C{C}_H{H}_W{W}_D{D}_bs{bs}"
        out[key + "_in0"], out[key + "_in1"] = a, b.astype(np.float32)
        out[key + "_out"] = ref.xcorrvol_cpu(t(a), t(out[key + "_in1"]), D, bs).numpy()
    flat = np.full((1, 6, 8), 0.25, np.float32)
    out["flat_out"] = ref.xcorrvol_cpu(t(flat), t(flat), 3, 3).numpy()        # flat windows -> 0
    r = rng.rand(2, 6, 9).astype(np.float32)
    out["self_in"] = r
    out["self_out"] = ref.xcorrvol_cpu(t(r), t(r), 1, 3).numpy()              # == C everywhere
    np.savez_compressed(os.path.join(HERE, "xcorrvol.npz"), **out)

    # ---- proj_nn (edge cases of SURVEY.md appendix A.4) ----
    rng = np.random.RandomState(5)
    out = {}
    K = np.array([[56.76, 0, 16.2], [0, 57.02, 12.5], [0, 0, 1]], np.float32)
    B, H, W = 2, 24, 32
    z = 1.0 + rng.rand(B, H, W).astype(np.float32)
    u, v = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    ray = np.stack(((u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)), -1)
    xyz1 = (ray[None] * z[..., None]).astype(np.float32)
    xyz0 = (xyz1 + 0.02 * rng.randn(B, H, W, 3)).astype(np.float32)
    xyz0[0, 0, 0] = [1, 1, 0]              # u = +inf
    xyz0[0, 0, 1] = [0, 0, 0]              # 0/0
    xyz0[0, 0, 2] = [np.nan, 1, 1]         # NaN input
    xyz0[0, 0, 3] = [1e30, 1, 1e-8]        # |u| out of int range
    xyz0[0, 0, 4] = [-0.3, -0.25, 1.0]     # u+0.5 in (-1, 0): truncation toward zero
    xyz0[0, 0, 5] = [0.1, 0.1, -1.0]       # z < 0 is not rejected
    xyz0[1, 5, 5] = [5.0, 0.0, 1.0]        # projects outside the image
    xyz1[1, 3:6, 3:6] = 7.0                # ties -> first in scan order
    xyz0[1, 4, 4] = ray[4, 4] * 1.5       # projects onto (4,4); its 3x3 patch is all-equal
    out["K"], out["xyz0"], out["xyz1"] = K, xyz0, xyz1
    for ps in (1, 2, 3, 5, 9):
        out[f"ps{ps}"] = ref.proj_nn_cpu(t(xyz0), t(xyz1), t(K), ps).numpy()
    np.savez_compressed(os.path.join(HERE, "proj_nn.npz"), **out)

    # ---- nn ----
    rng = np.random.RandomState(9)
    out = {}
    p0 = rng.randn(300, 3).astype(np.float32)
    p1 = rng.randn(517, 3).astype(np.float32)
    p1[100] = p1[7]                        # tie -> lowest index
    p0[0] = p1[7]
    p0[1] = [np.nan, 0, 0]                 # NaN -> -1
    p0[2] = [3e4, 3e4, 3e4]                # every dist >= 1e9 -> -1
    out["p0"], out["p1"] = p0, p1
    out["idx"] = ref.nn_cpu(t(p0), t(p1)).numpy()
    out["idx_empty"] = ref.nn_cpu(t(p0[:5]), t(p1[:0])).numpy()
    np.savez_compressed(os.path.join(HERE, "nn.npz"), **out)

    # ---- crosscheck ----
    rng = np.random.RandomState(3)
    out = {}
    n = 400
    perm = rng.permutation(n).astype(np.int64)
    inv = np.empty(n, np.int64)
    inv[perm] = np.arange(n)
    i0, i1 = perm.copy(), inv.copy()
    i0[rng.rand(n) < 0.2] = -1
    i1[rng.rand(n) < 0.2] = -1
    bad = rng.rand(n) < 0.1
    i1[bad] = rng.randint(0, n, bad.sum())
    out["in0"], out["in1"] = i0, i1
    out["out"] = ref.crosscheck_cpu(t(i0), t(i1)).numpy()
    out["kat_in0"] = np.array([1, 0, -1, 3, 2], np.int64)
    out["kat_in1"] = np.array([1, 0, 2, -1, 4], np.int64)
    out["kat_out"] = ref.crosscheck_cpu(t(out["kat_in0"]), t(out["kat_in1"])).numpy()
    np.savez_compressed(os.path.join(HERE, "crosscheck.npz"), **out)

    # ---- LCN (reference torch module on CPU) ----
    rng = np.random.RandomState(11)
    LCN = reference_lcn_class()
    out = {}
    x = rng.rand(2, 1, 24, 32).astype(np.float32)
    x[1, 0, :, :16] = 0.5                  # a flat half: var -> 0 (+1e-6)
    with torch.no_grad():
        for r, e in ((5, 0.05), (2, 0.1)):
            l, s = LCN(r, e)(t(x))
            out[f"r{r}_lcn"], out[f"r{r}_std"] = l.numpy(), s.numpy()
    out["x"] = x
    np.savez_compressed(os.path.join(HERE, "lcn.npz"), **out)

    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
