#!/usr/bin/env python
"""Golden vectors for the geometric loss and the disparity loss from the UNMODIFIED
reference classes:
/root/reference/model/networks.py:414-503 (ProjectionBaseLoss, ProjectionDepthSimilarityLoss) are exec'd from the
reference file at run time on CPU torch with a stub for their TimedModule base (networks.py:10-23) -- importing
model.networks wholesale needs matplotlib (networks.py:4).  Stores the inputs, the loss l0 + l1 and torch
autograd's gradients w.r.t. both depth maps, with and without the clamp the trainer uses (exp_synphge.py:83); and
for DisparityLoss / SobelFilter (networks.py:380-412, 537-565) the loss and gradients w.r.t. disp and edge.

    python tests/golden/make_golden_geometric.py        (needs /root/reference; writes tests/golden/geometric.npz, disparity_loss.npz)
"""
import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from connecting_the_dots_b200 import synth  # noqa: E402

REF_NETWORKS = "/root/reference/model/networks.py"


def reference_classes():
    src = open(REF_NETWORKS).read()
    tree = ast.parse(src)
    want = ("ProjectionBaseLoss", "ProjectionDepthSimilarityLoss", "SobelFilter", "DisparityLoss")
    cls_src = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body if isinstance(n, ast.ClassDef) and n.name in want)

    class TimedModule(torch.nn.Module):  # stub of networks.py:10-23 without the device syncs/timer
        def __init__(self, mod_name):
            super().__init__()
            self.mod_name = mod_name

        def forward(self, *a, **k):
            return self.tforward(*a, **k)

    ns = {"torch": torch, "np": np, "F": torch.nn.functional, "TimedModule": TimedModule}
    exec(compile(cls_src, REF_NETWORKS, "exec"), ns)
    return ns["ProjectionDepthSimilarityLoss"], ns["DisparityLoss"]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    Loss, DispLoss = reference_classes()
    out = {}
    B, H, W = 2, 30, 40
    d = synth.make_depth_pairs(B, H, W, seed=1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    for key, v in d.items():
        out[key] = v
    for name, clamp in (("noclamp", -1), ("clamp", 0.1), ("tight", 0.004)):
        loss = Loss(t(d["K"]), t(d["Ki"]), H, W, clamp=clamp)
        out["ray"] = loss.ray.numpy().reshape(-1, 3).copy()
        d0 = t(d["depth0"]).clone().requires_grad_(True)
        d1 = t(d["depth1"]).clone().requires_grad_(True)
        val = loss(d0, d1, t(d["R0"]), t(d["t0"]), t(d["R1"]), t(d["t1"]))
        val.backward()
        out[name + "_val"] = val.detach().numpy()
        out[name + "_g0"] = d0.grad.numpy()
        out[name + "_g1"] = d1.grad.numpy()
        print(name, float(val), float(d0.grad.abs().max()), float(d1.grad.abs().max()))
    np.savez_compressed(os.path.join(HERE, "geometric.npz"), **out)

    # DisparityLoss (networks.py:380-412) over SobelFilter (networks.py:537-565): with an edge map and without
    rng = np.random.RandomState(5)
    B, H, W = 2, 37, 45
    # smooth surface with gradient magnitudes around b0 .. 1 px/px (both mixture components and the clamps are exercised)
    disp = np.stack([synth.smooth_disparity(np.random.RandomState(40 + n), H, W) / (12.0 + 20.0 * n) for n in range(B)])[:, None].astype(np.float32)
    disp[1, 0, 10:20, 15:30] += 6.0                       # a depth discontinuity
    disp[0, 0, 25:, :8] = disp[0, 0, 25, 0]               # a flat patch (gradient magnitude ~1e-4)
    edge = rng.uniform(0, 1, (B, 1, H, W)).astype(np.float32)
    edge[1, 0, 8:22, 13:32] = 0.97
    dl = {"disp": disp, "edge": edge}
    ref = DispLoss()
    for name, e in (("edge", edge), ("noedge", None)):
        d = t(disp).clone().requires_grad_(True)
        ee = None if e is None else t(e).clone().requires_grad_(True)
        val = ref(d, ee)
        val.backward()
        dl[name + "_val"] = val.detach().numpy()
        dl[name + "_gdisp"] = d.grad.numpy()
        if ee is not None:
            dl[name + "_gedge"] = ee.grad.numpy()
        print("disparity_loss", name, float(val.detach()), float(d.grad.abs().max()))
    np.savez_compressed(os.path.join(HERE, "disparity_loss.npz"), **dl)


if __name__ == "__main__":
    main()
