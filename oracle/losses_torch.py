"""TEST INFRASTRUCTURE ONLY: torch (fp32, any device) restatements of the reference's per-pixel losses either side of
the torchext path.  They call the same torch operations in the same order as the
reference classes, without their TimedModule plumbing; tests/test_oracle.py pins them against golden vectors produced
by the reference classes themselves (tests/golden/make_golden_geometric.py).  Only tests/ and tools/bench_ops.py's
comparison legs may import this module.
"""
import torch

SOBEL_KX = [[-5, -4, 0, 4, 5], [-8, -10, 0, 10, 8], [-10, -20, 0, 20, 10], [-8, -10, 0, 10, 8], [-5, -4, 0, 4, 5]]  # networks.py:541-545
B0, B1 = 0.0503428816795, 1.07274045944  # networks.py:390-391


def depth_similarity(depth0, depth1, R0, t0, R1, t1, K, ray, clamp):
    """ProjectionDepthSimilarityLoss.tforward, model/networks.py:500-503: fwd(0 -> 1) + fwd(1 -> 0)."""
    B, _, H, W = depth0.shape

    def fwd(dA, dB, RA, tA, RB, tB):  # networks.py:483-498 over unproject / transform / project, :436-472
        xyz = dA.reshape(B, -1, 1) * ray.reshape(1, -1, 3)
        xyz = torch.bmm(xyz - tA.reshape(B, 1, 3), RA)
        xyz = torch.bmm(xyz, RB.transpose(1, 2)) + tB.reshape(B, 1, 3)
        uv = torch.bmm(xyz, K.reshape(1, 3, 3).transpose(1, 2).expand(B, -1, -1))
        d = uv[:, :, 2:3]
        uv = uv[:, :, :2] / (torch.nn.functional.relu(d) + 1e-12)
        g = torch.stack((2 * (uv[..., 0] / (W - 1) - 0.5), 2 * (uv[..., 1] / (H - 1) - 0.5)), -1).view(-1, H, W, 2)
        s = torch.nn.functional.grid_sample(dB, g, padding_mode="border", align_corners=False)
        diff = torch.abs(d.view(-1) - s.view(-1))
        if clamp > 0:
            diff = torch.clamp(diff, 0, clamp)
        return diff.mean()

    return fwd(depth0, depth1, R0, t0, R1, t1) + fwd(depth1, depth0, R1, t1, R0, t0)


def disparity_loss(disp, edge=None):
    """DisparityLoss.tforward, model/networks.py:395-411, over SobelFilter.tforward, networks.py:558-565."""
    kx = torch.tensor(SOBEL_KX, dtype=torch.float64, device=disp.device) / 240.0
    x = torch.nn.functional.pad(disp, (2, 2, 2, 2), "replicate")
    gx = torch.nn.functional.conv2d(x, kx.float()[None, None])
    gy = torch.nn.functional.conv2d(x, kx.t().contiguous().float()[None, None])
    grad = torch.sqrt(gx ** 2 + gy ** 2 + 1e-8)
    if edge is None:
        return torch.mean(torch.clamp(grad, 0, 1.0))
    pdf = (1 - edge) / B0 * torch.exp(-torch.abs(grad) / B0) + edge / B1 * torch.exp(-torch.abs(grad) / B1)
    return torch.mean(-torch.log(pdf.clamp(min=1e-4)))
