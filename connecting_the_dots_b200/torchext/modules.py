"""nn.Module layer of the drop-in torchext package.

CoordConv2d is the one module the reference ships (torchext/modules.py:7-27).  The others are thin
nn.Module faces of the ops for callers that want modules (BASELINE.json's naming); `LCN` is a drop-in
for model.networks.LCN (model/networks.py:507-533): same constructor, forward returns (lcn, std)."""
import torch

from .functions import *  # noqa: F401,F403
from . import functions as F


class CoordConv2d(torch.nn.Module):
    """Conv2d over the input with two extra channels holding the pixel coordinates in [-1, 1]."""

    def __init__(self, channels_in, channels_out, kernel_size, stride, padding):
        super().__init__()
        self.conv = torch.nn.Conv2d(channels_in + 2, channels_out, kernel_size=kernel_size, padding=padding, stride=stride)
        self.uv = None

    def _grid(self, x):
        h, w = x.shape[2], x.shape[3]
        if self.uv is None or self.uv.shape[2:] != (h, w) or self.uv.device != x.device:
            u = torch.linspace(-1, 1, w, device=x.device, dtype=torch.float32).view(1, 1, 1, w).expand(1, 1, h, w)
            v = torch.linspace(-1, 1, h, device=x.device, dtype=torch.float32).view(1, 1, h, 1).expand(1, 1, h, w)
            self.uv = torch.cat((u, v), dim=1)
        return self.uv

    def forward(self, x):
        uv = self._grid(x).expand(x.shape[0], -1, -1, -1)
        return self.conv(torch.cat((x, uv), dim=1))


class PhotometricLoss(torch.nn.Module):
    def __init__(self, block_size=9, type="census_sad", eps=0.5):
        super().__init__()
        self.block_size, self.type, self.eps = block_size, type, eps

    def forward(self, es, ta):
        return F.photometric_loss(es, ta, self.block_size, self.type, self.eps)


class XCorrVol(torch.nn.Module):
    def __init__(self, n_disps=128, block_size=9):
        super().__init__()
        self.n_disps, self.block_size = n_disps, block_size

    def forward(self, in0, in1):
        return F.xcorrvol(in0, in1, self.n_disps, self.block_size)


class CrossCheck(torch.nn.Module):
    def forward(self, in0, in1):
        return F.crosscheck(in0, in1)


class NN(torch.nn.Module):
    def forward(self, in0, in1):
        return F.nn(in0, in1)


class ProjNN(torch.nn.Module):
    def __init__(self, patch_size=3):
        super().__init__()
        self.patch_size = patch_size

    def forward(self, xyz0, xyz1, K):
        return F.proj_nn(xyz0, xyz1, K, self.patch_size)


class LCN(torch.nn.Module):
    """Drop-in for model.networks.LCN(radius, epsilon): forward(data) -> (lcn, std)."""

    def __init__(self, radius=5, epsilon=0.05):
        super().__init__()
        self.radius, self.epsilon = radius, epsilon

    def forward(self, data):
        return F.lcn(data, self.radius, self.epsilon)
