"""The reference's op API (torchext/functions.py:5-147), same names, signatures, defaults and return
types, dispatching to the B200 kernels.  CUDA tensors go to ext_cuda (libctd_b200.so); CPU tensors
reach ext_cpu, which raises -- there is no CPU implementation.

Only photometric_loss is differentiable (w.r.t. `es`); the index / cost-volume ops return None
gradients like the reference (functions.py:15-17,33-35,50-52,69-71).
"""
import torch

from . import ext_cpu
from . import ext_cuda

LOSS_TYPES = {"mse": 0, "sad": 1, "census_mse": 2, "census_sad": 3}  # ext.h:196-199


def _native(t, cuda_name, cpu_name):
    return getattr(ext_cuda, cuda_name) if t.is_cuda else getattr(ext_cpu, cpu_name)


class NNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in0, in1):
        return _native(in0, "nn_cuda", "nn_cpu")(in0, in1)

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


def nn(in0, in1):
    """For every row of in0 [N0,3] the index of its nearest row of in1 [N1,3] (int64, -1 if none)."""
    return NNFunction.apply(in0, in1)


class CrossCheckFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in0, in1):
        return _native(in0, "crosscheck_cuda", "crosscheck_cpu")(in0, in1)

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


def crosscheck(in0, in1):
    """uint8 mask: in0[i] = j >= 0 and in1[j] == i (mutual consistency of two index maps)."""
    return CrossCheckFunction.apply(in0, in1)


class ProjNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz0, xyz1, K, patch_size):
        return _native(xyz0, "proj_nn_cuda", "proj_nn_cpu")(xyz0, xyz1, K, patch_size)

    @staticmethod
    def backward(ctx, grad_out):
        return None, None, None, None


def proj_nn(xyz0, xyz1, K, patch_size):
    """Project xyz0 with K, search a patch_size^2 pixel patch of xyz1 for the nearest 3-D point."""
    return ProjNNFunction.apply(xyz0, xyz1, K, patch_size)


class XCorrVolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in0, in1, n_disps, block_size):
        return _native(in0, "xcorrvol_cuda", "xcorrvol_cpu")(in0, in1, n_disps, block_size)

    @staticmethod
    def backward(ctx, grad_out):
        return None, None, None, None


def xcorrvol(in0, in1, n_disps, block_size):
    """Normalised cross-correlation cost volume [n_disps,H,W] of in0 against in1 shifted by d."""
    return XCorrVolFunction.apply(in0, in1, n_disps, block_size)


class PhotometricLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, es, ta, block_size, type, eps):
        ctx.save_for_backward(es, ta)
        ctx.block_size, ctx.type, ctx.eps = block_size, type, eps
        return _native(es, "photometric_loss_forward", "photometric_loss_forward")(es, ta, block_size, type, eps)

    @staticmethod
    def backward(ctx, grad_out):
        es, ta = ctx.saved_tensors
        fn = _native(grad_out, "photometric_loss_backward", "photometric_loss_backward")
        grad_es = fn(es, ta, grad_out.contiguous(), ctx.block_size, ctx.type, ctx.eps)
        return grad_es, None, None, None, None


def _loss_type_id(type):
    try:
        return LOSS_TYPES[type.lower()]
    except KeyError:
        raise Exception("invalid loss type")


def photometric_loss(es, ta, block_size, type="mse", eps=0.1):
    """Block-window photometric loss map [B,1,H,W]; type in mse | sad | census_mse | census_sad."""
    return PhotometricLossFunction.apply(es, ta, block_size, _loss_type_id(type), eps)


def photometric_loss_pytorch(es, ta, block_size, type="mse", eps=0.1):
    """Plain-torch evaluation of the same loss (differentiable by autograd, any device); the
    counterpart of the reference's functions.py:120-147, written as a loop over window offsets on a
    replicate-padded copy instead of an unfold, so it needs O(1) extra image copies."""
    kind = type.lower()
    if kind not in LOSS_TYPES:
        raise Exception("invalid loss type")
    lo = block_size // 2
    hi = block_size - 1 - lo
    H, W = es.shape[2], es.shape[3]
    es_p = torch.nn.functional.pad(es, (lo, hi, lo, hi), mode="replicate")
    ta_p = torch.nn.functional.pad(ta, (lo, hi, lo, hi), mode="replicate")

    def soft_step(x):
        return 0.5 * (1 + x / torch.sqrt(x * x + eps))

    total = None
    for dy in range(block_size):
        for dx in range(block_size):
            e = es_p[:, :, dy:dy + H, dx:dx + W]
            t = ta_p[:, :, dy:dy + H, dx:dx + W]
            d = (soft_step(e - es) - soft_step(t - ta)) if kind.startswith("census") else (e - t)
            term = (d * d if kind.endswith("mse") else d.abs()).sum(dim=1, keepdim=True)
            total = term if total is None else total + term
    return total / block_size ** 2


class WeightedPhotometricLossFunction(torch.autograd.Function):
    """`(mask * photometric_loss(es, ta)).sum() / mask.sum()` -- the way the reference's only caller consumes the
    loss map (model/networks.py:377) -- as ONE op.  Its upstream gradient is a scalar, so the backward's grad_out is
    the mask up to that scalar and is known before the loss map exists: loss map, gradient and both sums come out of
    one fused pass (ctd_photometric_fwd_bwd_masked_f32, a single kernel for the census modes) and autograd's backward
    is a scalar multiply.  SURVEY section 8(f) rank 1, without the warp.  Returns (value, loss_map); mask is not
    differentiated (it is LCN's std in the reference)."""

    @staticmethod
    def forward(ctx, es, ta, mask, block_size, type, eps):
        if not es.is_cuda:
            raise RuntimeError("torchext.weighted_photometric_loss: connecting_the_dots_b200 has no CPU implementation")
        mask = mask.detach().contiguous()
        # grad_out = mask: the gradient comes back unnormalised and is divided by sum(mask) in backward
        out, grad_es, sums = ext_cuda.photometric_loss_forward_backward_masked(es.detach(), ta.detach(), mask, mask, block_size, type, eps)
        ctx.save_for_backward(grad_es, sums)
        ctx.mark_non_differentiable(out)
        return sums[0] / sums[1], out

    @staticmethod
    def backward(ctx, g_val, g_map):
        grad_es, sums = ctx.saved_tensors
        return grad_es * (g_val / sums[1]), None, None, None, None, None


def weighted_photometric_loss(es, ta, mask, block_size, type="mse", eps=0.1):
    """Masked mean of the photometric loss map and the map itself: (value, loss_map [B,1,H,W])."""
    return WeightedPhotometricLossFunction.apply(es, ta, mask, block_size, _loss_type_id(type), eps)


class WarpPatternFunction(torch.autograd.Function):
    """The disparity warp of the reference pattern (model/networks.py:362-371): one gather kernel instead of
    building uv grids and calling grid_sample; differentiable w.r.t. disp."""

    @staticmethod
    def forward(ctx, pattern, disp):
        if not disp.is_cuda:
            raise RuntimeError("torchext.warp_pattern: connecting_the_dots_b200 has no CPU implementation")
        pattern, disp = pattern.detach().contiguous(), disp.detach().contiguous()
        ctx.save_for_backward(pattern, disp)
        return ext_cuda.warp_pattern_forward(pattern, disp)

    @staticmethod
    def backward(ctx, grad_out):
        pattern, disp = ctx.saved_tensors
        return None, ext_cuda.warp_pattern_backward(pattern, disp, grad_out.contiguous())


def warp_pattern(pattern, disp):
    """pattern [Bp,1,Hp,Wp] sampled at (u - disp, v) with the reference's grid convention -> [B,1,H,W]."""
    return WarpPatternFunction.apply(pattern, disp)


class PatternSimilarityFunction(torch.autograd.Function):
    """RectifiedPatternSimilarityLoss.tforward (model/networks.py:358-378) with its backward to the disparity as ONE
    kernel (census modes, block 9): the disparity warp of the pattern happens inside the loss kernel's tile loader, so
    pattern_proj and d loss / d es never make a round trip through HBM as separate kernels' outputs and inputs.  The upstream
    gradient of the value is a scalar; autograd's backward is a scalar multiply."""

    @staticmethod
    def forward(ctx, disp, pattern, im, mask, type, eps):
        d = disp.detach().contiguous()
        m = mask.detach().contiguous()
        proj, out, gd, sums = ext_cuda.pattern_similarity(pattern.detach().contiguous(), d, im.detach().contiguous(), m, m, type, eps)
        ctx.save_for_backward(gd, sums)
        ctx.mark_non_differentiable(proj, out)
        return sums[0] / sums[1], proj, out

    @staticmethod
    def backward(ctx, g_val, g_proj, g_map):
        gd, sums = ctx.saved_tensors
        return gd * (g_val / sums[1]), None, None, None, None, None


def pattern_similarity_loss(disp, pattern, im, std=None, loss_type="census_sad", loss_eps=0.5, block_size=9, fused=False):
    """RectifiedPatternSimilarityLoss.tforward (model/networks.py:358-378).  Returns (val, pattern_proj).
    Default: three kernels (warp; fused loss forward + backward + masked mean; in autograd's backward the warp's gradient).
    fused=True (census modes, block 9, CUDA float32 -- the reference's configuration, networks.py:344,376): ONE kernel forward
    + backward (PatternSimilarityFunction).  Measured on B200 at batch 8 x 480x640 (profiles/r02_ops_b8.json): 175 us for
    the one kernel against 164 us for the three -- the tile loader warps 1.7 pixels per output pixel (halo) behind two
    dependent global loads, which costs more than the 27 us of the two stand-alone warp kernels -- so it is not the default."""
    mask = torch.ones_like(im) if std is None else std
    ty = _loss_type_id(loss_type)
    if (fused and ty >= 2 and block_size == 9 and disp.is_cuda and disp.dtype == torch.float32 and disp.dim() == 4 and disp.size(1) == 1
            and disp.size(2) >= 9 and disp.size(3) >= 9):
        val, pattern_proj, _ = PatternSimilarityFunction.apply(disp, pattern, im, mask, ty, loss_eps)
        return val, pattern_proj
    pattern_proj = warp_pattern(pattern, disp)
    val, _ = weighted_photometric_loss(pattern_proj, im.contiguous(), mask, block_size, loss_type, loss_eps)
    return val, pattern_proj


class ProjectionDepthSimilarityFunction(torch.autograd.Function):
    """The geometric loss of the stage-2 trainer, ProjectionDepthSimilarityLoss.tforward (model/networks.py:474-503,
    called per frame pair at model/exp_synphge.py:185-200): unproject, two rigid transforms, projection, bilinear
    sampling of the other frame's depth and the (clamped) absolute difference, both directions.  In the reference
    that is 4 bmm, 2 grid_sample and ~20 elementwise kernels plus their autograd twins; here two launches produce
    the loss and both depth gradients (the upstream gradient of a scalar loss is a scalar, applied in backward).
    Poses, K and the ray table are not differentiated (they are data in the reference)."""

    @staticmethod
    def forward(ctx, depth0, depth1, R0, t0, R1, t1, K, ray, clamp):
        if not depth0.is_cuda:
            raise RuntimeError("torchext.projection_depth_similarity_loss: connecting_the_dots_b200 has no CPU implementation")
        c = lambda t: t.detach().contiguous()
        sums, g0, g1 = ext_cuda.depth_similarity(c(depth0), c(depth1), c(ray), c(K), c(R0), c(t0), c(R1), c(t1), clamp)
        ctx.save_for_backward(g0, g1)
        return sums[0, 0] / sums[0, 1] + sums[1, 0] / sums[1, 1]

    @staticmethod
    def backward(ctx, g):
        g0, g1 = ctx.saved_tensors
        return g0 * g, g1 * g, None, None, None, None, None, None, None


def projection_depth_similarity_loss(depth0, depth1, R0, t0, R1, t1, K, ray, clamp=-1):
    """l0 + l1 of ProjectionDepthSimilarityLoss.tforward(depth0, depth1, R0, t0, R1, t1) (model/networks.py:500-503).
    K [3,3] and ray [H*W,3] are the module's constants (networks.py:421-434: ray = uv1 @ Ki^T in float32)."""
    return ProjectionDepthSimilarityFunction.apply(depth0, depth1, R0, t0, R1, t1, K, ray, clamp)


def projection_rays(Ki, im_height, im_width):
    """The ray table of ProjectionBaseLoss.__init__ (model/networks.py:427-434): pixel centres (u, v, 1) times
    Ki^T, computed in float64 by numpy and stored as float32 [H*W,3]."""
    import numpy as np
    u, v = np.meshgrid(range(im_width), range(im_height))
    uv = np.stack((u, v, np.ones_like(u)), axis=2).reshape(-1, 3)
    ray = uv @ np.asarray(Ki, dtype=np.float64).T if not torch.is_tensor(Ki) else uv @ Ki.cpu().numpy().T
    return torch.from_numpy(ray.reshape(-1, 3).astype(np.float32))


class DisparityLossFunction(torch.autograd.Function):
    """DisparityLoss.tforward (model/networks.py:395-411, called at model/exp_synph.py:116): 5x5 Sobel gradient
    magnitude of the disparity under a two-Laplacian mixture weighted by the edge probability (or, without an edge
    map, its clamped mean).  One kernel computes the loss and the gradients w.r.t. disp and edge; autograd's backward
    is a scalar multiply."""

    @staticmethod
    def forward(ctx, disp, edge):
        if not disp.is_cuda:
            raise RuntimeError("torchext.disparity_loss: connecting_the_dots_b200 has no CPU implementation")
        sums, gd, ge = ext_cuda.disparity_loss(disp.detach().contiguous(), None if edge is None else edge.detach().contiguous())
        ctx.has_edge = edge is not None
        ctx.save_for_backward(gd, ge if ge is not None else gd)
        return sums[0] / sums[1]

    @staticmethod
    def backward(ctx, g):
        gd, ge = ctx.saved_tensors
        return gd * g, (ge * g if ctx.has_edge else None)


def disparity_loss(disp, edge=None):
    """DisparityLoss()(disp, edge) of the reference (model/networks.py:380-412)."""
    return DisparityLossFunction.apply(disp, edge)


class LCNFunction(torch.autograd.Function):
    """Fused local contrast normalisation (model/networks.py:507-533) with its gradient: the forward is one kernel, the
    backward (ctd_lcn_bwd_f32) is what autograd produces through the reference's torch ops for upstream gradients of
    both outputs.  (In the reference the gradient flows to an input image nobody reads, exp_synph.py:80-91.)"""

    @staticmethod
    def forward(ctx, x, radius, epsilon):
        if not x.is_cuda:
            raise RuntimeError("torchext.lcn: connecting_the_dots_b200 has no CPU implementation")
        out, std = ext_cuda.lcn_forward(x, radius, epsilon)
        ctx.radius, ctx.epsilon = radius, epsilon
        if x.dtype == torch.float32 and x.requires_grad:
            ctx.save_for_backward(x.detach(), out, std)
        else:
            ctx.mark_non_differentiable(out, std)
        return out, std

    @staticmethod
    def backward(ctx, g_lcn, g_std):
        x, out, std = ctx.saved_tensors
        gl = None if g_lcn is None else g_lcn.contiguous()
        gs = None if g_std is None else g_std.contiguous()
        return ext_cuda.lcn_backward(x, out, std, gl, gs, ctx.radius, ctx.epsilon), None, None


def lcn(x, radius=5, epsilon=0.05):
    """(x - box_mean) / (box_std + epsilon) over a (2*radius+1)^2 reflection-padded window -> (lcn, std)."""
    return LCNFunction.apply(x, radius, epsilon)


def lcn_normalize(img, kernel_size=4, epsilon=0.01):
    """The data generator's offline LCN, data/lcn/lcn.pyx:16-58 `normalize(img, kernel_size, epsilon)` (called at
    data/create_syn_data.py:182): img [M,N] (or a batch [B,M,N]) -> (img_lcn, img_std), a border of kernel_size pixels left
    at 0, std = sqrt of the centred two-pass variance.  Bit-identical to the Cython build; CUDA tensors only."""
    if not img.is_cuda:
        raise RuntimeError("torchext.lcn_normalize: connecting_the_dots_b200 has no CPU implementation")
    return ext_cuda.lcn_cython(img.contiguous(), kernel_size, epsilon)


def pyramid_pattern_similarity_loss(disps, patterns, ims, stds, loss_type="census_sad", loss_eps=0.5, block_size=9):
    """The photometric part of the trainer's loss_forward (model/exp_synph.py:107-111) over the image pyramid
    (exp_synph.py:25-27: 480x640 halved three times): for every scale s, RectifiedPatternSimilarityLoss(disps[s],
    ims[s], stds[s]) with that scale's pattern (exp_synph.py:62-71).  `patterns` is one [1,1,Hp,Wp] tensor per scale (or
    one tensor used for all).  Returns (vals: list of scalar tensors, pattern_proj of scale 0, detached) as the reference
    keeps them; 3 kernels per scale forward+backward (warp, fused loss + gradient + masked mean, warp gradient), all
    capture-safe, so the whole pyramid replays as one CUDA graph."""
    if torch.is_tensor(patterns):
        patterns = [patterns] * len(disps)
    if not (len(disps) == len(patterns) == len(ims) == len(stds)):
        raise RuntimeError("pyramid_pattern_similarity_loss: one disparity, pattern, image and std per scale")
    vals, proj0 = [], None
    for s, (d, p, im, sd) in enumerate(zip(disps, patterns, ims, stds)):
        val, proj = pattern_similarity_loss(d, p, im, sd, loss_type, loss_eps, block_size)
        if s == 0:
            proj0 = proj.detach()
        vals.append(val)
    return vals, proj0


def masked_mean_terms(diff, mask):
    """(sum(mask*diff), sum(mask)) as a float32 [2] tensor: the two scalars of the caller's
    `(mask*diff).sum() / mask.sum()` (model/networks.py:377); under batch sharding they are what gets
    all-reduced.  Not differentiable (use photometric_loss's own backward with grad_out = mask / sum)."""
    return ext_cuda.masked_sums(diff.detach(), mask.detach())
