"""TEST / BENCH INFRASTRUCTURE ONLY: times the reference's OWN CUDA extension on this GPU.

oracle/_ref/ctd_ref_ext_cuda.so is /root/reference/torchext/ext/{ext_cuda.cpp,ext_kernel.cu} compiled unmodified
for sm_100 by oracle/build_ref.py (the generic grid-stride `iterate_kernel`, common_cuda.h:159-170, 1024 threads
per block, launched on the legacy default stream).  It is the "existing GPU kernel" speed bar of SURVEY.md
section 2.2 / 6.  LCN has no reference kernel: the reference runs `model/networks.py:523-533` as torch ops on the
GPU (two cuDNN convolutions + elementwise kernels), which is what `lcn_torch` restates.

Nothing in connecting_the_dots_b200/ imports this module; tools/bench_ops.py and bench.py's `gpu_reference`
block do, for the comparison leg only.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def load():
    """The reference CUDA module, or None when oracle/_ref holds no build of it."""
    try:
        import build_ref
        return build_ref.load_ref(cuda=True)
    except Exception:
        return None


def lcn_torch(x, radius=5, eps=0.05):
    """model/networks.py:523-533 with the same torch ops (reflection pad, all-ones conv2d)."""
    import torch
    k = 2 * radius + 1
    w = torch.ones(1, 1, k, k, device=x.device, dtype=x.dtype)
    pad = torch.nn.functional.pad(x, (radius,) * 4, mode="reflect")
    box = torch.nn.functional.conv2d(pad, w)
    box2 = torch.nn.functional.conv2d(pad * pad, w)
    avg = box / k ** 2
    std = torch.sqrt(box2 / k ** 2 - avg ** 2 + 1e-6) + eps
    return (x - avg) / std, std


def _time(fn, iters, warmup=2):
    """Per-call milliseconds (median, min): CUDA events on torch's default stream, which IS the legacy default
    stream the reference launches on.  Stream launches, as the reference issues them (it cannot be graph-captured:
    legacy stream)."""
    import numpy as np
    import torch
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    evs[0].record()
    for i in range(iters):
        fn(i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]
    return float(np.median(ts)), float(np.min(ts))


def time_ops(sets, B, H, W, iters=5, which=("lcn", "photometric", "xcorrvol", "proj_nn", "crosscheck", "nn"), clouds=None,
             xcorr_images=None):
    """Per-op timings of the reference CUDA path on rotating buffer sets `sets` (dicts of CUDA tensors with keys im, es, ta,
    go, pat_lcn as tools/bench_ops.py builds them).  Returns {op: {"ms": median, "ms_min": min, ...}} or None."""
    import torch
    ref = load()
    if ref is None:
        return None
    torch.backends.cudnn.allow_tf32 = False  # the reference's convolutions evaluated in fp32
    ns = len(sets)
    out = {}

    def put(name, t, **extra):
        out[name] = dict(ms=t[0], ms_min=t[1], **extra)

    if "lcn" in which:
        put("lcn_fwd", _time(lambda i: lcn_torch(sets[i % ns]["im"]), iters), note="torch ops on the GPU (2 cuDNN convs + elementwise), networks.py:523-533")
    if "photometric" in which:
        for ty, name in ((1, "sad"), (3, "census_sad")):
            put(name + "_fwd", _time(lambda i, ty=ty: ref.photometric_loss_forward(sets[i % ns]["es"], sets[i % ns]["ta"], 9, ty, 0.5), iters))
            put(name + "_bwd", _time(lambda i, ty=ty: ref.photometric_loss_backward(sets[i % ns]["es"], sets[i % ns]["ta"], sets[i % ns]["go"], 9, ty, 0.5), iters))
    if "xcorrvol" in which:
        nimg = B if xcorr_images is None else xcorr_images
        for bs in (9, 5):
            def f(i, bs=bs):
                d = sets[i % ns]
                for n in range(nimg):  # the reference entry point has no batch dimension (ext_cuda.cpp:73-86)
                    ref.xcorrvol_cuda(d["ta"][n], d["pat_lcn"][n], 128, bs)
            t = _time(f, max(2, iters // 2), warmup=1)
            put("xcorrvol_D128_bs%d" % bs, (t[0] * B / nimg, t[1] * B / nimg), images_timed=nimg, images_reported=B)
    if clouds is not None:
        x0, x1, Kd = clouds
        npairs = x0.shape[0]
        idx = None
        if "proj_nn" in which:
            for ps in (3, 5):
                put("proj_nn_ps%d_%dpairs" % (ps, npairs), _time(lambda i, ps=ps: ref.proj_nn_cuda(x0, x1, Kd, ps), iters))
            idx = ref.proj_nn_cuda(x0, x1, Kd, 3).view(-1)
        if "crosscheck" in which and idx is not None:
            rev = idx.view(npairs, -1).flip(0).contiguous().view(-1)
            put("crosscheck_%d" % idx.numel(), _time(lambda i: ref.crosscheck_cuda(idx, rev), iters))
    if "nn" in which:
        n = 16384
        g = torch.Generator(device="cpu").manual_seed(0)
        p0 = torch.randn(n, 3, generator=g).cuda()
        p1 = torch.randn(n, 3, generator=g).cuda()
        t = _time(lambda i: ref.nn_cuda(p0, p1), iters)
        put("nn_16384x16384", t, gpair_s=n * n / t[0] / 1e6)
    return out
