"""Seeded synthetic inputs of the op path: 480x640 Kinect-like dot-pattern image
pairs, LCN'd, with a smooth ground-truth disparity, plus point clouds for ProjNN/NN.  numpy only;
this produces INPUTS for tests and bench.py and is never part of a measured path.

Recipe anchors in the reference: the 10%-density random dot pattern of data/commons.py:9-12, the
pattern/ambient blend and sensor noise of data/commons.py:92-107 and create_syn_data.py:171,231, the
horizontal warp `pattern(x - disp)` of model/networks.py:362-371, LCN(5, 0.05) of exp_synph.py:41,
K and the disparity->depth relation of create_syn_data.py:228-230.
"""
import numpy as np

H0, W0 = 480, 640
K_REF = np.array([[567.6, 0, 324.7], [0, 570.2, 250.1], [0, 0, 1]], np.float32)
BASELINE, FOCAL = 0.075, 567.6


def dot_pattern(H=H0, W=W0):
    return (np.random.RandomState(42).uniform(0, 1, (H, W)) < 0.1).astype(np.float32)


def smooth_disparity(rng, H, W, lo=8.0, hi=120.0):
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    d = np.zeros((H, W), np.float32)
    for _ in range(3):
        fx, fy = rng.uniform(0.5, 2.5, 2) * 2 * np.pi / np.array([W, H])
        d += rng.uniform(0.5, 1.0) * np.sin(fx * x + fy * y + rng.uniform(0, 2 * np.pi)).astype(np.float32)
    d = (d - d.min()) / max(float(d.max() - d.min()), 1e-6)
    return (lo + (hi - lo) * d).astype(np.float32)


def warp_rows(img, disp):
    """img sampled at (y, x - disp), linear interpolation, border clamp."""
    H, W = img.shape
    xs = np.clip(np.arange(W, dtype=np.float32)[None, :] - disp, 0, W - 1)
    x0 = np.floor(xs).astype(np.int64)
    x1 = np.minimum(x0 + 1, W - 1)
    f = (xs - x0).astype(np.float32)
    rows = np.arange(H)[:, None]
    return ((1 - f) * img[rows, x0] + f * img[rows, x1]).astype(np.float32)


def lcn_np(x, radius=5, epsilon=0.05):
    """Input-generation LCN (float64 integral image); NOT the oracle and not the measured kernel."""
    r = radius
    p = np.pad(x.astype(np.float64), r, mode="reflect")

    def box(a):
        c = np.cumsum(np.cumsum(np.pad(a, ((1, 0), (1, 0))), 0), 1)
        k = 2 * r + 1
        return c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]

    n = (2 * r + 1) ** 2
    avg = box(p) / n
    std = np.sqrt(np.maximum(box(p * p) / n - avg * avg, 0) + 1e-6) + epsilon
    return ((x - avg) / std).astype(np.float32), std.astype(np.float32)


def make_pair(n, H=H0, W=W0, pattern=None):
    """One synthetic frame: dict with im, disp, es (warped LCN pattern), ta (LCN image), std, go."""
    P = dot_pattern(H, W) if pattern is None else pattern
    rng = np.random.RandomState(1000 + n)
    disp = smooth_disparity(rng, H, W, 8.0 * W / W0, 120.0 * W / W0)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    ambient = (0.5 + 0.3 * np.sin(2 * np.pi * (x / W * rng.uniform(0.3, 1.0) + y / H * rng.uniform(0.3, 1.0)))).astype(np.float32)
    sigma = rng.uniform(0, 3) / 255.0
    im = np.clip(0.6 * warp_rows(P, disp) + 0.4 * ambient + rng.normal(0, sigma, (H, W)), 0, 1).astype(np.float32)
    im_lcn, std = lcn_np(im)
    pat_lcn, _ = lcn_np(P)
    es = warp_rows(pat_lcn, disp + rng.normal(0, 0.5, (H, W)).astype(np.float32))
    return {"im": im, "disp": disp, "es": es, "ta": im_lcn, "std": std, "go": (std / std.sum()).astype(np.float32),
            "pat_lcn": pat_lcn}


def make_batch(B, H=H0, W=W0, distinct=8):
    """[B,1,H,W] arrays: im, es, ta, std, go, pat_lcn, disp.  Only `distinct` frames are generated; the
    rest are row-rolled copies (values identical in distribution, cheap to make for B = 64)."""
    P = dot_pattern(H, W)
    base = [make_pair(n, H, W, P) for n in range(min(B, distinct))]
    keys = ("im", "es", "ta", "std", "go", "pat_lcn", "disp")
    out = {k: np.empty((B, 1, H, W), np.float32) for k in keys}
    for b in range(B):
        src, shift = base[b % len(base)], (b // len(base)) * 7
        for k in keys:
            out[k][b, 0] = np.roll(src[k], shift, axis=0)
    return out


def make_clouds(n_frames, H=H0, W=W0, seed=0):
    """Point clouds of a frame track: returns (xyz [T,H,W,3] in each frame's own camera, K, poses), the
    scene being the disparity surfaces of make_pair turned into depth = baseline * focal / disp."""
    rng = np.random.RandomState(seed)
    K = K_REF.copy()
    K[0] *= W / W0
    K[1] *= H / H0
    v, u = np.mgrid[0:H, 0:W].astype(np.float32)
    ray = np.stack(((u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)), -1)
    xyz, poses = [], []
    for t in range(n_frames):
        disp = smooth_disparity(np.random.RandomState(2000 + seed), H, W) + 0.3 * t
        depth = (BASELINE * FOCAL / disp).astype(np.float32)
        xyz.append((ray * depth[..., None]).astype(np.float32))
        a = rng.uniform(-0.01, 0.01, 3)
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        poses.append(((Rx @ Ry @ Rz).astype(np.float32), rng.uniform(-0.01, 0.01, 3).astype(np.float32)))
    return np.stack(xyz), K, poses


def transform(xyz, pose):
    R, t = pose
    return (xyz @ R.T + t).astype(np.float32)


def make_depth_pairs(B, H=H0, W=W0, seed=0):
    """Inputs of the geometric loss (model/exp_synphge.py:185-200) for B frame pairs: depth0, depth1 [B,1,H,W]
    (depth = baseline * focal / disparity of two nearby smooth surfaces, as DispToDepth produces them), camera
    poses R0, t0, R1, t1 of a small rigid motion between the frames, and the intrinsics K, Ki scaled to (H, W)."""
    rng = np.random.RandomState(7000 + seed)
    K = K_REF.astype(np.float64).copy()
    K[0] *= W / W0
    K[1] *= H / H0
    Ki = np.linalg.inv(K)

    def rot(a):
        Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
        Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
        Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
        return Rx @ Ry @ Rz

    d0, d1, R0, t0, R1, t1 = [], [], [], [], [], []
    for n in range(B):
        disp = smooth_disparity(np.random.RandomState(3000 + seed + n % 8), H, W)
        d0.append((BASELINE * FOCAL * (W / W0) / disp).astype(np.float32))
        d1.append((BASELINE * FOCAL * (W / W0) / (disp + 0.3 + 0.1 * np.sin(np.arange(W) / 37.0))).astype(np.float32))
        R0.append(rot(rng.uniform(-0.01, 0.01, 3)))
        R1.append(rot(rng.uniform(-0.01, 0.01, 3)))
        t0.append(rng.uniform(-0.01, 0.01, 3))
        t1.append(rng.uniform(-0.01, 0.01, 3))
    f = lambda a: np.ascontiguousarray(np.stack(a), dtype=np.float32)
    return {"depth0": f(d0)[:, None], "depth1": f(d1)[:, None], "R0": f(R0), "t0": f(t0), "R1": f(R1), "t1": f(t1),
            "K": K.astype(np.float32), "Ki": Ki.astype(np.float32)}

