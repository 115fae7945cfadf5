"""CPU suite: pins the plain-C oracle (oracle/ctd_oracle.c) to the reference.

1. against the golden vectors in tests/golden/ (generated from the unmodified reference by
   tests/golden/make_golden.py) -- bit-exact for every torchext op;
2. against the compiled reference itself (oracle/_ref) on fresh random inputs, when shipped.
"""
import numpy as np
import pytest
import torch

import oracle
from conftest import assert_close

TYPES = ("mse", "sad", "census_mse", "census_sad")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


# ---------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("C", (1, 2))
@pytest.mark.parametrize("bs", (2, 3, 9))
@pytest.mark.parametrize("ty", range(4))
def test_photometric_golden_f32(golden, C, bs, ty):
    g = golden("photometric")
    es, ta, go = g[f"f32_C{C}_es"], g[f"f32_C{C}_ta"], g[f"f32_C{C}_go"]
    eps = 0.5 if ty == 3 else 0.1
    assert np.array_equal(oracle.photometric_loss_forward(es, ta, bs, ty, eps), g[f"f32_C{C}_bs{bs}_t{ty}_fwd"])
    assert np.array_equal(oracle.photometric_loss_backward(es, ta, go, bs, ty, eps), g[f"f32_C{C}_bs{bs}_t{ty}_bwd"])


@pytest.mark.parametrize("C", (1, 2))
@pytest.mark.parametrize("ty", range(4))
def test_photometric_golden_f64(golden, C, ty):
    g = golden("photometric")
    es, ta, go = g[f"f64_C{C}_es"], g[f"f64_C{C}_ta"], g[f"f64_C{C}_go"]
    eps = 0.5 if ty == 3 else 0.1
    assert np.array_equal(oracle.photometric_loss_forward(es, ta, 9, ty, eps), g[f"f64_C{C}_bs9_t{ty}_fwd"])
    assert np.array_equal(oracle.photometric_loss_backward(es, ta, go, 9, ty, eps), g[f"f64_C{C}_bs9_t{ty}_bwd"])


def test_xcorrvol_golden(golden):
    g = golden("xcorrvol")
    keys = [k[:-4] for k in g.files if k.endswith("_in0")]
    assert len(keys) == 4
    for k in keys:
        D = int(k.split("_D")[1].split("_")[0])
        bs = int(k.split("_bs")[1])
        assert np.array_equal(oracle.xcorrvol(g[k + "_in0"], g[k + "_in1"], D, bs), g[k + "_out"]), k
    flat = np.full((1, 6, 8), 0.25, np.float32)
    assert np.array_equal(oracle.xcorrvol(flat, flat, 3, 3), g["flat_out"])
    assert (g["flat_out"] == 0).all()                      # flat windows -> 0
    assert np.array_equal(oracle.xcorrvol(g["self_in"], g["self_in"], 1, 3), g["self_out"])
    assert np.allclose(g["self_out"], 2.0, atol=1e-5)      # xcorrvol(r, r, 1, bs) == C


@pytest.mark.parametrize("ps", (1, 2, 3, 5, 9))
def test_proj_nn_golden(golden, ps):
    g = golden("proj_nn")
    got = oracle.proj_nn(g["xyz0"], g["xyz1"], g["K"], ps)
    assert np.array_equal(got, g[f"ps{ps}"])
    if ps >= 3:
        assert (got[0, 0, :4] == -1).all()                 # inf, 0/0, NaN, out-of-int-range -> -1
        assert got[0, 0, 5] >= 0                           # z < 0 is not rejected
    if ps == 3:
        assert got[1, 4, 4] == (1 * 24 + 3) * 32 + 3       # ties -> first in (pv, pu) scan order


def test_nn_golden(golden):
    g = golden("nn")
    got = oracle.nn(g["p0"], g["p1"])
    assert np.array_equal(got, g["idx"])
    assert got[0] == 7 and got[1] == -1 and got[2] == -1
    assert np.array_equal(oracle.nn(g["p0"][:5], g["p1"][:0]), g["idx_empty"])
    assert (g["idx_empty"] == -1).all()


def test_crosscheck_golden(golden):
    g = golden("crosscheck")
    assert np.array_equal(oracle.crosscheck(g["in0"], g["in1"]), g["out"])
    assert 50 < g["out"].sum() < 400
    assert np.array_equal(oracle.crosscheck(g["kat_in0"], g["kat_in1"]), g["kat_out"])
    assert g["kat_out"].tolist() == [1, 1, 0, 0, 0]        # SURVEY.md A.5 known answer


@pytest.mark.parametrize("r,e", ((5, 0.05), (2, 0.1)))
def test_lcn_golden(golden, r, e):
    """The reference LCN is two fp32 library convolutions whose summation order is unspecified; its
    var = E[x^2] - E[x]^2 + 1e-6 cancels catastrophically where the image is flat, so two correct
    fp32 evaluations differ there by up to ~1e-4 relative in std.  The oracle uses exact (double)
    box sums; tolerance vs the reference's torch output: 2e-4 relative on std, 2e-4 * max|lcn| on
    lcn (measured worst case on this fixture: see DESIGN.md, LCN)."""
    g = golden("lcn")
    lcn, std = oracle.lcn(g["x"], r, e)
    assert np.abs(std - g[f"r{r}_std"]).max() <= 2e-4 * np.abs(g[f"r{r}_std"]).max()
    assert np.abs(lcn - g[f"r{r}_lcn"]).max() <= 2e-4 * np.abs(g[f"r{r}_lcn"]).max()
    tex = np.s_[0]                                         # textured image: much tighter
    assert np.abs(std[tex] - g[f"r{r}_std"][tex]).max() <= 2e-6
    assert np.abs(lcn[tex] - g[f"r{r}_lcn"][tex]).max() <= 2e-5


# ---------------------------------------------------------------- compiled reference, fresh inputs
@pytest.mark.parametrize("dt", (np.float32, np.float64))
def test_photometric_vs_ref(ref_ext, dt):
    rng = np.random.RandomState(0)
    for (B, C, H, W, bs) in [(2, 1, 13, 17, 9), (1, 2, 9, 11, 3), (1, 3, 7, 8, 2), (1, 1, 5, 6, 5), (1, 1, 3, 3, 9)]:
        es = rng.randn(B, C, H, W).astype(dt)
        ta = rng.randn(B, C, H, W).astype(dt)
        go = rng.rand(B, 1, H, W).astype(dt)
        for ty in range(4):
            a = oracle.photometric_loss_forward(es, ta, bs, ty, 0.5)
            b = ref_ext.photometric_loss_forward(t(es), t(ta), bs, ty, 0.5).numpy()
            assert np.array_equal(a, b)
            a = oracle.photometric_loss_backward(es, ta, go, bs, ty, 0.3)
            b = ref_ext.photometric_loss_backward(t(es), t(ta), t(go), bs, ty, 0.3).numpy()
            assert np.array_equal(a, b)


@pytest.mark.parametrize("dt", (np.float32, np.float64))
def test_index_ops_vs_ref(ref_ext, dt):
    rng = np.random.RandomState(1)
    for (C, H, W, D, bs) in [(1, 9, 20, 6, 5), (2, 8, 15, 17, 3), (1, 6, 9, 4, 2)]:
        a0, a1 = rng.rand(C, H, W).astype(dt), rng.rand(C, H, W).astype(dt)
        assert np.array_equal(oracle.xcorrvol(a0, a1, D, bs), ref_ext.xcorrvol_cpu(t(a0), t(a1), D, bs).numpy())
    K = np.array([[50., 0, 10], [0, 52, 8], [0, 0, 1]], dt)
    for ps in (1, 2, 3, 5):
        x0 = (rng.randn(2, 16, 20, 3) + [0, 0, 3]).astype(dt)
        x1 = (rng.randn(2, 16, 20, 3) + [0, 0, 3]).astype(dt)
        x0[0, 0, 0] = [1, 1, 0]
        x0[0, 0, 1] = [0, 0, 0]
        x0[0, 0, 2] = [np.nan, 1, 1]
        x0[0, 0, 3] = [1e30, 1, 1e-8]
        assert np.array_equal(oracle.proj_nn(x0, x1, K, ps), ref_ext.proj_nn_cpu(t(x0), t(x1), t(K), ps).numpy())
    p0, p1 = rng.randn(50, 3).astype(dt), rng.randn(70, 3).astype(dt)
    p1[5] = p1[3]
    assert np.array_equal(oracle.nn(p0, p1), ref_ext.nn_cpu(t(p0), t(p1)).numpy())
    i0 = rng.randint(-1, 40, 40).astype(np.int64)
    i1 = rng.randint(-1, 40, 40).astype(np.int64)
    i1[i0[i0 >= 0][:10]] = np.nonzero(i0 >= 0)[0][:10]
    assert np.array_equal(oracle.crosscheck(i0, i1), ref_ext.crosscheck_cpu(t(i0), t(i1)).numpy())


def test_photometric_matches_reference_torch_restatement():
    """functions.py:120-147 photometric_loss_pytorch is the reference's own second implementation;
    restated inline with torch ops (replicate pad + unfold) it must agree with the oracle <= 2e-6."""
    rng = np.random.RandomState(4)
    es = rng.randn(2, 2, 10, 12).astype(np.float32)
    ta = rng.randn(2, 2, 10, 12).astype(np.float32)
    bs, eps = 5, 0.5
    p = bs // 2
    E = torch.nn.functional.unfold(torch.nn.functional.pad(t(es), (p, p, p, p), mode="replicate"), bs).view(2, 2, -1, 10, 12)
    T = torch.nn.functional.unfold(torch.nn.functional.pad(t(ta), (p, p, p, p), mode="replicate"), bs).view(2, 2, -1, 10, 12)
    des, dta = E - t(es).unsqueeze(2), T - t(ta).unsqueeze(2)
    h = lambda x: 0.5 * (1 + x / torch.sqrt(x * x + eps))
    want = {"mse": (E - T) ** 2, "sad": (E - T).abs(), "census_mse": (h(des) - h(dta)) ** 2, "census_sad": (h(des) - h(dta)).abs()}
    for name, w in want.items():
        w = w.reshape(2, -1, 10, 12).sum(1, keepdim=True).numpy() / bs ** 2
        got = oracle.photometric_loss_forward(es, ta, bs, name, eps)
        assert np.abs(got - w).max() <= 2e-6 * max(1.0, np.abs(w).max()), name


# ---------------------------------------------------------------- next-row losses
@pytest.mark.parametrize("name,clamp", (("noclamp", -1), ("clamp", 0.1), ("tight", 0.004)))
def test_depth_similarity_restatement_golden(golden, name, clamp):
    """oracle/losses_torch.depth_similarity against the reference's own ProjectionDepthSimilarityLoss (golden)."""
    import torch
    import losses_torch
    g = golden("geometric")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    d0, d1 = t(g["depth0"]).requires_grad_(True), t(g["depth1"]).requires_grad_(True)
    val = losses_torch.depth_similarity(d0, d1, t(g["R0"]), t(g["t0"]), t(g["R1"]), t(g["t1"]), t(g["K"]), t(g["ray"]), clamp)
    val.backward()
    assert abs(float(val.detach()) - float(g[name + "_val"])) <= 1e-6 * abs(float(g[name + "_val"]))
    assert_close(d0.grad.numpy(), g[name + "_g0"], tol=1e-6, what=name + " grad depth0")
    assert_close(d1.grad.numpy(), g[name + "_g1"], tol=1e-6, what=name + " grad depth1")


@pytest.mark.parametrize("name", ("edge", "noedge"))
def test_disparity_loss_restatement_golden(golden, name):
    """oracle/losses_torch.disparity_loss against the reference's own DisparityLoss / SobelFilter (golden)."""
    import torch
    import losses_torch
    g = golden("disparity_loss")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    d = t(g["disp"]).requires_grad_(True)
    e = t(g["edge"]).requires_grad_(True) if name == "edge" else None
    val = losses_torch.disparity_loss(d, e)
    val.backward()
    assert abs(float(val.detach()) - float(g[name + "_val"])) <= 1e-6 * abs(float(g[name + "_val"]))
    assert_close(d.grad.numpy(), g[name + "_gdisp"], tol=1e-6, what=name + " grad disp")
    if e is not None:
        assert_close(e.grad.numpy(), g[name + "_gedge"], tol=1e-6, what=name + " grad edge")



def test_lcn_cython_oracle_matches_reference_build(golden):
    """oracle.lcn_cython (data/lcn/lcn.pyx:16-58 restated in C) against outputs of the reference's own Cython module
    (tests/golden/make_golden_lcn_cython.py): bit for bit, including the zero border and a flat window."""
    g = golden("lcn_cython")
    for n in "abcde":
        ks, eps = g[n + "_args"]
        l, s = oracle.lcn_cython(g[n + "_x"], int(ks), float(eps))
        assert np.array_equal(l, g[n + "_lcn"]) and np.array_equal(s, g[n + "_std"]), n
