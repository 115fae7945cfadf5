"""CPU suite for everything that does not need a GPU: the C-ABI library loads and exports every symbol
include/ctd_b200.h declares, argument validation (no compute), the drop-in torchext API surface and its
error behaviour on CPU tensors, batch sharding and the packed loss reduction over gloo (world size 2)."""
import ctypes
import inspect
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from connecting_the_dots_b200 import _build, _lib
    _build.build()
    return _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "ctd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    names = header_functions()
    assert len(names) >= 26
    L = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libctd_b200.so does not export " + n
    assert sorted(lib.EXPORTS) == names, "python binding and header disagree"
    assert b"sm_100a" in lib.lib().ctd_version()


def test_library_is_sm100a_only_and_torch_free(lib):
    out = subprocess.run(["cuobjdump", "--list-elf", lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    needed = subprocess.run(["readelf", "-d", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in needed and "c10" not in needed


def test_argument_validation_without_gpu(lib):
    """Validation happens before any CUDA call, so it is checkable here."""
    with pytest.raises(lib.CtdError, match="invalid loss type"):
        lib.call("ctd_photometric_fwd_f32", 16, 16, 16, 1, 1, 8, 8, 9, 9, 0.5, None)
    with pytest.raises(lib.CtdError, match="block_size"):
        lib.call("ctd_photometric_bwd_f32", 16, 16, 16, 16, 1, 1, 8, 8, 0, 1, 0.5, None)
    with pytest.raises(lib.CtdError, match="null pointer"):
        lib.call("ctd_photometric_fwd_f32", None, None, None, 1, 1, 8, 8, 9, 1, 0.5, None)
    with pytest.raises(lib.CtdError, match="radius"):
        lib.call("ctd_lcn_f32", 16, 16, 16, 1, 4, 40, 5, 0.05, None)
    with pytest.raises(lib.CtdError, match="negative"):
        lib.call("ctd_nn_f32", 16, 16, 16, -1, 4, None)
    with pytest.raises(lib.CtdError, match="unknown option"):
        lib.set_option("nope", 1)
    # empty problems succeed without touching the device
    lib.call("ctd_photometric_fwd_f32", None, None, None, 0, 1, 8, 8, 9, 1, 0.5, None)
    lib.call("ctd_crosscheck", None, None, None, 0, 0, None)
    lib.call("ctd_xcorrvol_f32", None, None, None, 1, 1, 8, 8, 0, 9, None)


def test_torchext_surface_matches_reference(lib):
    """Names, positional order and defaults of torchext/functions.py:5-147 and torchext/__init__.py:1-4."""
    import connecting_the_dots_b200 as ctd
    tx = ctd.torchext
    sig = lambda f: [(p.name, p.default) for p in inspect.signature(f).parameters.values()]
    E = inspect.Parameter.empty
    assert sig(tx.nn) == [("in0", E), ("in1", E)]
    assert sig(tx.crosscheck) == [("in0", E), ("in1", E)]
    assert sig(tx.proj_nn) == [("xyz0", E), ("xyz1", E), ("K", E), ("patch_size", E)]
    assert sig(tx.xcorrvol) == [("in0", E), ("in1", E), ("n_disps", E), ("block_size", E)]
    assert sig(tx.photometric_loss) == [("es", E), ("ta", E), ("block_size", E), ("type", "mse"), ("eps", 0.1)]
    assert sig(tx.photometric_loss_pytorch) == sig(tx.photometric_loss)
    for cls in ("NNFunction", "CrossCheckFunction", "ProjNNFunction", "XCorrVolFunction", "PhotometricLossFunction"):
        assert issubclass(getattr(tx, cls), torch.autograd.Function)
    for n in ("nn_cuda", "crosscheck_cuda", "proj_nn_cuda", "xcorrvol_cuda", "photometric_loss_forward", "photometric_loss_backward"):
        assert callable(getattr(tx.ext_cuda, n))
    for n in ("nn_cpu", "crosscheck_cpu", "proj_nn_cpu", "xcorrvol_cpu", "photometric_loss_forward", "photometric_loss_backward"):
        assert callable(getattr(tx.ext_cpu, n))
    for n in ("CoordConv2d", "Worker", "TestSets", "TestSet", "MultiDataset", "BaseDataset", "StopWatch", "ETA",
              "PhotometricLoss", "XCorrVol", "CrossCheck", "NN", "ProjNN", "LCN"):
        assert hasattr(tx, n), n
    assert tx.NNFunction.backward(None, None) == (None, None)
    assert tx.ProjNNFunction.backward(None, None) == (None, None, None, None)
    assert tx.XCorrVolFunction.backward(None, None) == (None, None, None, None)
    installed = ctd.install_as_torchext()
    import torchext
    assert torchext is installed and torchext.functions.photometric_loss is tx.photometric_loss
    del sys.modules["torchext"]


def test_no_cpu_fallback(lib):
    import connecting_the_dots_b200 as ctd
    tx = ctd.torchext
    x = torch.randn(1, 1, 12, 12)
    for call in (lambda: tx.photometric_loss(x, x, 9, "sad"), lambda: tx.nn(torch.zeros(2, 3), torch.zeros(2, 3)),
                 lambda: tx.crosscheck(torch.zeros(2, dtype=torch.int64), torch.zeros(2, dtype=torch.int64)),
                 lambda: tx.proj_nn(torch.zeros(1, 2, 2, 3), torch.zeros(1, 2, 2, 3), torch.eye(3), 3),
                 lambda: tx.xcorrvol(x[0], x[0], 4, 3), lambda: tx.lcn(x, 5, 0.05)):
        with pytest.raises(RuntimeError, match="no CPU implementation"):
            call()
    with pytest.raises(Exception, match="invalid loss type"):
        tx.photometric_loss(x, x, 9, "ssim")
    with pytest.raises(NotImplementedError):
        tx.Worker()


def test_torch_restatement_matches_oracle():
    """photometric_loss_pytorch (our counterpart of functions.py:120-147) on CPU against the oracle,
    including an even block size (window offsets [-bs/2, bs/2-1])."""
    import oracle
    import connecting_the_dots_b200 as ctd
    rng = np.random.RandomState(3)
    es = rng.randn(2, 2, 11, 13)
    ta = rng.randn(2, 2, 11, 13)
    for bs in (2, 5, 9):
        for ty in ("mse", "sad", "census_mse", "census_sad"):
            got = ctd.torchext.photometric_loss_pytorch(torch.from_numpy(es), torch.from_numpy(ta), bs, ty, 0.5).numpy()
            want = oracle.photometric_loss_forward(es, ta, bs, ty, 0.5)
            assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), (bs, ty)


def test_dataset_and_timer_shims():
    import connecting_the_dots_b200 as ctd
    tx = ctd.torchext

    class D(tx.BaseDataset):
        def __len__(self):
            return 10

    d = D(train=True)
    a = d.get_rng(3).rand()
    d.current_epoch = 1
    assert d.get_rng(3).rand() != a
    assert D(train=False).get_rng(3).rand() == np.random.RandomState(3).rand()
    m = tx.MultiDataset([1, 2, 3], [4, 5])
    assert len(m) == 5 and [m[i] for i in range(5)] == [1, 2, 3, 4, 5]
    ts = tx.TestSets()
    ts.append("val", [1], test_frequency=2)
    assert ts[0].name == "val" and ts[0].test_frequency == 2
    sw = tx.StopWatch()
    sw.start("a"); sw.stop("a")
    assert sw.get("a") >= 0
    conv = tx.CoordConv2d(1, 2, 3, 1, 1)
    assert conv(torch.zeros(2, 1, 5, 7)).shape == (2, 2, 5, 7)


def test_synthetic_inputs_are_seeded():
    from connecting_the_dots_b200 import synth
    a, b = synth.make_pair(0, 48, 64), synth.make_pair(0, 48, 64)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert abs(synth.dot_pattern().mean() - 0.1) < 0.01
    assert abs(a["go"].sum() - 1) < 1e-4 and a["ta"].std() > 0.3
    batch = synth.make_batch(3, 48, 64, distinct=2)
    assert batch["es"].shape == (3, 1, 48, 64)


def test_shard_range_partitions_any_batch():
    from connecting_the_dots_b200 import shard_range
    for n in (0, 1, 7, 8, 64, 65):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from connecting_the_dots_b200 import ShardedLoss, shard_range
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    diff = torch.from_numpy(rng.rand(8, 1, 6, 5).astype(np.float32))
    mask = torch.from_numpy(rng.rand(8, 1, 6, 5).astype(np.float32))
    lo, hi = shard_range(8, rank, world)
    sl = ShardedLoss("cpu")
    sl.add((mask[lo:hi] * diff[lo:hi]).sum(), mask[lo:hi].sum())
    sl.add((mask[lo:hi] * diff[lo:hi] * 2).sum(), mask[lo:hi].sum())
    vals, _ = sl.reduce()
    want = float((mask * diff).sum() / mask.sum())
    q.put((rank, float(vals[0]), float(vals[1]), want))
    dist.destroy_process_group()


def test_sharded_loss_reduction_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, v0, v1, want in res:
        assert abs(v0 - want) < 1e-6 and abs(v1 - 2 * want) < 1e-6
