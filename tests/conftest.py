import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def ref_ext():
    """The unmodified reference CPU extension (oracle/_ref), or skip when it was not shipped."""
    import build_ref
    mod = build_ref.load_ref()
    if mod is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return mod


@pytest.fixture(scope="session")
def tx():
    """The drop-in torchext package, with the kernel library built and loaded (fails loudly if not)."""
    import os
    import connecting_the_dots_b200 as ctd
    ctd._lib.lib()
    # CTD_TEST_OPTIONS="census_stream=1,host_graphs=1": run the whole suite with opt-in paths switched on (tests that set
    # an option themselves put the library default back afterwards)
    for item in filter(None, os.environ.get("CTD_TEST_OPTIONS", "").split(",")):
        name, value = item.split("=")
        ctd._lib.set_option(name.strip(), int(value))
    return ctd.torchext


def assert_close(got, ref, tol=1e-5, what=""):
    """The floating-point comparator of this repo: max |got - ref| <= tol * max |ref|.
    Gradients have mixed signs and near-zero entries, so the scale is the tensor's, not the element's."""
    import numpy as np
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()) if ref.size else 0.0, 1e-30)
    err = float(np.abs(got - ref).max()) if ref.size else 0.0
    assert np.isfinite(got).all(), what + ": non-finite values"
    assert err <= tol * scale, "%s: max abs err %.3e > %.0e * max|ref| (%.3e)" % (what, err, tol, scale)
    return err / scale
