"""connecting_the_dots_b200 -- B200 (sm_100a) kernels for the per-pixel custom ops of
"Connecting the Dots" behind the reference's own `torchext` op API.

    import connecting_the_dots_b200 as ctd
    ctd.torchext.photometric_loss(es, ta, 9, 'census_sad', 0.5)      # reference API, CUDA tensors
    ctd.install_as_torchext()                                         # `import torchext` -> this package

Layers: csrc/ (CUDA kernels + C ABI, include/ctd_b200.h)  ->  _lib.py (ctypes binding of the C ABI)
->  torchext/ext_cuda.py (the reference's native entry points, same names and checks)  ->
torchext/functions.py (the reference's autograd functions).  There is no CPU compute path.
"""
import sys

from . import _lib  # noqa: F401
from . import torchext  # noqa: F401
from .sharding import shard_range, ShardedLoss  # noqa: F401

__all__ = ["torchext", "install_as_torchext", "shard_range", "ShardedLoss"]


def install_as_torchext():
    """Make `import torchext` resolve to this package's drop-in (model/networks.py:9 does that)."""
    sys.modules["torchext"] = torchext
    for name in ("functions", "modules", "ext_cuda", "ext_cpu", "dataset", "worker"):
        sys.modules["torchext." + name] = getattr(torchext, name)
    return torchext
