"""Drop-in for the reference's `torchext` package (torchext/__init__.py:1-4): the same star-exports,
with the custom ops bound to the B200 kernels of libctd_b200.so."""
from . import dataset, worker, functions, modules, ext_cuda, ext_cpu  # noqa: F401
from .dataset import *  # noqa: F401,F403
from .worker import *  # noqa: F401,F403
from .functions import *  # noqa: F401,F403
from .modules import *  # noqa: F401,F403
