"""taseg_b200 — B200-native (sm_100a) implementation of TASeg's sparse-convolution hot path behind the
torchsparse 1.4.0 operator surface.

    import taseg_b200; taseg_b200.install_as_torchsparse()   # then `import torchsparse` resolves to this package

or put taseg_b200/dropin on PYTHONPATH.  All compute runs in hand-written CUDA kernels reached through the C ABI
of include/taseg_b200.h (libtaseg_b200.so); there is no CPU or eager fallback.
"""
import sys

from .operators import *
from .tensor import *
from . import nn, utils, backend
from .utils import collate, quantize  # noqa: F401  (submodules pcseg imports by path)

__version__ = '1.4.0+b200'

_ALIASES = ['', '.tensor', '.operators', '.backend', '.nn', '.nn.functional', '.nn.modules', '.nn.utils',
            '.nn.functional.conv', '.nn.functional.hash', '.nn.functional.query', '.nn.functional.count',
            '.nn.functional.voxelize', '.nn.functional.devoxelize', '.nn.functional.downsample',
            '.nn.functional.activation', '.nn.modules.conv', '.nn.modules.activation', '.nn.modules.norm',
            '.nn.utils.apply', '.nn.utils.kernel', '.utils', '.utils.utils', '.utils.quantize', '.utils.collate']


def install_as_torchsparse() -> None:
    """Register this package under the name `torchsparse` (and every submodule pcseg imports)."""
    import importlib
    for suffix in _ALIASES:
        sys.modules['torchsparse' + suffix] = importlib.import_module(__name__ + suffix)
