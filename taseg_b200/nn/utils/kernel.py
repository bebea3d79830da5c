"""get_kernel_offsets (TS/torchsparse/nn/utils/kernel.py:11-32).

The enumeration order defines which weight slice W[k] meets which offset: odd kernel volume -> x fastest
(z outermost), even volume -> z fastest (x outermost)."""
from itertools import product
from typing import Tuple, Union

import numpy as np
import torch

from ...utils import make_ntuple

__all__ = ['get_kernel_offsets', 'kernel_offsets_np']


def kernel_offsets_np(size, stride=1, dilation=1) -> np.ndarray:
    size, stride, dilation = make_ntuple(size, 3), make_ntuple(stride, 3), make_ntuple(dilation, 3)
    axes = [np.arange(-size[d] // 2 + 1, size[d] // 2 + 1) * stride[d] * dilation[d] for d in range(3)]
    if int(np.prod(size)) % 2 == 1:
        rows = [(x, y, z) for z, y, x in product(axes[2], axes[1], axes[0])]
    else:
        rows = [(x, y, z) for x, y, z in product(axes[0], axes[1], axes[2])]
    return np.asarray(rows, dtype=np.int32).reshape(-1, 3)


def get_kernel_offsets(size: Union[int, Tuple[int, ...]], stride: Union[int, Tuple[int, ...]] = 1,
                       dilation: Union[int, Tuple[int, ...]] = 1, device: str = 'cpu') -> torch.Tensor:
    return torch.from_numpy(kernel_offsets_np(size, stride, dilation)).to(device)
