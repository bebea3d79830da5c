"""spdownsample (TS/torchsparse/nn/functional/downsample.py:11-52): coarse voxel set of a strided convolution.

Only the branch the hot path takes is built (every stride component is 1 or the kernel size): coordinates are
truncated to multiples of stride*tensor_stride and de-duplicated; the result is ordered lexicographically by
(b, x, y, z) and stays in original units.  One radix sort + run-length pass on the device."""
from typing import Tuple, Union

import torch

from ... import ops
from ...utils import make_ntuple

__all__ = ['spdownsample']


def spdownsample(coords: torch.Tensor, stride: Union[int, Tuple[int, ...]] = 2,
                 kernel_size: Union[int, Tuple[int, ...]] = 2,
                 tensor_stride: Union[int, Tuple[int, ...]] = 1) -> torch.Tensor:
    stride, kernel_size, tensor_stride = make_ntuple(stride, 3), make_ntuple(kernel_size, 3), make_ntuple(tensor_stride, 3)
    if not all(stride[k] in (1, kernel_size[k]) for k in range(3)):
        raise NotImplementedError('spdownsample: the offset-expansion branch (stride not in {1, kernel}) is outside the hot path')
    sample = [stride[k] * tensor_stride[k] for k in range(3)]
    if len(set(sample)) != 1:
        raise NotImplementedError('spdownsample: anisotropic strides (Cylinder3D) are outside the hot path')
    return ops.unique_coords(coords, trunc_stride=sample[0])
