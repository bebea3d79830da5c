"""spdownsample (TS/torchsparse/nn/functional/downsample.py:11-52): coarse voxel set of a strided convolution.

Hot-path branch (every stride component is 1 or the kernel size, all axes alike): coordinates are truncated to
multiples of stride*tensor_stride inside the key kernel and de-duplicated — one radix sort + run-length pass.
The other two cases exist for the Cylinder3D-style geometries and do their cheap elementwise part
with torch on the device before the same CUDA de-duplication:
  * anisotropic strides, e.g. (2, 2, 1): per-axis truncation (downsample.py:24-27);
  * stride not in {1, kernel}, e.g. kernel 3 / stride 2: offset expansion + divisibility / lower-bound filter (:28-45).
The result is always ordered lexicographically by (b, x, y, z) and stays in original units (:47-51)."""
from typing import Tuple, Union

import torch

from ... import ops
from ...utils import make_ntuple
from ..utils.kernel import get_kernel_offsets

__all__ = ['spdownsample']


def spdownsample(coords: torch.Tensor, stride: Union[int, Tuple[int, ...]] = 2,
                 kernel_size: Union[int, Tuple[int, ...]] = 2,
                 tensor_stride: Union[int, Tuple[int, ...]] = 1) -> torch.Tensor:
    stride, kernel_size, tensor_stride = make_ntuple(stride, 3), make_ntuple(kernel_size, 3), make_ntuple(tensor_stride, 3)
    sample = [stride[k] * tensor_stride[k] for k in range(3)]
    if all(stride[k] in (1, kernel_size[k]) for k in range(3)):
        if len(set(sample)) == 1:
            return ops.unique_coords(coords, trunc_stride=sample[0])
        ss = torch.tensor(sample, dtype=torch.int32, device=coords.device).unsqueeze(0)
        c = coords.to(torch.int32).clone()
        c[:, :3] = torch.div(c[:, :3], ss, rounding_mode='trunc') * ss
        return ops.unique_coords(c.contiguous())
    ss = torch.tensor(sample, dtype=torch.int32, device=coords.device).unsqueeze(0)
    offsets = get_kernel_offsets(kernel_size, tensor_stride, device=coords.device)
    kv = offsets.shape[0]
    c = coords.to(torch.int32)
    cmin = c[:, :3].min(dim=0, keepdim=True).values
    x = (c[:, :3].unsqueeze(1) + offsets.unsqueeze(0)).reshape(-1, 3)
    b = c[:, 3:].repeat(1, kv).reshape(-1, 1)
    mask = ((x % ss == 0) & (x >= cmin)).all(dim=1)
    return ops.unique_coords(torch.cat([x, b], dim=1)[mask].contiguous())
