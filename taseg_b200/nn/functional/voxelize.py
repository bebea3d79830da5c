"""spvoxelize (TS/torchsparse/nn/functional/voxelize.py:10-56): scatter-mean of point features into voxels."""
import torch
from torch.autograd import Function

from ... import ops

__all__ = ['spvoxelize']


class VoxelizeFunction(Function):

    @staticmethod
    def forward(ctx, feats: torch.Tensor, coords: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
        coords = coords.int()
        ctx.for_backwards = (coords, counts, feats.shape[0])
        return ops.voxelize_forward(feats, coords, counts)

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        coords, counts, n = ctx.for_backwards
        return ops.voxelize_backward(grad_output, coords, counts, n), None, None


def spvoxelize(feats: torch.Tensor, coords: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    if torch.is_autocast_enabled():
        feats = feats.to(torch.get_autocast_dtype('cuda'))   # the reference casts to half under --amp (voxelize.py:13)
    return VoxelizeFunction.apply(feats, coords, counts)
