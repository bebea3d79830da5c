"""spdevoxelize / calc_ti_weights (TS/torchsparse/nn/functional/devoxelize.py:10-98)."""
import torch
from torch.autograd import Function

from ... import ops

__all__ = ['spdevoxelize', 'calc_ti_weights']


def calc_ti_weights(coords: torch.Tensor, idx_query: torch.Tensor, scale: float = 1) -> torch.Tensor:
    """Trilinear weights (8,N) of each point w.r.t. the 8 corners of its stride-`scale` cell, zeroed where the corner
    voxel is absent and renormalised by (sum + 1e-8).  Corner order: z fastest (get_kernel_offsets(2, s))."""
    with torch.no_grad():
        p = coords[:, :3]
        lo = torch.floor(p / scale) * scale if scale != 1 else torch.floor(p)
        hi = lo + scale
        near = (hi - p).float()      # weight toward the low corner along each axis
        far = (p - lo).float()
        rows = []
        for k in range(8):
            wx = far[:, 0] if k & 4 else near[:, 0]
            wy = far[:, 1] if k & 2 else near[:, 1]
            wz = far[:, 2] if k & 1 else near[:, 2]
            rows.append(wx * wy * wz)
        w = torch.stack(rows, dim=0)
        if scale != 1:
            w /= scale ** 3
        w[idx_query == -1] = 0
        w /= torch.sum(w, dim=0) + 1e-8
    return w


class DevoxelizeFunction(Function):

    @staticmethod
    def forward(ctx, feats: torch.Tensor, coords: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        coords = coords.int()
        ctx.for_backwards = (coords, weights, feats.shape[0])
        return ops.devoxelize_forward(feats, coords, weights)

    @staticmethod
    def backward(ctx, grad_output: torch.Tensor):
        coords, weights, m = ctx.for_backwards
        return ops.devoxelize_backward(grad_output, coords, weights, m), None, None


def spdevoxelize(feats: torch.Tensor, coords: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    if torch.is_autocast_enabled():
        feats = feats.to(torch.get_autocast_dtype('cuda'))   # devoxelize.py:54
    return DevoxelizeFunction.apply(feats, coords, weights)
