"""sparse_quantize (TS/torchsparse/utils/quantize.py:24-46) on the device.

Voxels come out sorted lexicographically by (x, y, z); `indices` is the FIRST point of each voxel in input order and
`inverse_indices` maps every point to its voxel — np.unique's contract, obtained here from one stable radix sort of
packed coordinate keys plus a run-length pass.  numpy in -> numpy out (the reference's DataLoader call sites,
R/pcseg/data/dataset/semantickitti/semantickitti_voxel_ms.py:153-165); CUDA tensors in -> CUDA tensors out.
numpy input is moved to the current CUDA device: DataLoader workers that call this must be started with the `spawn`
method (a forked worker cannot initialise CUDA) — or, better, leave quantisation to the device front end
(taseg_b200.frontend), which is what the benchmark path does.
"""
from itertools import repeat
from typing import List, Tuple, Union

import numpy as np
import torch

from .. import ops

__all__ = ['sparse_quantize']


def sparse_quantize(coords, voxel_size: Union[float, Tuple[float, ...]] = 1, *, return_index: bool = False,
                    return_inverse: bool = False) -> List:
    if isinstance(voxel_size, (float, int)):
        voxel_size = tuple(repeat(voxel_size, 3))
    assert isinstance(voxel_size, tuple) and len(voxel_size) == 3
    as_numpy = isinstance(coords, np.ndarray)
    c = torch.from_numpy(np.ascontiguousarray(coords)).cuda() if as_numpy else coords
    if c.dtype.is_floating_point:       # the reference always floors: coords = np.floor(coords / voxel_size).astype(np.int32)
        vs = torch.tensor(voxel_size, dtype=torch.float64, device=c.device)
        c = torch.floor(c.to(torch.float64) / vs)
    elif any(v != 1 for v in voxel_size):
        vs = torch.tensor(voxel_size, dtype=torch.float64, device=c.device)
        c = torch.floor(c.to(torch.float64) / vs)
    c = c.to(torch.int32)
    c4 = torch.cat([c[:, :3], torch.zeros((c.shape[0], 1), dtype=torch.int32, device=c.device)], dim=1)
    uniq, first, inv = ops.unique_coords(c4, want_index=True, want_inverse=True)
    outs = [uniq[:, :3].contiguous()]
    if return_index:
        outs.append(first.long())
    if return_inverse:
        outs.append(inv.long())
    if as_numpy:
        outs = [o.cpu().numpy() for o in outs]
    return outs[0] if len(outs) == 1 else outs
