"""sphash (TS/torchsparse/nn/functional/hash.py:10-37) on the sm_100a hash kernels."""
from typing import Optional

import torch

from ... import ops

__all__ = ['sphash']


def sphash(coords: torch.Tensor, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert coords.dtype == torch.int, coords.dtype
    assert coords.ndim == 2 and coords.shape[1] == 4, coords.shape
    if offsets is not None:
        assert offsets.dtype == torch.int, offsets.dtype
        assert offsets.ndim == 2 and offsets.shape[1] == 3, offsets.shape
        offsets = offsets.to(coords.device)
    return ops.sphash(coords, offsets)
