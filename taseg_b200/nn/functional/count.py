"""spcount (TS/torchsparse/nn/functional/count.py:8-16)."""
import torch

from ... import ops

__all__ = ['spcount']


def spcount(coords: torch.Tensor, num) -> torch.Tensor:
    return ops.spcount(coords, int(num))
