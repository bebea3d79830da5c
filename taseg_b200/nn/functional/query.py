"""sphashquery (TS/torchsparse/nn/functional/query.py:8-33): position of each query key in `references`, -1 on miss."""
import torch

from ... import ops

__all__ = ['sphashquery']


def sphashquery(queries: torch.Tensor, references: torch.Tensor) -> torch.Tensor:
    shape = queries.shape
    out = ops.sphashquery(queries.to(torch.int64).reshape(-1), references.to(torch.int64).reshape(-1))
    return out.view(*shape)
