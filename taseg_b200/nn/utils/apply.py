"""fapply (TS/torchsparse/nn/utils/apply.py:10-16)."""
from typing import Callable

import torch

from ...tensor import SparseTensor

__all__ = ['fapply']


def fapply(input: SparseTensor, fn: Callable[..., torch.Tensor], *args, **kwargs) -> SparseTensor:
    return input.derive(fn(input.feats, *args, **kwargs))
