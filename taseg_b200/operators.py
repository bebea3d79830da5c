"""torchsparse.cat (TS/torchsparse/operators.py:10-17): channel concat of tensors on the same voxels."""
from typing import List

import torch

from .tensor import SparseTensor

__all__ = ['cat']


def cat(inputs: List[SparseTensor]) -> SparseTensor:
    return inputs[0].derive(torch.cat([x.feats for x in inputs], dim=1))
