"""sparse_collate / sparse_collate_fn (TS/torchsparse/utils/collate.py:11-59): batch index as 4th coordinate column."""
from typing import Any, List

import numpy as np
import torch

from ..tensor import SparseTensor

__all__ = ['sparse_collate', 'sparse_collate_fn']


def sparse_collate(inputs: List[SparseTensor]) -> SparseTensor:
    stride = inputs[0].stride
    coords, feats = [], []
    for b, x in enumerate(inputs):
        if isinstance(x.coords, np.ndarray):
            x.coords = torch.tensor(x.coords)
        if isinstance(x.feats, np.ndarray):
            x.feats = torch.tensor(x.feats)
        assert isinstance(x.coords, torch.Tensor), type(x.coords)
        assert isinstance(x.feats, torch.Tensor), type(x.feats)
        assert x.stride == stride, (x.stride, stride)
        col = torch.full((x.coords.shape[0], 1), b, device=x.coords.device, dtype=torch.int)
        coords.append(torch.cat((x.coords, col), dim=1))
        feats.append(x.feats)
    return SparseTensor(coords=torch.cat(coords, dim=0), feats=torch.cat(feats, dim=0), stride=stride)


def sparse_collate_fn(inputs: List[Any]) -> Any:
    if not isinstance(inputs[0], dict):
        return inputs
    output = {}
    for name, first in inputs[0].items():
        column = [sample[name] for sample in inputs]
        if isinstance(first, dict):
            output[name] = sparse_collate_fn(column)
        elif isinstance(first, np.ndarray):
            output[name] = torch.stack([torch.tensor(v) for v in column], dim=0)
        elif isinstance(first, torch.Tensor):
            output[name] = torch.stack(column, dim=0)
        elif isinstance(first, SparseTensor):
            output[name] = sparse_collate(column)
        else:
            output[name] = column
    return output
