"""make_ntuple — same contract as TS/torchsparse/utils/utils.py:9-20."""
from typing import List, Tuple, Union

import torch

__all__ = ['make_ntuple']


def make_ntuple(x: Union[int, List[int], Tuple[int, ...], torch.Tensor], ndim: int) -> Tuple[int, ...]:
    if isinstance(x, torch.Tensor):
        x = [int(v) for v in x.reshape(-1).tolist()]
    if isinstance(x, int):
        x = (x,) * ndim
    x = tuple(x)
    assert len(x) == ndim, x
    return x
