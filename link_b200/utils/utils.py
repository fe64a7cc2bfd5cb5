from typing import List, Tuple, Union

import torch

__all__ = ['make_ntuple']


def make_ntuple(x: Union[int, List[int], Tuple[int, ...], torch.Tensor], ndim: int) -> Tuple[int, ...]:
    """Same contract as the reference helper (torchsparse/utils/utils.py:9-21)."""
    if isinstance(x, torch.Tensor):
        x = x.reshape(-1).tolist()
    if isinstance(x, int):
        return (x,) * ndim
    x = tuple(int(v) for v in x)
    assert len(x) == ndim, x
    return x
