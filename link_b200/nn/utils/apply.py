from typing import Callable

import torch

from link_b200.tensor import SparseTensor

__all__ = ['fapply']


def fapply(input: SparseTensor, fn: Callable[..., torch.Tensor], *args, **kwargs) -> SparseTensor:
    """Apply a feature-wise function, keeping coordinates and the shared caches
    (reference: torchsparse/nn/utils/apply.py:10-16)."""
    output = SparseTensor(coords=input.coords, feats=fn(input.feats, *args, **kwargs),
                          stride=input.stride)
    output.cmaps = input.cmaps
    output.kmaps = input.kmaps
    return output
