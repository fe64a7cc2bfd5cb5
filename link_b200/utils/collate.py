"""Batch collation with the reference's layout (torchsparse/utils/collate.py:11-66):
the batch index is appended as the 4th coordinate column."""
from typing import Any, List

import numpy as np
import torch

from link_b200.tensor import SparseTensor

__all__ = ['sparse_collate', 'sparse_collate_fn']


def sparse_collate(inputs: List[SparseTensor]) -> SparseTensor:
    coords, feats = [], []
    stride = inputs[0].stride
    for k, x in enumerate(inputs):
        if isinstance(x.coords, np.ndarray):
            x.coords = torch.tensor(x.coords)
        if isinstance(x.feats, np.ndarray):
            x.feats = torch.tensor(x.feats)
        assert isinstance(x.coords, torch.Tensor), type(x.coords)
        assert isinstance(x.feats, torch.Tensor), type(x.feats)
        assert x.stride == stride, (x.stride, stride)
        batch = torch.full((x.coords.shape[0], 1), k, device=x.coords.device, dtype=torch.int)
        coords.append(torch.cat((x.coords, batch), dim=1))
        feats.append(x.feats)
    return SparseTensor(coords=torch.cat(coords, dim=0), feats=torch.cat(feats, dim=0),
                        stride=stride)


def sparse_collate_fn(inputs: List[Any]) -> Any:
    if not isinstance(inputs[0], dict):
        return inputs
    output = {}
    for name in inputs[0].keys():
        first = inputs[0][name]
        if isinstance(first, dict):
            output[name] = sparse_collate_fn([inp[name] for inp in inputs])
        elif isinstance(first, np.ndarray):
            output[name] = torch.stack([torch.tensor(inp[name]) for inp in inputs], dim=0)
        elif isinstance(first, torch.Tensor):
            output[name] = torch.stack([inp[name] for inp in inputs], dim=0)
        elif isinstance(first, SparseTensor):
            output[name] = sparse_collate([inp[name] for inp in inputs])
        else:
            output[name] = [inp[name] for inp in inputs]
    return output
