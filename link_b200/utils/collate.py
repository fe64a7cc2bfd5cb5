"""Batch collation for data loaders.

Contract = the reference's `sparse_collate` / `sparse_collate_fn` (torchsparse/utils/collate.py:11-66):
a list of per-frame SparseTensors with [N_i, 3] coordinates becomes ONE SparseTensor whose 4th
coordinate column is the frame index (the column the hash / kernel-map / block kernels treat as the
batch), and a list of sample dicts is merged key by key (nested dicts recursively, arrays and tensors
stacked, SparseTensors collated, everything else kept as a list)."""
from typing import Any, Dict, List

import numpy as np
import torch

from link_b200.tensor import SparseTensor

__all__ = ['sparse_collate', 'sparse_collate_fn']


def _as_tensor(x) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(x)).clone()
    if not isinstance(x, torch.Tensor):
        raise AssertionError(type(x))
    return x


def sparse_collate(inputs: List[SparseTensor]) -> SparseTensor:
    stride = inputs[0].stride
    sizes = []
    for x in inputs:
        if x.stride != stride:
            raise AssertionError((x.stride, stride))
        x.coords, x.feats = _as_tensor(x.coords), _as_tensor(x.feats)   # the reference converts in place too
        sizes.append(x.coords.shape[0])
    total, dev = sum(sizes), inputs[0].coords.device
    # one [total, 4] buffer: xyz blocks copied in, the frame index written as a run per frame
    coords = torch.empty(total, 4, dtype=inputs[0].coords.dtype, device=dev)
    row = 0
    for frame, (x, n) in enumerate(zip(inputs, sizes)):
        coords[row:row + n, :3] = x.coords[:, :3]
        coords[row:row + n, 3] = frame
        row += n
    feats = torch.cat([x.feats for x in inputs], dim=0)
    return SparseTensor(coords=coords, feats=feats, stride=stride)


def _merge(values: List[Any]) -> Any:
    head = values[0]
    if isinstance(head, dict):
        return sparse_collate_fn(values)
    if isinstance(head, np.ndarray):
        return torch.stack([torch.from_numpy(np.ascontiguousarray(v)).clone() for v in values], dim=0)
    if isinstance(head, torch.Tensor):
        return torch.stack(values, dim=0)
    if isinstance(head, SparseTensor):
        return sparse_collate(values)
    return values


def sparse_collate_fn(inputs: List[Any]) -> Any:
    if not inputs or not isinstance(inputs[0], dict):
        return inputs
    merged: Dict[Any, Any] = {}
    for name in inputs[0]:
        merged[name] = _merge([sample[name] for sample in inputs])
    return merged
