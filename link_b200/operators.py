from typing import List

import torch

from link_b200.tensor import SparseTensor

__all__ = ['cat']


def cat(inputs: List[SparseTensor]) -> SparseTensor:
    """Channel-wise concatenation of tensors on the same coordinates
    (reference: torchsparse/operators.py:10-17)."""
    output = SparseTensor(coords=inputs[0].coords,
                          feats=torch.cat([x.feats for x in inputs], dim=1),
                          stride=inputs[0].stride)
    output.cmaps = inputs[0].cmaps
    output.kmaps = inputs[0].kmaps
    return output
