import torch
from torch import nn

from link_b200.nn.utils import fapply
from link_b200.tensor import SparseTensor

__all__ = ['BatchNorm', 'GroupNorm']


class BatchNorm(nn.BatchNorm1d):
    """BatchNorm1d over the feature rows (reference: torchsparse/nn/modules/norm.py:10-13).  In
    training mode on fp32 CUDA rows the statistics / normalisation / backward run on the fused
    kernels of csrc/bn.cu (nn.functional.batch_norm_act); same values and running statistics."""

    def forward(self, input: SparseTensor) -> SparseTensor:
        from link_b200.nn.functional.norm import batch_norm_act
        return fapply(input, batch_norm_act, self)


class GroupNorm(nn.GroupNorm):
    """Per-sample GroupNorm (reference: norm.py:16-43).  Not used by the LinK models."""

    def forward(self, input: SparseTensor) -> SparseTensor:
        coords, feats = input.coords, input.feats
        batch_size = int(torch.max(coords[:, -1]).item()) + 1
        nfeats = torch.zeros_like(feats)
        for k in range(batch_size):
            sel = coords[:, -1] == k
            b = feats[sel].transpose(0, 1).reshape(1, feats.shape[1], -1)
            b = super().forward(b)
            nfeats[sel] = b.reshape(feats.shape[1], -1).transpose(0, 1)
        output = SparseTensor(coords=coords, feats=nfeats, stride=input.stride)
        output.cmaps = input.cmaps
        output.kmaps = input.kmaps
        return output
