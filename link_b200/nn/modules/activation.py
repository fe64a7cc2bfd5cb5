from torch import nn

from link_b200.nn.utils import fapply
from link_b200.tensor import SparseTensor

__all__ = ['ReLU', 'LeakyReLU']


class ReLU(nn.ReLU):

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)


class LeakyReLU(nn.LeakyReLU):

    def forward(self, input: SparseTensor) -> SparseTensor:
        return fapply(input, super().forward)
