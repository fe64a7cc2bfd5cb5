import math
from typing import Tuple, Union

import numpy as np
import torch
from torch import nn

from link_b200.nn import functional as F
from link_b200.tensor import SparseTensor
from link_b200.utils import make_ntuple

__all__ = ['Conv3d']


class Conv3d(nn.Module):
    """Sparse 3D convolution module; parameter names, shapes ([K, Cin, Cout], or [Cin, Cout]
    when K == 1) and initialisation follow the reference (torchsparse/nn/modules/conv.py:15-72)
    so its state dicts load unchanged."""

    def __init__(self, in_channels: int, out_channels: int,
                 kernel_size: Union[int, Tuple[int, ...]] = 3,
                 stride: Union[int, Tuple[int, ...]] = 1, dilation: int = 1, bias: bool = False,
                 transposed: bool = False) -> None:
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = make_ntuple(kernel_size, ndim=3)
        self.stride = make_ntuple(stride, ndim=3)
        self.dilation = dilation
        self.transposed = transposed
        self.kernel_volume = int(np.prod(self.kernel_size))
        shape = ((self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1
                 else (in_channels, out_channels))
        self.kernel = nn.Parameter(torch.zeros(*shape))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def extra_repr(self) -> str:
        s = f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}'
        if self.stride != (1,) * len(self.stride):
            s += f', stride={self.stride}'
        if self.dilation != 1:
            s += f', dilation={self.dilation}'
        if self.bias is None:
            s += ', bias=False'
        if self.transposed:
            s += ', transposed=True'
        return s

    def reset_parameters(self) -> None:
        fan = (self.out_channels if self.transposed else self.in_channels) * self.kernel_volume
        std = 1 / math.sqrt(fan)
        self.kernel.data.uniform_(-std, std)
        if self.bias is not None:
            self.bias.data.uniform_(-std, std)

    def forward(self, input: SparseTensor) -> SparseTensor:
        return F.conv3d(input, self.kernel, kernel_size=self.kernel_size, bias=self.bias,
                        stride=self.stride, dilation=self.dilation, transposed=self.transposed)
