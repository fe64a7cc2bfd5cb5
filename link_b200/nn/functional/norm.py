"""Training-mode BatchNorm over sparse feature rows, fused with the shortcut add and the ReLU that
follow it (csrc/bn.cu: lk_bn_train_fwd / lk_bn_train_bwd).

Reference: spnn.BatchNorm is nn.BatchNorm1d over `.feats` (torchsparse/nn/modules/norm.py:10-13),
followed by spnn.ReLU or by the ResidualBlock's add + ReLU (linkencoder.py:26-37, 64-91); the
detection backbone applies nn.BatchNorm1d + nn.ReLU to `.features` (scn.py:64-107).  Same values,
same running-statistics updates, one autograd node instead of three."""
from typing import Optional

import torch
from torch.autograd import Function

from link_b200 import _capi

__all__ = ['batch_norm_act', 'bn_act_supported']

USE_FUSED_BN = __import__('os').environ.get('LINKB200_FUSED_BN', '1') != '0'


def bn_act_supported(bn: torch.nn.modules.batchnorm._BatchNorm, x: torch.Tensor) -> bool:
    """The fused kernels serve nn.BatchNorm1d's default configuration in TRAINING mode: batch statistics,
    affine, running statistics with a fixed momentum, fp32 CUDA rows, >= 2 rows (nn.BatchNorm1d raises
    for a single value per channel; that check stays with PyTorch)."""
    return (USE_FUSED_BN and bn.training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
            and x.shape[0] >= 2 and bn.affine and bn.track_running_stats and bn.momentum is not None
            and x.shape[1] % 4 == 0 and x.shape[1] <= 1024 and bn.weight.dtype == torch.float32)


class BatchNormActFunction(Function):
    """y = relu?(BN_train(x) [+ residual]); backward through lk_bn_train_bwd."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, nbt, eps, momentum, relu):
        _capi.check_device(x)
        x = x.contiguous()
        n, c = x.shape
        L = _capi.lib()
        res = residual.contiguous() if residual is not None else None
        y = torch.empty_like(x)
        stats = torch.empty(2, c, dtype=torch.float32, device=x.device)          # mean, invstd
        ws = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        _capi.check(L.lk_bn_train_fwd(_capi.ptr(x), _capi.ptr(res), n, c, _capi.ptr(weight), _capi.ptr(bias), eps,
                                      momentum, 1 if relu else 0, _capi.ptr(running_mean), _capi.ptr(running_var),
                                      _capi.ptr(nbt), stats[0].data_ptr(), stats[1].data_ptr(), _capi.ptr(y),
                                      ws.data_ptr(), 16 * c, _capi.stream()), 'lk_bn_train_fwd')
        ctx.save_for_backward(x, y if relu else None, weight, stats)
        ctx.relu, ctx.has_res = relu, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, stats = ctx.saved_tensors
        n, c = x.shape
        dy = dy.contiguous()
        if dy.dtype != torch.float32:
            dy = dy.float()
        L = _capi.lib()
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if (ctx.has_res and ctx.needs_input_grad[3]) else None
        dwb = torch.empty(2, c, dtype=torch.float32, device=x.device)
        ws = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        _capi.check(L.lk_bn_train_bwd(_capi.ptr(dy), _capi.ptr(x), _capi.ptr(y), n, c, stats[0].data_ptr(),
                                      stats[1].data_ptr(), _capi.ptr(weight), _capi.ptr(dx), _capi.ptr(dres),
                                      dwb[0].data_ptr(), dwb[1].data_ptr(), ws.data_ptr(), 16 * c, _capi.stream()),
                    'lk_bn_train_bwd')
        return dx, dwb[0], dwb[1], dres, None, None, None, None, None, None


def batch_norm_act(x: torch.Tensor, bn, relu: bool = False, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """relu?(bn(x) [+ residual]) for a BatchNorm module `bn`: the fused kernels in training mode on
    fp32 CUDA rows (see bn_act_supported), PyTorch's ops otherwise (eval mode, other dtypes, CPU)."""
    if bn_act_supported(bn, x) and (residual is None or (residual.dtype == torch.float32 and residual.shape == x.shape)):
        from link_b200 import _ext
        ext = _ext.module()
        if ext is not None:      # the same two kernels behind a C++ autograd node (no python in the backward)
            _capi.check_device(x)
            return ext.batch_norm_act(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var,
                                      bn.num_batches_tracked, float(bn.eps), float(bn.momentum), bool(relu))
        return BatchNormActFunction.apply(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var,
                                          bn.num_batches_tracked, float(bn.eps), float(bn.momentum), bool(relu))
    y = torch.nn.modules.batchnorm._BatchNorm.forward(bn, x)       # (not bn(x): spnn.BatchNorm takes SparseTensors)
    if residual is not None:
        y = y + residual
    return torch.relu(y) if relu else y
