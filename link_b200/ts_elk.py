"""Detection-side LinK block: TSELKBlock and the spconv <-> torchsparse adapters.

Reference: detection/det3d/models/utils/ts_elk.py -- spconv2ts / ts2spconv (10-59),
large_to_small / small_to_large_v2 (68-107), TSELKBlock (110-230).  spconv itself is an
un-vendored third-party dependency of the reference; the adapters below are
duck-typed on the five attributes the reference reads (`features`, `indices` [N,4] = (b,z,y,x),
`spatial_shape`, `batch_size`, plus pass-through bookkeeping) and `SparseConvTensor` is the minimal
container with that surface, so the block can be driven without spconv installed."""
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

import link_b200.nn as spnn
from link_b200 import elk
from link_b200.tensor import SparseTensor

__all__ = ['SparseConvTensor', 'spconv2ts', 'ts2spconv', 'large_to_small', 'small_to_large_v2',
           'TSELKBlock']


class BevScatterFunction(torch.autograd.Function):
    """features [n, C], indices int32 [n, 4] (b, z, y, x) -> dense [B, C, D, H, W]; the backward is the
    gather of the same positions (lk_bev_gather)."""

    @staticmethod
    def forward(ctx, feats, indices, batch, d, h, w):
        from link_b200 import _capi
        feats = feats.contiguous()
        n, c = feats.shape
        out = torch.empty(batch, c, d, h, w, dtype=torch.float32, device=feats.device)
        _capi.check(_capi.lib().lk_bev_scatter(_capi.ptr(feats), _capi.ptr(indices), n, c, batch, d, h, w,
                                               _capi.ptr(out), _capi.stream()), 'lk_bev_scatter')
        ctx.save_for_backward(indices)
        ctx.shape = (n, c, batch, d, h, w)
        return out

    @staticmethod
    def backward(ctx, grad):
        from link_b200 import _capi
        (indices,) = ctx.saved_tensors
        n, c, batch, d, h, w = ctx.shape
        g = torch.empty(n, c, dtype=torch.float32, device=grad.device)
        _capi.check(_capi.lib().lk_bev_gather(_capi.ptr(grad.contiguous()), _capi.ptr(indices), n, c, batch, d, h, w,
                                              _capi.ptr(g), _capi.stream()), 'lk_bev_gather')
        return g, None, None, None, None, None


class SparseConvTensor:
    """Minimal stand-in for spconv.pytorch.SparseConvTensor (scn.py:581): features [N,C],
    indices int32 [N,4] = (batch, z, y, x), spatial_shape [D,H,W], batch_size."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None,
                 indice_dict=None, benchmark=False):
        self.features = features
        self.indices = indices
        self.spatial_shape = list(spatial_shape)
        self.batch_size = batch_size
        self.grid = grid
        self.voxel_num = voxel_num
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.benchmark = benchmark
        self.benchmark_record = {}

    def replace_feature(self, feature):
        out = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid,
                               self.voxel_num, self.indice_dict, self.benchmark)
        out.benchmark_record = self.benchmark_record
        return out

    def dense(self, channels_first: bool = True):
        """[B, C, D, H, W] dense tensor (scn.py:614).  CUDA fp32 features take the fused scatter kernel
        (lk_bev_scatter: zero fill + one pass, channels-first written directly; differentiable)."""
        d, h, w = self.spatial_shape
        c = self.features.shape[1]
        if channels_first and self.features.is_cuda and self.features.dtype == torch.float32:
            return BevScatterFunction.apply(self.features, self.indices.int().contiguous(), self.batch_size, d, h, w)
        out = torch.zeros(self.batch_size, d, h, w, c, dtype=self.features.dtype,
                          device=self.features.device)
        i = self.indices.long()
        out[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous() if channels_first else out


def spconv2ts(sct):
    """(b,z,y,x) indices -> (x,y,z,b) coords, stride 1 at every level (ts_elk.py:10-33)."""
    coords = sct.indices[:, [3, 2, 1, 0]].contiguous()
    st = SparseTensor(sct.features, coords, 1)
    save = {k: getattr(sct, k, None) for k in ('batch_size', 'benchmark', 'benchmark_record', 'grid',
                                               'indice_dict', 'spatial_shape', 'voxel_num')}
    save['cls'] = type(sct)
    shape, bs = getattr(sct, 'spatial_shape', None), getattr(sct, 'batch_size', None)
    if shape is not None and bs and coords.is_cuda:
        # indices lie inside the declared grid: bounds from the shape, no min / max read-back per block
        from link_b200.nn.functional import _index
        _index.set_coord_bounds(st.kmaps, (0, 0, 0, 0),
                                (int(shape[2]) - 1, int(shape[1]) - 1, int(shape[0]) - 1, int(bs) - 1))
    return st, save


def ts2spconv(st: SparseTensor, save: Dict[str, Any]):
    """Inverse adapter (ts_elk.py:36-59); rebuilds the same tensor class that came in."""
    indices = st.coords[:, [3, 2, 1, 0]].contiguous()
    sct = save['cls'](st.feats, indices, spatial_shape=save['spatial_shape'],
                      batch_size=save['batch_size'], grid=save['grid'], voxel_num=save['voxel_num'],
                      indice_dict=save['indice_dict'], benchmark=save['benchmark'])
    sct.benchmark_record = save['benchmark_record']
    return sct


def large_to_small(large_x: SparseTensor, stride):
    """== voxel_to_aux (ts_elk.py:68-81)."""
    return elk.voxel_to_aux(large_x, stride)


def small_to_large_v2(small_x, large_x, idx, counts):
    """== aux_to_voxel with the hard-wired 3^3 block neighbourhood (ts_elk.py:84-107)."""
    return elk.aux_to_voxel(small_x, large_x, idx, counts, 3)


class TSELKBlock(nn.Module):
    """LinK block of the detection backbone (ts_elk.py:110-230): `TSELKBlock(inc, outc, baseop)`
    called as `blk(sct, stride)` on an spconv tensor, or `blk.forward_(st, stride)` on a
    SparseTensor.  r = 3 is hard-wired (ts_elk.py:87); 'cos' uses the first inc/2 rows of the
    Linear(3, inc) twice ("channel grouping", ts_elk.py:168)."""

    def __init__(self, inc, outc, baseop='cos'):
        super().__init__()
        self.inc, self.outc, self.baseop = inc, outc, baseop
        self.pre_mix = nn.Sequential(nn.Linear(inc, inc, bias=False), nn.LayerNorm(inc, eps=1e-6))
        self.local_mix = nn.Sequential(spnn.Conv3d(inc, inc, kernel_size=3, dilation=1, stride=1))
        self.pos_weight = nn.Sequential(nn.Linear(3, inc, bias=False))
        self.norm = nn.LayerNorm(inc, eps=1e-6)
        self.norm_local = nn.LayerNorm(inc, eps=1e-6)
        self.activate = nn.ReLU(True)

    def forward(self, sct, stride):
        st, save = spconv2ts(sct)
        return ts2spconv(self.forward_(st, stride), save)

    def _phase_rows(self) -> torch.Tensor:
        w = self.pos_weight[0].weight
        return w[:self.inc // 2] if self.baseop == 'cos' else w

    def forward_(self, st: SparseTensor, stride):
        if self.baseop not in ('sin', 'cos', 'cos_sin', 'x'):
            # 'cos_x_alpha' reads an undefined self.alpha in the reference (ts_elk.py:181)
            raise AttributeError(f"TSELKBlock: baseop {self.baseop!r} is not runnable (reference "
                                 "ts_elk.py:181 references an undefined alpha)")
        needs_grad = torch.is_grad_enabled() and (st.F.requires_grad or
                                                  any(p.requires_grad for p in self.parameters()))
        fused_ok = (self.baseop in ('sin', 'cos') and not needs_grad and st.F.dtype == torch.float32
                    and self.inc % 4 == 0 and self.inc <= 128)
        if fused_ok:
            st.F = elk.elk_forward_fused(st, stride, 3, op=self.baseop, pre_mix=self.pre_mix,
                                         conv=self.local_mix[0], pos_weight=self._phase_rows(),
                                         alpha=None, coord_scale=1.0, norm=self.norm,
                                         norm_local=self.norm_local)
            return st
        if (needs_grad and elk.FUSED_BACKWARD and self.baseop in ('sin', 'cos') and st.F.dtype == torch.float32
                and self.inc in (16, 32, 64, 128)):
            # training: fused forward + hand-written backward of the linear-kernel path
            F_input, local_mix = self.pre_mix(st.F), self.local_mix(st)
            st.F = elk.LinkAggregateFunction.apply(
                F_input, local_mix.F, self._phase_rows(), self.norm.weight, self.norm.bias,
                self.norm_local.weight, self.norm_local.bias, st.C.contiguous(), elk.block_index(st, stride), 3,
                self.baseop)
            return st
        return self._forward_composed(st, stride)

    def _forward_composed(self, st: SparseTensor, stride):
        """The reference's op sequence (ts_elk.py:144-230) on differentiable kernels."""
        C_ = self.inc
        F_input = self.pre_mix(st.F)
        local_mix = self.local_mix(st)
        pos = self.pos_weight(st.C[:, :3].float())
        if self.baseop in ('cos', 'x'):
            pos = pos[:, :C_ // 2].repeat([1, 2])
        sin, cos = torch.sin(pos), torch.cos(pos)
        if self.baseop == 'sin':
            planes = [F_input * sin, F_input * cos]
        elif self.baseop in ('cos', 'cos_sin'):
            planes = [F_input * cos, F_input * sin]
        else:
            lin = F_input * pos
            planes = [lin]
        st.F = torch.cat(planes, dim=1).contiguous()
        small_st, idx, counts = large_to_small(st, stride)
        large_st = small_to_large_v2(small_st, st, idx, counts)
        vf = large_st.F
        if self.baseop == 'sin':
            new = vf[:, :C_] * cos - vf[:, C_:] * sin
        elif self.baseop == 'cos':
            new = vf[:, :C_] * cos + vf[:, C_:] * sin
        elif self.baseop == 'cos_sin':
            new = (vf[:, :C_] * cos + vf[:, C_:] * sin) + (vf[:, C_:] * cos - vf[:, :C_] * sin)
        else:
            new = vf - lin
        large_st.F = self.activate(self.norm(new) + self.norm_local(local_mix.F))
        return large_st
