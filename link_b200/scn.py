"""spconv-free detection backbone: SubMConv3d / SparseConv3d / SparseSequential /
SparseBasicBlock and SpMiddleResNetFHDELKv3 on liblinkb200 kernels.

Reference: detection/det3d/models/backbones/scn.py (SparseBasicBlock :64-107,
SpMiddleResNetFHDELKv3 :453-626).  The reference builds these layers from `spconv.pytorch`, an
un-vendored third-party package; here they are restated on our
kernel-map + sparse-conv kernels:

  * tensors are `link_b200.ts_elk.SparseConvTensor` (features, indices (b,z,y,x), spatial_shape);
  * SubMConv3d: output sites == input sites, out[o] = sum_kappa W[kappa] in[o + kappa - centre];
  * SparseConv3d(k, s, p): out shape floor((in + 2p - k)/s) + 1 per axis; site o is active iff some
    input i and tap kappa satisfy i + p - kappa = o*s; rows are ordered by (b, z, y, x)
    (spconv's own row order is implementation defined);
  * parameters keep spconv-2.x names and layout (`weight` [C_out, kz, ky, kx, C_in], `bias`), so
    reference checkpoints load (det3d/torchie/trainer/checkpoint.py:78-92).
Parity: spconv is not installable here, so these layers are checked against dense
torch.nn.functional.conv3d on the densified input (tests/test_gpu_parity.py)."""
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn

from link_b200 import _capi
from link_b200.nn.functional import _index
from link_b200.nn.functional.conv import KernelMap, ConvolutionFunction, _conv_fwd, _folded_bn
from link_b200.nn.functional.hash import sphash
from link_b200.nn.functional.norm import batch_norm_act
from link_b200.nn.functional.query import HashTable
from link_b200.ts_elk import SparseConvTensor, TSELKBlock

__all__ = ['SubMConv3d', 'SparseConv3d', 'SparseSequential', 'SparseBasicBlock',
           'SpMiddleResNetFHDELKv3', 'SparseConvTensor']


def _triple(v) -> Tuple[int, int, int]:
    if isinstance(v, int):
        return (v, v, v)
    v = tuple(int(a) for a in v)
    assert len(v) == 3, v
    return v


def _xyzb(indices: torch.Tensor) -> torch.Tensor:
    return indices[:, [3, 2, 1, 0]].contiguous()


def _taps_xyz(kernel: Tuple[int, int, int], shift: Tuple[int, int, int], device) -> torch.Tensor:
    """int32 [K,3] (x,y,z) offsets kappa - shift, kappa enumerated (kz, ky, kx) row-major so that
    tap K-index == flattened spconv weight index."""
    kz, ky, kx = kernel
    offs = [[x - shift[2], y - shift[1], z - shift[0]]
            for z in range(kz) for y in range(ky) for x in range(kx)]
    return torch.tensor(offs, dtype=torch.int32, device=device)


class _SparseConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None):
        super().__init__()
        assert groups == 1 and _triple(dilation) == (1, 1, 1)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _triple(kernel_size)          # (kz, ky, kx)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=np.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * int(np.prod(self.kernel_size))
            bound = 1 / np.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def _weight_kio(self) -> torch.Tensor:
        """[K, C_in, C_out] view of the spconv-layout parameter, cached on the parameter."""
        w = self.weight
        ver = (w._version, w.data_ptr())
        hit = w.__dict__.get('_lk_kio')
        if hit is not None and hit[0] == ver and not (torch.is_grad_enabled() and w.requires_grad):
            return hit[1]
        kio = w.permute(1, 2, 3, 4, 0).reshape(-1, self.in_channels, self.out_channels)
        if torch.is_grad_enabled() and w.requires_grad:
            return kio.contiguous()                      # keep the autograd graph to `weight`
        kio = kio.detach().contiguous()
        w.__dict__['_lk_kio'] = (ver, kio)
        return kio

    def _run(self, x: SparseConvTensor, kmap: KernelMap, out_indices, out_shape,
               scale=None, shift=None, relu=False, residual=None) -> SparseConvTensor:
        w = self._weight_kio()
        fused = not torch.is_grad_enabled()
        if fused:
            if self.bias is not None:
                b = self.bias.detach()
                shift = b if shift is None else shift + (b * scale if scale is not None else b)
            feats = _conv_fwd(x.features.contiguous().float(), w, kmap.nbr, kmap.n_out, None, scale,
                              shift, residual, relu, kmap=kmap, cache_on=self.weight)
        else:
            feats = ConvolutionFunction.apply(x.features, w, kmap, False)
            if self.bias is not None:
                feats = feats + self.bias
        out = SparseConvTensor(feats, out_indices, out_shape, x.batch_size, x.grid, x.voxel_num,
                               x.indice_dict, x.benchmark)
        out.benchmark_record = x.benchmark_record
        return out


class SubMConv3d(_SparseConvBase):
    """Submanifold conv: same active sites in and out; layers sharing an `indice_key` share the
    kernel map (scn.py:480-490)."""

    def kernel_map(self, x: SparseConvTensor) -> KernelMap:
        key = ('subm', self.indice_key if self.indice_key is not None else id(self), self.kernel_size)
        kmap = x.indice_dict.get(key)
        n = x.indices.shape[0]
        if kmap is None or kmap.n_out != n:
            coords = _xyzb(x.indices.int())
            centre = tuple(k // 2 for k in self.kernel_size)
            taps = _taps_xyz(self.kernel_size, centre, coords.device)
            table = x.indice_dict.get(('table', n, x.indices.data_ptr()))
            if table is None:
                table = HashTable(sphash(coords))
                x.indice_dict[('table', n, x.indices.data_ptr())] = table
            k = taps.shape[0]
            nbr = torch.empty(k, n, dtype=torch.int32, device=coords.device)
            L = _capi.lib()
            fn = L.lk_kmap_query_subm if k % 2 == 1 else L.lk_kmap_query
            _capi.check(fn(_capi.ptr(coords), n, _capi.ptr(taps), k, _capi.ptr(table.table),
                           table.capacity, _capi.ptr(nbr), _capi.stream()), 'lk_kmap_query')
            kmap = KernelMap(nbr, n, n, coords)
            kmap.offsets = taps                      # classes of the tile-skipping plan
            kmap.subm = all(k % 2 == 1 for k in self.kernel_size)   # symmetric taps: taps[K-1-k] == -taps[k]
            x.indice_dict[key] = kmap
        return kmap

    def forward(self, x: SparseConvTensor, **epilogue) -> SparseConvTensor:
        return self._run(x, self.kernel_map(x), x.indices, x.spatial_shape, **epilogue)


class SparseConv3d(_SparseConvBase):
    """Strided / padded sparse conv that creates new active sites (scn.py:494-566)."""

    def forward(self, x: SparseConvTensor, **epilogue) -> SparseConvTensor:
        dev = x.indices.device
        ks, st, pd = self.kernel_size, self.stride, self.padding
        in_shape = list(x.spatial_shape)
        out_shape = [(in_shape[a] + 2 * pd[a] - ks[a]) // st[a] + 1 for a in range(3)]
        idx = x.indices.int().contiguous()
        # candidate outputs o = (i + p - kappa) / s where divisible and in range: at most prod ceil(k/s)
        # per input site, written by one kernel (lk_strided_candidates), sorted + uniqued on the device;
        # the only host read-back is the (count, sentinel flag) pair that sizes the output
        import ctypes as C
        n = idx.shape[0]
        bs = int(x.batch_size)
        cap = 1
        for a in range(3):
            cap *= (ks[a] + st[a] - 1) // st[a]
        cand = torch.empty(n * cap, 4, dtype=torch.int32, device=dev)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        i3 = C.c_int32 * 3
        _capi.check(_capi.lib().lk_strided_candidates(_capi.ptr(idx), n, i3(*ks), i3(*st), i3(*pd), i3(*out_shape), bs,
                                                      _capi.ptr(cand), _capi.ptr(flag), _capi.stream()),
                    'lk_strided_candidates')
        spec, bits = _index.make_keyspec(((0, 0, 0, 0), (out_shape[2] - 1, out_shape[1] - 1, out_shape[0] - 1, bs)),
                                         (1, 1, 1), (3, 2, 1, 0))
        su = _index.sort_unique(_index.pack_keys(cand, spec), bits)
        m, any_invalid = torch.cat([su.num, flag]).tolist()
        uniq = _index.unpack_keys(su.unique, m - (1 if any_invalid else 0), spec)         # (x,y,z,b), sorted (b,z,y,x)
        out_indices = uniq[:, [3, 2, 1, 0]].contiguous()
        # kernel map: input site feeding output o through tap kappa is  o*s - p + kappa
        q = uniq.clone()
        q[:, 0] *= st[2]; q[:, 1] *= st[1]; q[:, 2] *= st[0]
        taps = _taps_xyz(ks, pd, dev)
        coords = _xyzb(idx)
        table = HashTable(sphash(coords))
        k, n_out = taps.shape[0], q.shape[0]
        nbr = torch.empty(k, n_out, dtype=torch.int32, device=dev)
        _capi.check(_capi.lib().lk_kmap_query(_capi.ptr(q), n_out, _capi.ptr(taps), k,
                                              _capi.ptr(table.table), table.capacity,
                                              _capi.ptr(nbr), _capi.stream()), 'lk_kmap_query')
        kmap = KernelMap(nbr, coords.shape[0], n_out, q)
        kmap.offsets = taps
        return self._run(x, kmap, out_indices, out_shape, **epilogue)


class SparseSequential(nn.Sequential):
    """spconv.SparseSequential: sparse modules take/return SparseConvTensor, dense modules
    (BatchNorm1d, ReLU) act on `.features`.  In inference a `conv -> BatchNorm1d(eval) [-> ReLU]`
    run collapses into the conv kernel's fused epilogue."""

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, _SparseConvBase) and not torch.is_grad_enabled():
                ep, j = {}, i + 1
                if j < len(mods) and isinstance(mods[j], nn.BatchNorm1d) and not mods[j].training:
                    ep['scale'], ep['shift'] = _folded_bn(mods[j])
                    j += 1
                if j < len(mods) and isinstance(mods[j], nn.ReLU):
                    ep['relu'] = True
                    j += 1
                x = m(x, **ep)
                i = j
                continue
            if isinstance(m, (_SparseConvBase, SparseSequential, SparseBasicBlock)):
                x = m(x)
            elif isinstance(m, nn.BatchNorm1d) and m.training:
                # BatchNorm1d(train) [-> ReLU] as one fused op (csrc/bn.cu)
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                x = x.replace_feature(batch_norm_act(x.features, m, relu))
                i += 1 if relu else 0
            else:
                x = x.replace_feature(m(x.features))
            i += 1
        return x


def _bn1d(planes, norm_cfg=None):
    cfg = dict(eps=1e-3, momentum=0.01)
    if norm_cfg:
        cfg.update({k: v for k, v in norm_cfg.items() if k in ('eps', 'momentum')})
    return nn.BatchNorm1d(planes, **cfg)


class SparseBasicBlock(nn.Module):
    """Two 3^3 submanifold convs + identity shortcut (scn.py:64-107); convs carry a bias because
    norm_cfg is always set there (scn.py:81-86)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, norm_cfg=None, downsample=None, indice_key=None):
        super().__init__()
        self.conv1 = SubMConv3d(inplanes, planes, 3, stride, padding=1, bias=True, indice_key=indice_key)
        self.bn1 = _bn1d(planes, norm_cfg)
        self.relu = nn.ReLU()
        self.conv2 = SubMConv3d(planes, planes, 3, padding=1, bias=True, indice_key=indice_key)
        self.bn2 = _bn1d(planes, norm_cfg)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        identity = x if self.downsample is None else self.downsample(x)
        if not torch.is_grad_enabled() and not self.bn1.training:
            s1, h1 = _folded_bn(self.bn1)
            s2, h2 = _folded_bn(self.bn2)
            out = self.conv1(x, scale=s1, shift=h1, relu=True)
            return self.conv2(out, scale=s2, shift=h2, relu=True,
                              residual=identity.features.contiguous().float())
        out = self.conv1(x)
        out = out.replace_feature(batch_norm_act(out.features, self.bn1, True))
        out = self.conv2(out)
        return out.replace_feature(batch_norm_act(out.features, self.bn2, True, identity.features))


class SpMiddleResNetFHDELKv3(nn.Module):
    """LinK detection backbone (scn.py:453-626): same constructor, sub-module names, forward
    signature `(voxel_features, coors(b,z,y,x), batch_size, input_shape)` and return value
    `(dense [B, C*D, H, W], {'conv1'..'conv4': SparseConvTensor})`."""

    def __init__(self, num_input_features=128, norm_cfg=None, name='SpMiddleResNetFHDELKv3', **kwargs):
        super().__init__()
        self.name = name
        self.dcn = None
        self.zero_init_residual = False
        p = self.planes = [16, 32, 64, 128]
        self.block_sz = 7

        def tail(c, key):
            return SparseSequential(SubMConv3d(c, c, 3, bias=False, indice_key=key), _bn1d(c, norm_cfg))

        self.conv_input = SparseSequential(
            SubMConv3d(num_input_features, p[0], 3, bias=False, indice_key='res0'),
            _bn1d(p[0], norm_cfg), nn.ReLU(inplace=True))
        for lv in (1, 2, 3, 4):
            c = p[lv - 1]
            if lv > 1:
                pad = 1 if lv < 4 else [0, 1, 1]
                setattr(self, f'down{lv}', SparseSequential(
                    SparseConv3d(p[lv - 2], c, 3, 2, padding=pad, bias=False),
                    _bn1d(c, norm_cfg), nn.ReLU(inplace=True)))
            setattr(self, f'conv{lv}', SparseSequential(
                SparseBasicBlock(c, c, norm_cfg=norm_cfg, indice_key=f'res{lv}'),
                SparseBasicBlock(c, c, norm_cfg=norm_cfg, indice_key=f'res{lv}')))
            setattr(self, f'conv{lv}_tail', tail(c, f'res{lv}_tail'))
            setattr(self, f'elk{lv}', TSELKBlock(c, c))
            setattr(self, f'elk{lv}_tail', tail(c, f'elk{lv}_tail'))
            setattr(self, f'act{lv}', nn.ReLU(inplace=True))
        self.extra_conv = SparseSequential(
            SparseConv3d(p[3], p[3], (3, 1, 1), (2, 1, 1), bias=False), _bn1d(p[3], norm_cfg), nn.ReLU())

    def forward(self, voxel_features, coors, batch_size, input_shape):
        sparse_shape = (np.array(input_shape[::-1]) + [1, 0, 0]).tolist()
        x = self.conv_input(SparseConvTensor(voxel_features, coors.int(), sparse_shape, batch_size))
        multi = {}
        for lv in (1, 2, 3, 4):
            if lv > 1:
                x = getattr(self, f'down{lv}')(x)
            x_conv = getattr(self, f'conv{lv}_tail')(getattr(self, f'conv{lv}')(x))
            x_lk = getattr(self, f'elk{lv}_tail')(getattr(self, f'elk{lv}')(x, self.block_sz))
            x = x_conv.replace_feature(getattr(self, f'act{lv}')(x_conv.features + x_lk.features))
            multi[f'conv{lv}'] = x
        ret = self.extra_conv(x).dense()
        n, c, d, h, w = ret.shape
        return ret.view(n, c * d, h, w), multi
