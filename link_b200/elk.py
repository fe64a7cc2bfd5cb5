"""The LinK block on B200: block index build, the reference's voxel_to_aux / aux_to_voxel API and
the fused ELKBlock.

Reference: ELKBlock (segmentation/core/models/semantic_kitti/linkencoder.py:94-185; UNet variant
linkunet.py:94-185), voxel_to_aux / aux_to_voxel / upsample_voxel / initial_voxelize
(segmentation/core/models/utils.py:44-84, 234-254, 327-340).

Two execution paths, both on liblinkb200 kernels (there is no PyTorch/CPU fallback):
  * fused (no autograd needed): block index -> lk_link_preagg_fwd -> lk_link_window_mean ->
    lk_link_apply_fwd with both LayerNorms, the add and the ReLU folded in.  No host sync.
  * composed (autograd): the reference's own op sequence on our differentiable
    spvoxelize / spdevoxelize kernels, used when gradients are required.
"""
import ctypes as C
import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from link_b200 import _capi
import link_b200.nn as spnn
import link_b200.nn.functional as F
from link_b200.nn.functional import _index
from link_b200.nn.functional import conv as _conv_mod
from link_b200.nn.utils import get_kernel_offsets
from link_b200.tensor import PointTensor, SparseTensor

__all__ = ['BlockIndex', 'block_index', 'voxel_to_aux', 'aux_to_voxel', 'upsample_voxel', 'upsample_index',
           'initial_voxelize', 'link_aggregate', 'elk_forward_fused', 'ELKBlock', 'LinKBlock']

_OPS = {'cos': 0, 'sin': 1, 'cos_x': 2}
# dense / sparse GEMMs on tcgen05 (3xTF32, fp32-level accuracy); False selects the FFMA kernels
# 1: libdevice sincosf (~1 ulp) in the kernel generator; 0 (default): two-term Cody-Waite reduction
# + SFU sin/cos (absolute error 2^-20.9), ~2x fewer instructions in the pre-aggregation kernel
ACCURATE_TRIG = os.environ.get('LINKB200_ACCURATE_TRIG', '0') == '1'
# 1 (default): one lk_elk_block_fwd call per block; 0: one python-level call per kernel
NATIVE_EXECUTOR = os.environ.get('LINKB200_NATIVE_EXECUTOR', '1') != '0'
USE_TENSOR_CORES = os.environ.get('LINKB200_TENSOR_CORES', '1') != '0'
# 1: keep the whole block on the caller's stream (default 0: two-chain schedule inside the executor)
SINGLE_STREAM = os.environ.get('LINKB200_SINGLE_STREAM', '0') == '1'
# 1 (default): training runs the fused forward + hand-written backward (cos / sin, C in {16,32,64,128});
# 0: the reference's op sequence on differentiable voxelize / devoxelize kernels
FUSED_BACKWARD = os.environ.get('LINKB200_FUSED_BACKWARD', '1') != '0'


class BlockIndex:
    """Device-resident index maps of one (coords, s) block partition (+ r^3 neighbour table).

    Everything is allocated for the worst case M = N and the true M stays on the device
    (`num`), so building and using the index needs no host synchronisation.  `.m` reads M back
    on demand (used only by the shape-returning reference API and by tests)."""

    def __init__(self, coords: torch.Tensor, s: int, cache: Optional[Dict]):
        self.n = coords.shape[0]
        self.s = int(s)
        bounds = _index.coord_bounds(coords, cache)
        self.spec, self.bits = _index.make_keyspec(bounds, (s, s, s), (0, 1, 2, 3))
        keys = _index.pack_keys(coords, self.spec)
        su = _index.sort_unique(keys, self.bits, want_order=True)
        self.order = su.order              # [n] int32  voxel row at each block-sorted position
        self.sorted_rank = su.sorted_rank  # [n] int32  block row at each block-sorted position
        self.seg = su.seg                  # [n+1] int32 start of each block's run in `order`
        self.unique_keys = su.unique       # [n] int64, first M valid, ascending == torch.unique order
        self.idx_query = su.inverse        # [n] int32  voxel -> block row
        self.counts = su.counts            # [n] int32, first M valid
        self.num = su.num                  # [1] int32 device scalar M
        self._m = None
        self._nbr = {}
        self._small_C = None

    @property
    def m(self) -> int:
        if self._m is None:
            self._m = int(self.num.item())
        return self._m

    @property
    def small_C(self) -> torch.Tensor:
        """[M,4] int32 block coordinates == torch.unique(x_C, dim=0) of the reference."""
        if self._small_C is None:
            self._small_C = _index.unpack_keys(self.unique_keys, self.m, self.spec)
        return self._small_C

    def neighbors(self, r: int) -> torch.Tensor:
        """[n, r^3] int32 (first M rows valid): row of each neighbour block in
        get_kernel_offsets(r,1,1) order, -1 if that block is empty (utils.py:65-73)."""
        nbr = self._nbr.get(r)
        if nbr is None:
            offsets = get_kernel_offsets(r, 1, 1, device=self.unique_keys.device)
            R = offsets.shape[0]
            nbr = torch.empty(self.n, R, dtype=torch.int32, device=self.unique_keys.device)
            _capi.check(_capi.lib().lk_block_neighbors(
                _capi.ptr(self.unique_keys), _capi.ptr(self.num), self.n, C.byref(self.spec),
                _capi.ptr(offsets), R, _capi.ptr(nbr), _capi.stream()), 'lk_block_neighbors')
            self._nbr[r] = nbr
        return nbr


    def neighbors_t(self, r: int) -> torch.Tensor:
        """Neighbour table of the TRANSPOSED relation {b : b' in N(b)} (used by the backward pass):
        the table of the negated offsets; for odd r the offset set is symmetric and this is
        `neighbors(r)` itself (up to the column order, which the window sum does not depend on)."""
        if r % 2 == 1:
            return self.neighbors(r)
        nbr = self._nbr.get(-r)
        if nbr is None:
            offsets = (-get_kernel_offsets(r, 1, 1, device=self.unique_keys.device)).contiguous()
            R = offsets.shape[0]
            nbr = torch.empty(self.n, R, dtype=torch.int32, device=self.unique_keys.device)
            _capi.check(_capi.lib().lk_block_neighbors(
                _capi.ptr(self.unique_keys), _capi.ptr(self.num), self.n, C.byref(self.spec),
                _capi.ptr(offsets), R, _capi.ptr(nbr), _capi.stream()), 'lk_block_neighbors')
            self._nbr[-r] = nbr
        return nbr


def block_index(st: SparseTensor, s: int) -> BlockIndex:
    """Block partition of `st` with block edge `s` (absolute voxel units), cached in the tensor
    family's shared `kmaps` (the reference recomputes it on every call)."""
    key = ('lk', 'blocks', st.stride, int(s), st.coords.data_ptr(), st.coords.shape[0])
    bi = st.kmaps.get(key)
    if bi is None:
        bi = BlockIndex(st.coords.contiguous(), s, st.kmaps)
        st.kmaps[key] = bi
    return bi


# ----------------------------------------------------------------------------- reference API
def voxel_to_aux(large_x: SparseTensor, s: int):
    """Block-local pre-aggregation with the reference signature and return values
    (utils.py:44-58): (aux SparseTensor with per-block MEAN features and stride s,
    idx_query [N] int64, counts [M] int32)."""
    bi = block_index(large_x, s)
    m = bi.m
    counts = bi.counts[:m]
    inserted = F.spvoxelize(large_x.F, bi.idx_query, counts)
    small_x = SparseTensor(inserted, bi.small_C, s)
    small_x.cmaps = large_x.cmaps
    small_x.kmaps = large_x.kmaps
    small_x._lk_block_index = bi
    return small_x, bi.idx_query.long(), counts


def aux_to_voxel(small_x: SparseTensor, large_x: SparseTensor, idx: torch.Tensor,
                 counts: torch.Tensor, r: int = 2) -> SparseTensor:
    """Outer-block reuse with the reference signature (utils.py:61-84): window mean over the r^3
    neighbour blocks, gathered back to the voxels.  Mutates and returns `large_x`."""
    bi = getattr(small_x, '_lk_block_index', None)
    m = small_x.F.shape[0]
    if bi is not None:
        nbr = bi.neighbors(r)[:m]
    else:   # foreign aux tensor: rebuild the neighbour table from its coordinates
        offsets = get_kernel_offsets(r, 1, 1, device=large_x.F.device)
        nbr = F.sphashquery(F.sphash(small_x.C.contiguous(), offsets),
                            F.sphash(small_x.C.contiguous())).t().contiguous().int()
    f = torch.cat([small_x.F, torch.ones_like(small_x.F[:, :1])], dim=1)
    f = f * counts.unsqueeze(dim=-1)
    weights = (nbr != -1).float()
    new_feat = F.spdevoxelize(f, nbr, weights, r)
    new_feat = new_feat[:, :-1] / new_feat[:, -1:]
    large_x.F = new_feat[idx.long()]
    return large_x


def upsample_index(x: SparseTensor, ref_x: SparseTensor) -> torch.Tensor:
    """int64 [N_ref]: row of `x` (coarse level) that is the parent of each voxel of `ref_x`
    (the index map of upsample_voxel, utils.py:329-335), cached in the family's kmaps."""
    stride = x.s[0]
    key = ('lk', 'upsample', x.s, ref_x.s, x.C.data_ptr(), ref_x.C.data_ptr())
    idx_query = x.kmaps.get(key)
    if idx_query is None:
        # hash(floor(C / stride)) of the coarse level -> table; probe with the fine level's
        # floor-divided coordinates (floor-division folded into the kernels)
        L, st = _capi.lib(), _capi.stream()
        xc, rc = x.C.contiguous(), ref_x.C.contiguous()
        h = torch.empty(xc.shape[0], dtype=torch.int64, device=xc.device)
        _capi.check(L.lk_hash_div(_capi.ptr(xc, torch.int32), xc.shape[0], int(stride), _capi.ptr(h), st),
                    'lk_hash_div')
        table = F.HashTable(h)
        idx_query = torch.empty(rc.shape[0], dtype=torch.int64, device=rc.device)
        _capi.check(L.lk_table_query_div(_capi.ptr(rc, torch.int32), rc.shape[0], int(stride),
                                         _capi.ptr(table.table), table.capacity,
                                         _capi.ptr(idx_query), st), 'lk_table_query_div')
        x.kmaps[key] = idx_query
    return idx_query


def upsample_voxel(x: SparseTensor, ref_x: SparseTensor) -> SparseTensor:
    """Nearest-parent gather from a coarse level to the fine level (utils.py:327-340)."""
    idx_query = upsample_index(x, ref_x)
    new_tensor = SparseTensor(x.F[idx_query], ref_x.C, ref_x.s)
    new_tensor.cmaps.setdefault(new_tensor.stride, new_tensor.coords)
    return new_tensor


def initial_voxelize(z: PointTensor, init_res, after_res) -> SparseTensor:
    """Points -> voxels (utils.py:234-254).  Voxel order = ascending FNV hash, exactly like the
    reference's `torch.unique(pc_hash)`; the unique runs on our radix sort (60-bit keys)."""
    new_float_coord = torch.cat([(z.C[:, :3] * init_res) / after_res, z.C[:, -1].view(-1, 1)], 1)
    fl = torch.floor(new_float_coord)
    pc_hash = F.sphash(fl.int())
    su = _index.sort_unique(pc_hash, 60)
    m = int(su.num.item())
    idx_query = su.inverse.long()
    counts = su.counts[:m].contiguous()
    inserted_coords = torch.round(F.spvoxelize(fl, su.inverse, counts)).int()
    inserted_feat = F.spvoxelize(z.F, su.inverse, counts)
    new_tensor = SparseTensor(inserted_feat, inserted_coords, 1)
    new_tensor.cmaps.setdefault(new_tensor.stride, new_tensor.coords)
    z.additional_features['idx_query'][1] = idx_query
    z.additional_features['counts'][1] = counts
    z.C = new_float_coord
    return new_tensor


# ----------------------------------------------------------------------------- fused path
def _kernel_gen(op: str, c: int, pos_weight: torch.Tensor, alpha: Optional[torch.Tensor],
                coord_scale: float) -> _capi.KernelGen:
    g = _capi.KernelGen()
    g.op = _OPS[op]
    g.c = c
    g.wrows = pos_weight.shape[0]
    g.coord_scale = float(coord_scale)
    g.d_pos_weight = _capi.ptr(pos_weight, torch.float32)
    g.d_alpha = _capi.ptr(alpha, torch.float32) if alpha is not None else None
    g.accurate_trig = 1 if ACCURATE_TRIG else 0
    return g


def link_aggregate(f_input: torch.Tensor, coords: torch.Tensor, bi: BlockIndex, r: int, op: str,
                   pos_weight: torch.Tensor, alpha: Optional[torch.Tensor] = None,
                   coord_scale: float = 1.0, local: Optional[torch.Tensor] = None,
                   norm: Optional[Tuple[torch.Tensor, ...]] = None,
                   save: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """Fused kernel generator + block pre-aggregation + outer-block reuse + combine.

    f_input [N,C] fp32 (output of pre_mix), coords int32 [N,4].  Returns the pre-LayerNorm value
    of linkencoder.py:162/148/176, or, when `local` [N,C] and `norm` = (g1, b1, g2, b2) are
    given, relu(LN(value) + LN(local)) of linkencoder.py:178-181.  Forward only (the differentiable
    form is `LinkAggregateFunction`); `save` = (mean [n,kC], tot [n]) buffers that receive the window
    means / populations the backward pass needs."""
    n, c = f_input.shape
    _capi.check_device(f_input)
    f_input = f_input.contiguous()
    coords = coords.contiguous()
    pos_weight = pos_weight.detach().contiguous().float()
    alpha_v = alpha.detach().reshape(-1).contiguous().float() if alpha is not None else None
    gen = _kernel_gen(op, c, pos_weight, alpha_v, coord_scale)
    k = 3 if op == 'cos_x' else 2
    L, st = _capi.lib(), _capi.stream()
    dev = f_input.device
    sums = torch.empty(n, k * c, dtype=torch.float32, device=dev)
    out = torch.empty(n, c, dtype=torch.float32, device=dev)
    nbr = bi.neighbors(r)
    m_hint = bi._m if bi._m is not None else 0      # only for the byte accounting of bench.py
    _capi.check(L.lk_zero_rows(_capi.ptr(sums), _capi.ptr(bi.num), n, k * c, st), 'lk_zero_rows')
    # algorithmic bytes: read F_in + coords + block index once, write the M block-sum rows
    # (the segmented kernel reads the 4-byte permutation entry in addition: not counted)
    with _capi.timed('lk_link_preagg_fwd', n * (4 * c + 16 + 4) + m_hint * 4 * k * c):
        _capi.check(L.lk_link_preagg_seg_fwd(_capi.ptr(f_input, torch.float32),
                                             _capi.ptr(coords, torch.int32), _capi.ptr(bi.order),
                                             _capi.ptr(bi.sorted_rank), n, C.byref(gen),
                                             _capi.ptr(sums), st), 'lk_link_preagg_seg_fwd')
    fuse = 1 if (local is not None and norm is not None) else 0
    g1 = b1 = g2 = b2 = None
    if fuse:
        local = local.contiguous()
        g1, b1, g2, b2 = (t.detach().contiguous().float() for t in norm)
    # read coords + block index (+ local, + F_in for cos_x) once, the M block-sum rows, write out
    nb = n * (16 + 4 + 4 * c * (1 + fuse + (1 if op == 'cos_x' else 0))) + m_hint * 4 * k * c
    mean = save[0] if save else torch.empty(n, k * c, dtype=torch.float32, device=dev)
    with _capi.timed('lk_link_window_mean', m_hint * (2 * 4 * k * c + 4 * nbr.shape[1] + 4)):
        _capi.check(L.lk_link_window_mean_tot(_capi.ptr(sums), _capi.ptr(bi.seg), _capi.ptr(nbr),
                                              _capi.ptr(bi.num), n, nbr.shape[1], k * c, _capi.ptr(mean),
                                              _capi.ptr(save[1]) if save else None, st), 'lk_link_window_mean_tot')
    with _capi.timed('lk_link_apply_fwd', nb):
        _capi.check(L.lk_link_apply_fwd(_capi.ptr(mean), _capi.ptr(f_input), _capi.ptr(coords),
                                        _capi.ptr(bi.idx_query), n, C.byref(gen), fuse,
                                        _capi.ptr(local) if fuse else None, _capi.ptr(g1),
                                        _capi.ptr(b1), _capi.ptr(g2), _capi.ptr(b2), _capi.ptr(out),
                                        st), 'lk_link_apply_fwd')
    return out


class LinkAggregateFunction(torch.autograd.Function):
    """Differentiable linear-kernel path with the fused norms:
    out = relu(LN(window_mean_combine(f_input)) + LN(local))  (linkencoder.py:150-162, 178-181),
    gradients for f_input, local, pos_weight and both LayerNorms by the hand-written backward
    kernels (lk_link_bwd_norm / lk_link_bwd_apply).  The reference reaches the
    same gradients through autograd over cat / spvoxelize / spdevoxelize / index
    (devoxelize.py:75-98, voxelize.py:33-56)."""

    @staticmethod
    def forward(ctx, f_input, local, pos_weight, g1, b1, g2, b2, coords, bi, r, op):
        n, c = f_input.shape
        k = 2
        dev = f_input.device
        f_input = f_input.contiguous().float()
        local = local.contiguous().float()
        mean = torch.empty(n, k * c, dtype=torch.float32, device=dev)
        tot = torch.empty(n, dtype=torch.float32, device=dev)
        out = link_aggregate(f_input, coords, bi, r, op, pos_weight, None, 1.0, local, (g1, b1, g2, b2),
                             save=(mean, tot))
        ctx.save_for_backward(f_input, local, pos_weight.detach().contiguous().float(), g1, b1, g2, b2, mean,
                              tot, coords)
        ctx.bi, ctx.r, ctx.op = bi, r, op
        return out

    @staticmethod
    def backward(ctx, dout):
        f_input, local, pw, g1, b1, g2, b2, mean, tot, coords = ctx.saved_tensors
        bi, r, op = ctx.bi, ctx.r, ctx.op
        n, c = f_input.shape
        dev = f_input.device
        dout = dout.contiguous().float()
        L, st = _capi.lib(), _capi.stream()
        gen = _kernel_gen(op, c, pw, None, 1.0)
        g1c, b1c, g2c, b2c = (t.detach().contiguous().float() for t in (g1, b1, g2, b2))
        dy = torch.empty_like(f_input)
        dlocal = torch.empty_like(f_input)
        gsum = torch.empty(n, 2 * c, dtype=torch.float32, device=dev)
        dparam = torch.zeros(4, c, dtype=torch.float32, device=dev)
        _capi.check(L.lk_link_bwd_norm(
            _capi.ptr(mean), _capi.ptr(tot), _capi.ptr(bi.seg), _capi.ptr(bi.order), _capi.ptr(bi.num), n,
            _capi.ptr(coords), C.byref(gen), _capi.ptr(local), _capi.ptr(dout), _capi.ptr(g1c), _capi.ptr(b1c),
            _capi.ptr(g2c), _capi.ptr(b2c), _capi.ptr(dy), _capi.ptr(dlocal), _capi.ptr(gsum), _capi.ptr(dparam),
            st), 'lk_link_bwd_norm')
        nbr_t = bi.neighbors_t(r)
        dfin = torch.empty_like(f_input)
        dw = torch.zeros_like(pw)
        _capi.check(L.lk_link_bwd_apply(
            _capi.ptr(gsum), _capi.ptr(mean), _capi.ptr(nbr_t), _capi.ptr(bi.seg), _capi.ptr(bi.order),
            _capi.ptr(bi.num), n, nbr_t.shape[1], _capi.ptr(coords), C.byref(gen), _capi.ptr(f_input),
            _capi.ptr(dy), _capi.ptr(dfin), _capi.ptr(dw), st), 'lk_link_bwd_apply')
        return dfin, dlocal, dw, dparam[0], dparam[1], dparam[2], dparam[3], None, None, None, None


def _pre_mix_fused(pre_mix, x: torch.Tensor) -> torch.Tensor:
    """pre_mix = Linear(no bias) + LayerNorm in one liblinkb200 kernel (forward only)."""
    lin, ln = pre_mix[0], pre_mix[1]
    x = x.contiguous()
    out = torch.empty_like(x)
    n, c = x.shape
    L = _capi.lib()
    fn = L.lk_linear_ln_tc_fwd if (c in (32, 64) and USE_TENSOR_CORES) else L.lk_linear_ln_fwd
    with _capi.timed('lk_linear_ln_fwd', n * 8 * c + 4 * c * c):
        _capi.check(fn(
            _capi.ptr(x, torch.float32), _capi.ptr(lin.weight.detach().contiguous()),
            _capi.ptr(ln.weight.detach().contiguous()), _capi.ptr(ln.bias.detach().contiguous()),
            float(ln.eps), n, c, _capi.ptr(out), _capi.stream()), 'lk_linear_ln_fwd')
    return out


def _native_template(op, c, pre_mix, conv, pos_weight, alpha, coord_scale, norm, norm_local, dev):
    """(version, ElkBlockArgs with every parameter-only field filled, tensors kept alive): the 12
    parameter pointers, the kernel-generator struct and the packed conv image are filled once per
    parameter version (cached on the conv module) and copied per call."""
    params = (pre_mix[0].weight, pre_mix[1].weight, pre_mix[1].bias, conv.kernel, pos_weight, alpha,
              norm.weight, norm.bias, norm_local.weight, norm_local.bias)
    ver = tuple((p._version, p.data_ptr()) if p is not None else None for p in params) + (
        op, c, float(coord_scale), USE_TENSOR_CORES, ACCURATE_TRIG, _conv_mod.precision_code(), SINGLE_STREAM, str(dev))
    hit = conv.__dict__.get('_lk_native_args')
    if hit is None or hit[0] != ver:
        from link_b200.nn.functional.conv import _tc_image
        t = _capi.ElkBlockArgs()
        lin, ln = pre_mix[0], pre_mix[1]
        keep = [lin.weight.detach().contiguous(), ln.weight.detach(), ln.bias.detach(), conv.kernel.detach().contiguous(),
                pos_weight.detach().contiguous().float(),
                alpha.detach().reshape(-1).contiguous().float() if alpha is not None else None,
                norm.weight.detach(), norm.bias.detach(), norm_local.weight.detach(), norm_local.bias.detach()]
        t.d_premix_w, t.d_premix_g, t.d_premix_b = _capi.ptr(keep[0]), _capi.ptr(keep[1]), _capi.ptr(keep[2])
        t.premix_eps = float(ln.eps)
        t.kvol = conv.kernel_volume
        t.d_conv_w = _capi.ptr(keep[3])
        wt = _tc_image(conv.kernel) if (USE_TENSOR_CORES and c in (32, 64, 128)) else None   # cached on the Parameter
        keep.append(wt)
        t.d_conv_wt = _capi.ptr(wt)
        t.gen = _kernel_gen(op, c, keep[4], keep[5], coord_scale)
        t.d_g1, t.d_b1, t.d_g2, t.d_b2 = (_capi.ptr(x) for x in keep[6:10])
        t.use_tensor_cores = 1 if USE_TENSOR_CORES else 0
        t.conv_precision = _conv_mod.precision_code()
        t.single_stream = 1 if SINGLE_STREAM else 0
        hit = conv.__dict__['_lk_native_args'] = (ver, t, keep)
    return hit


def _forward_native(st: SparseTensor, s, r, *, op, pre_mix, conv, pos_weight, alpha, coord_scale, norm,
                    norm_local) -> torch.Tensor:
    """Whole block through lk_elk_block_fwd: one FFI call, one workspace allocation."""
    from link_b200.nn.functional.conv import KernelMap, _tc_image
    L = _capi.lib()
    _capi.check_device(st._feats)
    x, ready = st.take_feats_event()          # a pending async upload is joined ON THE DEVICE, after
    x = x.contiguous()                         # the index-only kernels (SparseTensor.from_host)
    coords = st.C.contiguous()
    n, c = x.shape
    dev = x.device
    key = (st.stride, conv.kernel_size, conv.stride, (1, 1, 1))
    kmap = st.kmaps.get(key)
    build = kmap is None
    if build:
        kmap = KernelMap(torch.empty(conv.kernel_volume, n, dtype=torch.int32, device=dev), n, n, coords)
        kmap.subm = conv.kernel_volume % 2 == 1 and all(s == 1 for s in conv.stride)
        st.kmaps[key] = kmap
    conv_off = get_kernel_offsets(conv.kernel_size, stride=st.stride, device=dev)
    blk_off = get_kernel_offsets(r, 1, 1, device=dev)
    r3 = blk_off.shape[0]
    hit = _native_template(op, c, pre_mix, conv, pos_weight, alpha, coord_scale, norm, norm_local, dev)
    a = _capi.ElkBlockArgs.from_buffer_copy(hit[1])
    a.n = n
    a.d_coords, a.d_feats = _capi.ptr(coords, torch.int32), _capi.ptr(x, torch.float32)
    out = torch.empty_like(x)
    a.d_out = _capi.ptr(out)
    a.d_conv_offsets = _capi.ptr(conv_off)
    a.d_kmap, a.build_kmap = _capi.ptr(kmap.nbr), 1 if build else 0
    if build:
        kmap.offsets = conv_off
    a.build_plan = 0
    if USE_TENSOR_CORES and c in (32, 64, 128) and conv.kernel_volume <= 32 and kmap._plan is not False:
        if kmap._plan is None and _conv_mod.USE_PLAN:     # the executor fills the plan buffers
            kmap._plan = kmap.plan_buffers()
            a.build_plan = 1
        if kmap._plan:
            a.d_plan_perm, a.d_plan_mask = (_capi.ptr(t_) for t_ in kmap._plan)
    bounds = _index.coord_bounds(coords, st.kmaps)
    spec, bits = _index.make_keyspec(bounds, (s, s, s), (0, 1, 2, 3))
    a.keyspec, a.key_bits, a.r3 = spec, bits, r3
    a.d_block_offsets = _capi.ptr(blk_off)
    ws_bytes = L.lk_elk_block_ws_bytes(n, c, a.gen.op, r3, a.kvol, a.build_kmap)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    a.d_ws, a.ws_bytes = _capi.ptr(ws), ws_bytes
    a.feats_ready = ready.cuda_event if ready is not None else None
    _capi.check(L.lk_elk_block_fwd(C.byref(a), _capi.stream()), 'lk_elk_block_fwd')
    return out


def elk_forward_fused(st: SparseTensor, s, r, *, op, pre_mix, conv, pos_weight, alpha, coord_scale,
                      norm, norm_local) -> torch.Tensor:
    """Fused forward of a LinK block given its sub-modules (shared by ELKBlock and TSELKBlock):
    the native executor when available, else one python-level call per kernel."""
    c = st._feats.shape[1]
    # native executor for the channel counts its tensor-core conv serves directly; narrower blocks
    # (C = 16) take the per-kernel path, whose conv runs zero-padded on the tensor cores (the FFMA
    # conv the executor would fall back to is ~5x slower)
    if NATIVE_EXECUTOR and _capi.TIMERS is None and c in (32, 64, 128):
        return _forward_native(st, s, r, op=op, pre_mix=pre_mix, conv=conv, pos_weight=pos_weight,
                               alpha=alpha, coord_scale=coord_scale, norm=norm, norm_local=norm_local)
    if c in (16, 32, 64, 128):
        f_input = _pre_mix_fused(pre_mix, st.F)
    else:
        f_input = pre_mix(st.F)
    local = F.conv3d(st, conv.kernel, kernel_size=conv.kernel_size, bias=conv.bias,
                     stride=conv.stride, dilation=conv.dilation)
    bi = block_index(st, s)
    return link_aggregate(f_input, st.C, bi, r, op, pos_weight, alpha, coord_scale, local.F,
                          (norm.weight, norm.bias, norm_local.weight, norm_local.bias))


class ELKBlock(nn.Module):
    """LinK block.  Constructor, parameters (names and shapes) and call signature are the
    reference's (linkencoder.py:94-185): `ELKBlock(inc, outc, groups, baseop)(st, s, r)`.

    `variant='unet'` reproduces linkunet.py:165 (cos_x phase not divided by the tensor stride)."""

    def __init__(self, inc, outc, groups=1, baseop='cos_x', variant='encoder'):
        super().__init__()
        self.inc, self.outc, self.groups, self.baseop = inc, outc, groups, baseop
        self.variant = variant
        assert inc % groups == 0
        assert baseop in ['cos', 'sin', 'cos_x']
        if baseop == 'cos_x':
            self.alpha = nn.Parameter(torch.ones(1, inc // groups).float(), requires_grad=True)
        self.pos_weight = nn.Sequential(nn.Linear(3, inc // groups, bias=False))
        self.pre_mix = nn.Sequential(nn.Linear(inc, inc, bias=False), nn.LayerNorm(inc, eps=1e-6))
        self.local_mix = nn.Sequential(spnn.Conv3d(inc, inc, kernel_size=3, dilation=1, stride=1))
        self.norm_local = nn.LayerNorm(inc, eps=1e-6)
        self.norm = nn.LayerNorm(inc, eps=1e-6)
        self.activate = nn.ReLU(True)

    def _needs_grad(self, st: SparseTensor) -> bool:
        return torch.is_grad_enabled() and (st._feats.requires_grad or
                                            any(p.requires_grad for p in self.parameters()))

    def forward(self, st: SparseTensor, s, r):
        needs_grad = self._needs_grad(st)
        if not needs_grad and st._feats.dtype in (torch.bfloat16, torch.float16):
            # reduced-precision activations at inference: the fused block computes in fp32 (its kernels read
            # fp32 rows) and hands the activation dtype back -- two casts instead of the composed op sequence
            dt = st._feats.dtype
            st.F = st.F.float()
            return self._cast_back(self.forward(st, s, r), dt)
        composed = needs_grad or st._feats.dtype != torch.float32   # (no upload join here)
        if not composed:
            if self.baseop == 'cos_x' and self.groups != 1:
                raise RuntimeError("baseop='cos_x' needs groups == 1 (the reference's phase tensor "
                                   "is not repeated over groups, linkencoder.py:165)")
            scale = float(st.s[0]) if (self.baseop == 'cos_x' and self.variant == 'encoder') else 1.0
            st.F = elk_forward_fused(st, s, r, op=self.baseop, pre_mix=self.pre_mix,
                                     conv=self.local_mix[0], pos_weight=self.pos_weight[0].weight,
                                     alpha=getattr(self, 'alpha', None), coord_scale=scale,
                                     norm=self.norm, norm_local=self.norm_local)
            return st          # like aux_to_voxel, the input tensor object carries the result
        F_input, local_mix = self.pre_mix(st.F), self.local_mix(st)
        c = self.inc
        if (FUSED_BACKWARD and self.baseop in ('cos', 'sin') and st._feats.dtype == torch.float32
                and _capi.lib().lk_link_bwd_supported(c)):
            # training: fused forward + hand-written backward of the linear-kernel path; pre_mix and
            # local_mix stay autograd nodes of their own (dense Linear+LayerNorm, sparse conv)
            st.F = LinkAggregateFunction.apply(
                F_input, local_mix.F, self.pos_weight[0].weight, self.norm.weight, self.norm.bias,
                self.norm_local.weight, self.norm_local.bias, st.C.contiguous(), block_index(st, s), r,
                self.baseop)
            return st
        return self._forward_composed(st, F_input, local_mix, s, r)

    @staticmethod
    def _cast_back(st, dtype):
        st.F = st.F.to(dtype)
        return st

    def _forward_composed(self, st, F_input, local_mix, s, r):
        """The reference's op sequence (linkencoder.py:135-183) on differentiable kernels."""
        C_ = self.inc
        xyz = st.C[:, :3].float()
        if self.baseop == 'cos_x':
            if self.variant == 'encoder':
                xyz = xyz / st.s[0]
            pos = self.pos_weight(xyz) * self.alpha
        else:
            pos = self.pos_weight(xyz).repeat([1, self.groups])
        sin, cos = torch.sin(pos), torch.cos(pos)
        if self.baseop == 'sin':
            planes = [F_input * sin, F_input * cos]
        elif self.baseop == 'cos':
            planes = [F_input * cos, F_input * sin]
        else:
            lin = F_input * pos
            planes = [F_input * cos, F_input * sin, lin]
        st.F = torch.cat(planes, dim=1).contiguous()
        aux_st, idx, counts = voxel_to_aux(st, s)
        voxel_st = aux_to_voxel(aux_st, st, idx, counts, r)
        vf = voxel_st.F
        if self.baseop == 'sin':
            new = vf[:, :C_] * cos - vf[:, C_:] * sin
        elif self.baseop == 'cos':
            new = vf[:, :C_] * cos + vf[:, C_:] * sin
        else:
            new = vf[:, :C_] * cos + vf[:, C_:2 * C_] * sin + (vf[:, 2 * C_:] - lin)
        new = self.activate(self.norm(new) + self.norm_local(local_mix.F))
        voxel_st.F = new
        return voxel_st


LinKBlock = ELKBlock
