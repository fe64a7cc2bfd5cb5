"""CPU oracle for the LinK hot path -- TEST INFRASTRUCTURE ONLY.

A plain numpy / torch-CPU restatement of the reference algorithm
(MCG-NJU/LinK @ f939adc).  Integer / index work is done in numpy (bit-exact by
construction), floating-point work in torch CPU fp32 with differentiable ops so
that torch autograd of this file is the gradient oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (link_b200/) never does.

Parity pin: the reference has no tests / golden vectors, so this
oracle is pinned against outputs of the reference itself, executed in the build
container (oracle/_ref/backend.so = the reference's CPU backend compiled from
/root/reference + the reference's unmodified python layer).  See
tests/golden/make_golden.py and tests/test_oracle_vs_golden.py.  Known envelope
of the reference CPU path: r=2 and batch==1 only (devoxelize_cpu.cpp:19-24
hard-codes 8 neighbours, hash_cpu.cpp:29 reads row 0's batch index); outside it
this file follows the reference's CUDA sources, which are the correct ones.

Paths below are relative to /root/reference;
ts/ = segmentation/torchsparse-u/torchsparse/, seg/ = segmentation/.
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as TF

FNV_OFFSET = np.uint64(14695981039346656037)
FNV_PRIME = np.uint64(1099511628211)
MASK60 = np.uint64(0x0FFFFFFFFFFFFFFF)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def make_ntuple(x, ndim=3):
    # ts/utils/utils.py:9-21
    if isinstance(x, int):
        return (x,) * ndim
    if isinstance(x, torch.Tensor):
        x = x.view(-1).tolist()
    x = tuple(int(v) for v in x)
    assert len(x) == ndim, x
    return x


# --------------------------------------------------------------------------
# integer ops
# --------------------------------------------------------------------------
def _fnv(c: np.ndarray) -> np.ndarray:
    """c: [..., 4] int32 -> [...] int64.  ts/backend/hash/hash_cuda.cu:14-21."""
    c = c.astype(np.int32, copy=False).view(np.uint32).astype(np.uint64)
    h = np.full(c.shape[:-1], FNV_OFFSET, dtype=np.uint64)
    with np.errstate(over='ignore'):
        for j in range(4):
            h = h ^ c[..., j]
            h = h * FNV_PRIME
    h = (h >> np.uint64(60)) ^ (h & MASK60)
    return h.astype(np.int64)


def sphash(coords, offsets=None) -> np.ndarray:
    """ts/nn/functional/hash.py:10-37; hash_cuda.cu:10-23 (no offsets) and
    hash_cuda.cu:27-55 (offsets -> layout [K, N], batch column kept)."""
    coords = _np(coords)
    assert coords.dtype == np.int32 and coords.ndim == 2 and coords.shape[1] == 4
    if offsets is None:
        return _fnv(coords)
    offsets = _np(offsets)
    assert offsets.dtype == np.int32 and offsets.ndim == 2 and offsets.shape[1] == 3
    K = offsets.shape[0]
    cur = np.empty((K, coords.shape[0], 4), dtype=np.int32)
    with np.errstate(over='ignore'):
        cur[:, :, :3] = coords[None, :, :3] + offsets[:, None, :]
    cur[:, :, 3] = coords[None, :, 3]
    return _fnv(cur)


def sphashquery(queries, references) -> np.ndarray:
    """ts/nn/functional/query.py:8-33: index of `query` in `references` or -1.
    Duplicate reference hashes: the first one inserted wins (query_cpu.cpp:22-26,
    insert() does not overwrite)."""
    q = _np(queries).astype(np.int64)
    ref = _np(references).astype(np.int64)
    shape = q.shape
    q = q.reshape(-1)
    if ref.size == 0:
        return np.full(shape, -1, dtype=np.int64)
    uniq, first = np.unique(ref, return_index=True)
    pos = np.searchsorted(uniq, q)
    pos_c = np.minimum(pos, len(uniq) - 1)
    hit = uniq[pos_c] == q
    out = np.where(hit, first[pos_c], -1).astype(np.int64)
    return out.reshape(shape)


def spcount(idx, num: int) -> np.ndarray:
    """ts/backend/others/count_cuda.cu:10-16 (negative indices skipped)."""
    idx = _np(idx).astype(np.int64)
    return np.bincount(idx[idx >= 0], minlength=int(num)).astype(np.int32)[:int(num)]


def get_kernel_offsets(size, stride=1, dilation=1) -> np.ndarray:
    """ts/nn/utils/kernel.py:11-32.  Odd volume: z-major (x fastest);
    even volume: x-major (z fastest)."""
    size, stride, dilation = make_ntuple(size), make_ntuple(stride), make_ntuple(dilation)
    offs = [np.arange(-size[k] // 2 + 1, size[k] // 2 + 1) * stride[k] * dilation[k]
            for k in range(3)]
    if int(np.prod(size)) % 2 == 1:
        out = [[x, y, z] for z in offs[2] for y in offs[1] for x in offs[0]]
    else:
        out = [[x, y, z] for x in offs[0] for y in offs[1] for z in offs[2]]
    return np.asarray(out, dtype=np.int32).reshape(-1, 3)


def unique_rows(c: np.ndarray) -> np.ndarray:
    """torch.unique(x, dim=0): rows in ascending signed lexicographic order."""
    return np.unique(c, axis=0)


def spdownsample(coords, stride=2, kernel_size=2, tensor_stride=1) -> np.ndarray:
    """ts/nn/functional/downsample.py:11-51."""
    coords = _np(coords).astype(np.int32)
    stride, kernel_size, tensor_stride = (make_ntuple(stride), make_ntuple(kernel_size),
                                          make_ntuple(tensor_stride))
    ss = np.asarray([stride[k] * tensor_stride[k] for k in range(3)], dtype=np.int32)[None]
    if all(stride[k] in [1, kernel_size[k]] for k in range(3)):
        coords = coords.copy()
        coords[:, :3] = np.floor_divide(coords[:, :3], ss) * ss
    else:
        offsets = get_kernel_offsets(kernel_size, tensor_stride)
        K = offsets.shape[0]
        cmin = coords[:, :3].min(axis=0, keepdims=True)
        x = (coords[:, None, :3] + offsets[None]).reshape(-1, 3)
        b = np.repeat(coords[:, 3:], K, axis=1).reshape(-1, 1)
        coords = np.concatenate([x, b], axis=1)
        mask = (np.mod(coords[:, :3], ss) == 0) & (coords[:, :3] >= cmin)
        coords = coords[mask.all(axis=1)]
    coords = unique_rows(coords[:, [3, 0, 1, 2]])
    return np.ascontiguousarray(coords[:, [1, 2, 3, 0]])


def build_kmap(coords, tensor_stride, kernel_size, stride=1, dilation=1):
    """ts/nn/functional/conv.py:103-122.  Returns (nbmaps [P,2] int64 with
    column 0 = input row, column 1 = output row, ordered by (k, out_row);
    nbsizes [K] int64; out_coords; results [K, Nout])."""
    coords = _np(coords).astype(np.int32)
    kernel_size, stride, tensor_stride = (make_ntuple(kernel_size), make_ntuple(stride),
                                          make_ntuple(tensor_stride))
    offsets = get_kernel_offsets(kernel_size, stride=tensor_stride)
    references = sphash(coords)
    out_coords = coords
    if any(s > 1 for s in stride):
        out_coords = spdownsample(coords, stride, kernel_size, tensor_stride)
    queries = sphash(out_coords, offsets)
    results = sphashquery(queries, references)
    nbsizes = (results != -1).sum(axis=1).astype(np.int64)
    k_idx, o_idx = np.nonzero(results != -1)
    nbmaps = np.stack([results[k_idx, o_idx], o_idx], axis=1).astype(np.int64)
    return nbmaps, nbsizes, out_coords, results


# --------------------------------------------------------------------------
# float ops (torch CPU fp32, differentiable)
# --------------------------------------------------------------------------
def _t(x, dtype=None):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    return t if dtype is None else t.to(dtype)


def spvoxelize(feats: torch.Tensor, idx, counts) -> torch.Tensor:
    """ts/backend/voxelize/voxelize_cuda.cu:12-25: out[idx[i]] += feats[i] / counts[idx[i]]."""
    idx = _t(idx, torch.long)
    counts = _t(counts)
    M = counts.shape[0]
    valid = (idx >= 0) & (idx < M)
    safe = idx.clamp(0, max(M - 1, 0))
    cnt = counts[safe].to(feats.dtype)
    valid = valid & (cnt > 0)
    contrib = torch.where(valid[:, None], feats / cnt.clamp(min=1)[:, None],
                          torch.zeros((), dtype=feats.dtype))
    out = torch.zeros(M, feats.shape[1], dtype=feats.dtype)
    return out.index_add(0, safe, contrib)


def spdevoxelize(feats: torch.Tensor, idx, weights: torch.Tensor) -> torch.Tensor:
    """ts/backend/devoxelize/devoxelize_cuda.cu:11-34 with R = idx.shape[1] (= r^3):
    out[i] = sum_k w[i,k] * feats[idx[i,k]] (idx<0 contributes 0), k ascending."""
    idx = _t(idx, torch.long)
    N, R = idx.shape
    out = torch.zeros(N, feats.shape[1], dtype=feats.dtype)
    for k in range(R):
        col = idx[:, k]
        g = feats[col.clamp(min=0)]
        g = torch.where((col >= 0)[:, None], g, torch.zeros((), dtype=feats.dtype))
        out = out + weights[:, k:k + 1] * g
    return out


def conv_apply(feats: torch.Tensor, weight: torch.Tensor, nbmaps, nbsizes,
               sizes: Tuple[int, int], transposed: bool = False) -> torch.Tensor:
    """Arithmetic of ConvolutionFunction: ts/nn/functional/conv.py:47-61
    (gather -> mm -> index-add per kernel offset); convolution_cuda.cu:101-164."""
    nbmaps = _t(nbmaps, torch.long)
    nbsizes = [int(v) for v in _np(nbsizes)]
    n_out = sizes[1] if not transposed else sizes[0]
    out = torch.zeros(n_out, weight.shape[-1], dtype=feats.dtype)
    cur = 0
    for k in range(weight.shape[0]):
        n = nbsizes[k]
        in_map, out_map = nbmaps[cur:cur + n, 0], nbmaps[cur:cur + n, 1]
        cur += n
        if n == 0:
            continue
        if transposed:
            in_map, out_map = out_map, in_map
        out = out.index_add(0, out_map, feats[in_map] @ weight[k])
    return out


class OTensor:
    """Minimal stand-in for ts/tensor.py:10-70 (feats, coords, stride, shared caches)."""

    def __init__(self, feats, coords, stride=1):
        self.F = feats
        self.C = _np(coords).astype(np.int32)
        self.s = make_ntuple(stride)
        self.cmaps: Dict = {}
        self.kmaps: Dict = {}


def conv3d(x: OTensor, weight: torch.Tensor, kernel_size, bias=None, stride=1,
           dilation=1, transposed=False) -> OTensor:
    """ts/nn/functional/conv.py:83-147."""
    kernel_size, stride, dilation = make_ntuple(kernel_size), make_ntuple(stride), make_ntuple(dilation)
    feats, coords = x.F, x.C
    if kernel_size == (1, 1, 1) and stride == (1, 1, 1) and dilation == (1, 1, 1):
        feats = feats @ weight
        if bias is not None:
            feats = feats + bias
        out = OTensor(feats, coords, x.s)
    elif not transposed:
        key = (x.s, kernel_size, stride, dilation)
        kmap = x.kmaps.get(key)
        if kmap is None:
            nbmaps, nbsizes, oc, _ = build_kmap(coords, x.s, kernel_size, stride, dilation)
            kmap = [nbmaps, nbsizes, (feats.shape[0], oc.shape[0]), oc]
            x.kmaps[key] = kmap
        feats = conv_apply(feats, weight, kmap[0], kmap[1], kmap[2], False)
        if bias is not None:
            feats = feats + bias
        out = OTensor(feats, kmap[3], tuple(x.s[k] * stride[k] for k in range(3)))
    else:
        ts = tuple(x.s[k] // stride[k] for k in range(3))
        kmap = x.kmaps[(ts, kernel_size, stride, dilation)]
        feats = conv_apply(feats, weight, kmap[0], kmap[1], kmap[2], True)
        if bias is not None:
            feats = feats + bias
        out = OTensor(feats, x.cmaps[ts], ts)
    out.cmaps = x.cmaps
    out.cmaps.setdefault(out.s, out.C)
    out.kmaps = x.kmaps
    return out


# --------------------------------------------------------------------------
# LinK block: voxel_to_aux / aux_to_voxel / ELKBlock
# --------------------------------------------------------------------------
def block_index(coords, s: int):
    """Integer part of voxel_to_aux, seg/core/models/utils.py:44-51
    (== large_to_small, detection/det3d/models/utils/ts_elk.py:68-75).
    Returns (small_C [M,4] int32 in torch.unique(dim=0) order, idx_query [N] int64,
    counts [M] int32)."""
    coords = _np(coords).astype(np.int32)
    x_C = np.concatenate([np.floor_divide(coords[:, :3], np.int32(s)).astype(np.int32),
                          coords[:, 3:]], axis=1)
    large_hash = sphash(x_C)
    small_C = unique_rows(x_C)
    small_hash = sphash(small_C)
    idx_query = sphashquery(large_hash, small_hash)
    counts = spcount(idx_query.astype(np.int32), len(small_hash))
    return small_C, idx_query, counts


def block_neighbors(small_C, r: int) -> np.ndarray:
    """Integer part of aux_to_voxel, seg/core/models/utils.py:65-73: [M, r^3] int64,
    -1 where the neighbour block is empty; column order = get_kernel_offsets(r,1,1)."""
    offsets = get_kernel_offsets(r, 1, 1)
    nh = sphash(small_C, offsets)
    sh = sphash(small_C)
    return np.ascontiguousarray(sphashquery(nh, sh).T)


def voxel_to_aux(feats: torch.Tensor, coords, s: int):
    """seg/core/models/utils.py:44-58."""
    small_C, idx_query, counts = block_index(coords, s)
    aux_F = spvoxelize(feats, idx_query, counts)
    return aux_F, small_C, idx_query, counts


def aux_to_voxel(aux_F: torch.Tensor, small_C, idx, counts, r: int = 2) -> torch.Tensor:
    """seg/core/models/utils.py:61-84.  Returns the new voxel features [N, c]."""
    nbr = block_neighbors(small_C, r)
    cnt = _t(counts).to(aux_F.dtype)
    f = torch.cat([aux_F, torch.ones_like(aux_F[:, :1])], dim=1) * cnt[:, None]
    w = torch.ones(aux_F.shape[0], r ** 3, dtype=torch.float32)
    w[_t(nbr) == -1] = 0
    new = spdevoxelize(f, nbr, w)
    new = new[:, :-1] / new[:, -1:]
    return new[_t(idx, torch.long)]


def window_mean_bruteforce(feats: torch.Tensor, coords, s: int, r: int) -> torch.Tensor:
    """Semantic definition used as an oracle-of-the-oracle (small N only): the mean
    of feats over every voxel of the same batch whose s-block lies in the r^3 offset
    set around the voxel's own block."""
    coords = _np(coords).astype(np.int64)
    blk = np.concatenate([np.floor_divide(coords[:, :3], s), coords[:, 3:]], axis=1)
    offs = get_kernel_offsets(r, 1, 1).astype(np.int64)
    out = torch.zeros_like(feats)
    for i in range(coords.shape[0]):
        d = blk[:, :3] - blk[i, :3]
        m = (blk[:, 3] == blk[i, 3]) & (d[:, None, :] == offs[None]).all(-1).any(-1)
        out[i] = feats[torch.from_numpy(m)].double().mean(0).to(feats.dtype)
    return out


def elk_pos(coords, tensor_stride, p: Dict[str, torch.Tensor], baseop: str, groups: int,
            variant: str = 'encoder') -> torch.Tensor:
    """Kernel-generator phase [N, C].
    encoder: seg/core/models/semantic_kitti/linkencoder.py:136-137,151-152,165;
    unet:    seg/core/models/semantic_kitti/linkunet.py:165 (no stride division);
    det:     detection/det3d/models/utils/ts_elk.py:154,167-168 (Linear(3,C), first
             half repeated twice for 'cos', full width for 'sin')."""
    xyz = _t(_np(coords)[:, :3].astype(np.float32))
    W = p['pos_weight.0.weight']
    if variant == 'det':
        pos = TF.linear(xyz, W)
        if baseop == 'cos':
            pos = pos[:, :W.shape[0] // 2].repeat(1, 2)
        return pos
    if baseop in ('sin', 'cos'):
        return TF.linear(xyz, W).repeat(1, groups)
    if variant == 'encoder':
        xyz = xyz / tensor_stride[0]
    return TF.linear(xyz, W) * p['alpha']


def elk_block_forward(feats: torch.Tensor, coords, tensor_stride, p: Dict[str, torch.Tensor],
                      s: int, r: int, baseop: str = 'cos', groups: int = 1,
                      variant: str = 'encoder', kmaps: Optional[Dict] = None,
                      return_parts: bool = False):
    """ELKBlock.forward, seg/core/models/semantic_kitti/linkencoder.py:124-185
    (TSELKBlock.forward_, ts_elk.py:144-230 with variant='det', r=3)."""
    C = feats.shape[1]
    tensor_stride = make_ntuple(tensor_stride)
    F_input = TF.layer_norm(TF.linear(feats, p['pre_mix.0.weight']), (C,),
                            p['pre_mix.1.weight'], p['pre_mix.1.bias'], 1e-6)
    st = OTensor(feats, coords, tensor_stride)
    if kmaps is not None:
        st.kmaps = kmaps
    local = conv3d(st, p['local_mix.0.kernel'], 3).F
    pos = elk_pos(coords, tensor_stride, p, baseop, groups, variant)
    sin, cos = torch.sin(pos), torch.cos(pos)
    if baseop == 'sin':
        planes = torch.cat([F_input * sin, F_input * cos], dim=1)
    elif baseop == 'cos':
        planes = torch.cat([F_input * cos, F_input * sin], dim=1)
    elif baseop == 'cos_x':
        lin = F_input * pos
        planes = torch.cat([F_input * cos, F_input * sin, lin], dim=1)
    else:
        raise ValueError(baseop)
    aux_F, small_C, idx, counts = voxel_to_aux(planes.contiguous(), coords, s)
    vf = aux_to_voxel(aux_F, small_C, idx, counts, r)
    if baseop == 'sin':
        new = vf[:, :C] * cos - vf[:, C:] * sin
    elif baseop == 'cos':
        new = vf[:, :C] * cos + vf[:, C:] * sin
    else:
        new = vf[:, :C] * cos + vf[:, C:2 * C] * sin + (vf[:, 2 * C:] - lin)
    pre_norm = new
    new = TF.layer_norm(new, (C,), p['norm.weight'], p['norm.bias'], 1e-6)
    loc = TF.layer_norm(local, (C,), p['norm_local.weight'], p['norm_local.bias'], 1e-6)
    out = torch.relu(new + loc)
    if return_parts:
        return out, dict(F_input=F_input, local=local, pos=pos, pre_norm=pre_norm, pre_act=new + loc,
                         small_C=small_C, idx=idx, counts=counts)
    return out


# --------------------------------------------------------------------------
# encoder (linkencoder.py:188-381) driven by a state dict
# --------------------------------------------------------------------------
def _sub(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _bn(x: torch.Tensor, p: Dict[str, torch.Tensor], training: bool) -> torch.Tensor:
    # spnn.BatchNorm == nn.BatchNorm1d on feats (ts/nn/modules/norm.py:10-13)
    return TF.batch_norm(x, None if training else p['running_mean'],
                         None if training else p['running_var'],
                         p['weight'], p['bias'], training, 0.1, 1e-5)


def _conv_bn(x: OTensor, sd, prefix, ks, stride=1, relu=False, training=False) -> OTensor:
    y = conv3d(x, sd[prefix + '0.kernel'], ks, stride=stride)
    y.F = _bn(y.F, _sub(sd, prefix + '1.'), training)
    if relu:
        y.F = torch.relu(y.F)
    return y


def _residual(x: OTensor, sd, prefix, training) -> OTensor:
    # linkencoder.py:61-91: identity shortcut when inc == outc, else 1x1 conv + BN
    y = _conv_bn(x, sd, prefix + 'net.', 3, relu=True, training=training)
    y2 = conv3d(y, sd[prefix + 'net.3.kernel'], 3)
    y2.F = _bn(y2.F, _sub(sd, prefix + 'net.4.'), training)
    short = x.F
    if prefix + 'downsample.0.kernel' in sd:
        short = _bn(x.F @ sd[prefix + 'downsample.0.kernel'], _sub(sd, prefix + 'downsample.1.'), training)
    y2.F = torch.relu(y2.F + short)
    return y2


def upsample_voxel(xF: torch.Tensor, xC, x_stride, refC) -> torch.Tensor:
    """seg/core/models/utils.py:327-340."""
    stride = make_ntuple(x_stride)[0]
    xC, refC = _np(xC).astype(np.int32), _np(refC).astype(np.int32)
    a = np.concatenate([np.floor_divide(xC[:, :3], np.int32(stride)), xC[:, 3:]], 1).astype(np.int32)
    b = np.concatenate([np.floor_divide(refC[:, :3], np.int32(stride)), refC[:, 3:]], 1).astype(np.int32)
    idx = sphashquery(sphash(b), sphash(a))
    return xF[_t(idx, torch.long)]


def elk_unet_forward(sd: Dict[str, torch.Tensor], feats: torch.Tensor, coords, *, s: int, r: int,
                     baseop: str, groups: int, training: bool = False) -> torch.Tensor:
    """ELKUNet.forward, segmentation/core/models/semantic_kitti/linkunet.py:331-385."""
    _, (x0, x1, x2, x3, x4) = elk_encoder_forward(sd, feats, coords, s=s, r=r, baseop=baseop,
                                                  groups=groups, training=training,
                                                  return_levels=True, variant='unet', head=False)
    y = x4
    for u, skip in ((1, x3), (2, x2), (3, x1), (4, x0)):
        d = conv3d(y, sd[f'up{u}.0.net.0.kernel'], 2, stride=2, transposed=True)
        d.F = torch.relu(_bn(d.F, _sub(sd, f'up{u}.0.net.1.'), training))
        cat = OTensor(torch.cat([d.F, skip.F], dim=1), d.C, d.s)
        cat.cmaps, cat.kmaps = d.cmaps, d.kmaps
        y = _residual(_residual(cat, sd, f'up{u}.1.0.', training), sd, f'up{u}.1.1.', training)
    return TF.linear(y.F, sd['classifier.0.weight'], sd['classifier.0.bias'])


def elk_encoder_forward(sd: Dict[str, torch.Tensor], feats: torch.Tensor, coords, *, s: int,
                        r: int, baseop: str, groups: int, training: bool = False,
                        return_levels: bool = False, variant: str = 'encoder', head: bool = True):
    """ELKEncoder.forward, linkencoder.py:339-381."""
    x = OTensor(feats, coords, 1)
    x.cmaps[x.s] = x.C
    x0 = _conv_bn(x, sd, 'stem.', 3, relu=True, training=training)
    y = conv3d(x0, sd['stem.3.kernel'], 3)
    y.F = torch.relu(_bn(y.F, _sub(sd, 'stem.4.'), training))
    x0 = y
    cur, levels = x0, []
    for l in (1, 2, 3, 4):
        xl0 = _conv_bn(cur, sd, f'down{l}.0.net.', 2, stride=2, relu=True, training=training)
        z = _residual(xl0, sd, f'stage{l}.0.', training)
        z = _residual(z, sd, f'stage{l}.1.', training)
        xl = _conv_bn(z, sd, f'stage{l}_tail.', 3, training=training)
        lk_F = elk_block_forward(xl0.F, xl0.C, xl0.s, _sub(sd, f'elk{l}.'), xl0.s[0] * s, r,
                                 baseop, groups, variant, kmaps=xl0.kmaps)
        lk = OTensor(lk_F, xl0.C, xl0.s)
        lk.cmaps, lk.kmaps = xl0.cmaps, xl0.kmaps
        lk = _conv_bn(lk, sd, f'elk{l}_tail.', 3, training=training)
        xl.F = torch.relu(xl.F + lk.F)
        levels.append(xl)
        cur = xl
    if not head:
        return None, [x0] + levels
    ups = [upsample_voxel(lv.F, lv.C, lv.s, x0.C) for lv in reversed(levels)]
    F_cat = torch.cat(ups + [x0.F], dim=1).unsqueeze(0).permute(0, 2, 1)
    h = torch.relu(TF.conv1d(F_cat, sd['classifier.0.weight'], sd['classifier.0.bias'], groups=5))
    out = TF.conv1d(h, sd['classifier.2.weight'], sd['classifier.2.bias']).squeeze(0).T
    if return_levels:
        return out, [x0] + levels
    return out


# --------------------------------------------------------------------------
# voxelisation front-ends
# --------------------------------------------------------------------------
def ravel_hash(x: np.ndarray) -> np.ndarray:
    """ts/utils/quantize.py:9-21."""
    x = x - x.min(axis=0)
    x = x.astype(np.uint64)
    xmax = x.max(axis=0).astype(np.uint64) + np.uint64(1)
    h = np.zeros(x.shape[0], dtype=np.uint64)
    for k in range(x.shape[1] - 1):
        h += x[:, k]
        h *= xmax[k + 1]
    h += x[:, -1]
    return h


def sparse_quantize(coords, voxel_size=1.0):
    """ts/utils/quantize.py:24-46 -> (coords [N,3] int32, indices, inverse)."""
    vs = np.asarray(make_ntuple(voxel_size) if not isinstance(voxel_size, float)
                    else (voxel_size,) * 3)
    c = np.floor(_np(coords) / vs).astype(np.int32)
    _, ind, inv = np.unique(ravel_hash(c), return_index=True, return_inverse=True)
    # Reference quirk: ravel_hash does `x -= np.min(x, axis=0)` IN PLACE
    # (quantize.py:12), so the coords sparse_quantize returns are min-shifted.
    c = c - c.min(axis=0)
    return c[ind], ind, inv.reshape(-1)


def initial_voxelize(pF: torch.Tensor, pC: torch.Tensor, init_res, after_res):
    """seg/core/models/utils.py:234-254 -> (voxel feats, voxel coords int32 [M,4],
    idx_query [P], counts [M]).  Voxel order = ascending FNV hash (torch.unique)."""
    nf = torch.cat([(pC[:, :3] * init_res) / after_res, pC[:, -1].view(-1, 1)], 1)
    fl = torch.floor(nf)
    pc_hash = sphash(fl.int().numpy())
    sparse_hash = np.unique(pc_hash)
    idx = sphashquery(pc_hash, sparse_hash)
    counts = spcount(idx.astype(np.int32), len(sparse_hash))
    vc = torch.round(spvoxelize(fl, idx, counts)).int()
    vf = spvoxelize(pF, idx, counts)
    return vf, vc.numpy(), idx, counts


# --------------------------------------------------------------------------
# detection voxelisation front-end
# --------------------------------------------------------------------------
def points_to_voxel(points: np.ndarray, voxel_size, coors_range, max_points=35, reverse_index=True,
                    max_voxels=20000):
    """det3d/ops/point_cloud/point_cloud_ops.py:8-56 / 112-183 restated without the sequential
    loop: a point is kept iff floor((p - lo) / vs) is inside round((hi - lo) / vs) (float32
    arithmetic, like the reference's numpy float32 arrays); voxels are numbered by first appearance,
    the first `max_voxels` exist; each keeps its first `max_points` points in point order."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    gs = np.round((cr[3:] - cr[:3]) / vs).astype(np.int32)
    c = np.floor((points[:, :3] - cr[:3]) / vs)                      # float32
    ok = np.all((c >= 0) & (c < gs), axis=1)
    ci = c[ok].astype(np.int64)
    idx = np.nonzero(ok)[0]
    key = (ci[:, 2] * gs[1] + ci[:, 1]) * gs[0] + ci[:, 0]
    uq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    vid_of_cell = np.empty(len(uq), np.int64)
    vid_of_cell[np.argsort(first, kind='stable')] = np.arange(len(uq))
    vid = vid_of_cell[inv.reshape(-1)]
    m = min(len(uq), max_voxels)
    voxels = np.zeros((m, max_points, points.shape[1]), np.float32)
    coors = np.zeros((m, 3), np.int32)
    num = np.zeros(m, np.int32)
    order = np.argsort(vid, kind='stable')                            # points grouped by voxel, in point order
    vs_sorted = vid[order]
    start = np.searchsorted(vs_sorted, np.arange(m))
    rank = np.arange(len(order)) - start[np.minimum(vs_sorted, m - 1)] if m else np.zeros(0, np.int64)
    keep = (vs_sorted < m) & (rank < max_points)
    voxels[vs_sorted[keep], rank[keep]] = points[idx[order[keep]]]
    np.add.at(num, vs_sorted[keep], 1)
    cell_first = ci[first]                                            # (x, y, z) of each cell
    sel = vid_of_cell < m
    coors[vid_of_cell[sel]] = cell_first[sel][:, ::-1] if reverse_index else cell_first[sel]
    return voxels, coors, num
