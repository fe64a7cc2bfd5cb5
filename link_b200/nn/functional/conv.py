"""Sparse 3D convolution: kernel-map build + output-stationary fused gather-GEMM.

Public surface = the reference's `F.conv3d(input, weight, kernel_size, bias, stride, dilation,
transposed)` (torchsparse/nn/functional/conv.py:83-147) with the same caching contract: kernel
maps live in `input.kmaps[(input.stride, kernel_size, stride, dilation)]` and every derived
tensor shares the `cmaps` / `kmaps` dict objects."""
import ctypes as C
import os
from typing import Optional, Tuple, Union

import torch
from torch.autograd import Function

from link_b200 import _capi
from link_b200.nn.functional.downsample import spdownsample
from link_b200.nn.functional.hash import sphash
from link_b200.nn.functional.query import HashTable
from link_b200.nn.utils import get_kernel_offsets
from link_b200.tensor import SparseTensor
from link_b200.utils import make_ntuple

__all__ = ['conv3d', 'conv_bn_act', 'fusable', 'KernelMap', 'build_kernel_map', 'set_precision', 'invalidate_caches']


class KernelMap:
    """Output-stationary kernel map: `nbr` int32 [K, n_out], nbr[k, o] = input row that feeds
    output row o through offset k, or -1 (the reference's `results`, conv.py:114).

    Indexing (`kmap[0]`, `kmap[1]`, `kmap[2]`) yields the reference's list layout
    [nbmaps [P,2] (input row, output row) ordered by (k, output row), nbsizes [K], (n_in, n_out)]
    -- materialised lazily, only parity tests and foreign code need it."""

    def __init__(self, nbr: torch.Tensor, n_in: int, n_out: int, out_coords: torch.Tensor):
        self.nbr = nbr
        self.n_in = n_in
        self.n_out = n_out
        self.out_coords = out_coords
        self._inv = None
        self._ref = None
        self.offsets = None      # int32 [K,3] offsets the map was built with (classes of the plan)
        self._plan = None        # (perm [n_out], tile_mask [tiles]) or False
        self._wgrad = {}         # direction -> (nbrp, perm, masks) of the tensor-core weight gradient
        self.subm = False        # submanifold map (stride 1, odd kernel): inv[k] == nbr[K - 1 - k]

    @property
    def inv(self) -> torch.Tensor:
        """Transposed relation int32 [K, n_in]: inv[k, i] = o whenever nbr[k, o] = i."""
        if self._inv is None:
            k = self.nbr.shape[0]
            inv = torch.empty(k, self.n_in, dtype=torch.int32, device=self.nbr.device)
            _capi.check(_capi.lib().lk_kmap_invert(_capi.ptr(self.nbr), self.n_out, k, self.n_in,
                                                   _capi.ptr(inv), _capi.stream()),
                        'lk_kmap_invert')
            self._inv = inv
        return self._inv

    def plan_buffers(self):
        """Uninitialised buffers of the tile-skipping plan (filled by lk_conv_plan, here or inside
        the native block executor)."""
        k, dev = self.nbr.shape[0], self.nbr.device
        return (torch.empty(self.n_out, dtype=torch.int32, device=dev),
                torch.empty((self.n_out + 127) // 128, dtype=torch.int32, device=dev))

    def plan(self):
        """Tile-skipping plan of this map for the tensor-core conv (lk_conv_plan), built once and
        shared by every conv that reuses the map; None when the map is too wide (K > 32)."""
        if self._plan is None:
            k = self.nbr.shape[0]
            if k > 32 or self.n_out == 0 or not USE_PLAN:
                self._plan = False
            else:
                perm, tmask = self.plan_buffers()
                L = _capi.lib()
                ws_bytes = L.lk_conv_plan_ws_bytes(self.n_out)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.nbr.device)
                with _capi.timed('lk_conv_plan', self.n_out * (4 * k + 8)):
                    _capi.check(L.lk_conv_plan(_capi.ptr(self.nbr), self.n_out, k, _capi.ptr(self.offsets),
                                               _capi.ptr(perm), _capi.ptr(tmask),
                                               _capi.ptr(ws), ws_bytes, _capi.stream()), 'lk_conv_plan')
                self._plan = (perm, tmask)
        return self._plan or None

    def wgrad_relation(self, transposed: bool):
        """Relation of the weight-gradient contraction as the tensor-core kernel walks it
        (lk_conv_wgrad_prepass, once per map and direction, shared by every conv on the map):
        rows = forward OUTPUT rows (grad rows), in the order of the forward tile plan when there is
        one.  Returns (nbrp [K, n], perm [n] or None, masks [ceil(n/64)])."""
        hit = self._wgrad.get(transposed)
        if hit is None:
            rel = self.inv if transposed else self.nbr
            k, n = rel.shape
            plan = None if transposed else self.plan()
            perm = plan[0] if plan is not None else None
            masks = torch.empty((n + 63) // 64, dtype=torch.int32, device=rel.device)
            nbrp = torch.empty_like(rel) if perm is not None else None
            _capi.check(_capi.lib().lk_conv_wgrad_prepass(_capi.ptr(rel), _capi.ptr(perm), n, k, _capi.ptr(nbrp),
                                                          _capi.ptr(masks), _capi.stream()),
                        'lk_conv_wgrad_prepass')
            hit = self._wgrad[transposed] = (nbrp if nbrp is not None else rel, perm, masks)
        return hit

    def _reference_layout(self):
        if self._ref is None:
            hit = self.nbr != -1
            nbsizes = torch.sum(hit, dim=1)
            nbmaps = torch.nonzero(hit)
            nbmaps[:, 0] = self.nbr[hit].long()
            self._ref = [nbmaps, nbsizes, (self.n_in, self.n_out)]
        return self._ref

    def __getitem__(self, i):
        return self._reference_layout()[i]

    def __len__(self):
        return 3

    def __iter__(self):
        return iter(self._reference_layout())


def _table_for(input: SparseTensor) -> HashTable:
    key = ('lk', 'table', input.stride)
    tab = input.kmaps.get(key)
    if tab is None or tab.n != input.coords.shape[0]:
        tab = HashTable(sphash(input.coords.contiguous()))
        input.kmaps[key] = tab
    return tab


def build_kernel_map(input: SparseTensor, kernel_size, stride, dilation, want_plan: bool = False,
                     out_coords: Optional[torch.Tensor] = None) -> KernelMap:
    """Kernel map of a conv over `input` (the kmap branch of the reference's F.conv3d,
    conv.py:103-121) through ONE library call (lk_kmap_build: hash -> table -> query [-> plan]);
    the hash table of the input level is kept in `kmaps` and reused by later maps of that level."""
    coords = input.coords.contiguous()
    dev = coords.device
    # NB: like the reference (conv.py:105-107) the offsets ignore `dilation`.
    offsets = get_kernel_offsets(kernel_size, stride=input.stride, device=dev)
    if out_coords is None:          # (a caller that derived the output sites already passes them in)
        out_coords = coords
        if any(s > 1 for s in stride):
            out_coords = spdownsample(coords, stride, kernel_size, input.stride, cache=input.kmaps)
    k, n_in, n_out = offsets.shape[0], coords.shape[0], out_coords.shape[0]
    L = _capi.lib()
    tkey = ('lk', 'table', input.stride)
    tab = input.kmaps.get(tkey)
    build_table = tab is None or tab.n != n_in
    if build_table:
        tab = HashTable.__new__(HashTable)
        tab.n, tab.capacity = n_in, int(L.lk_table_capacity(n_in))
        tab.table = torch.empty(tab.capacity * 16, dtype=torch.uint8, device=dev)
        input.kmaps[tkey] = tab
    nbr = torch.empty(k, n_out, dtype=torch.int32, device=dev)
    kmap = KernelMap(nbr, n_in, n_out, out_coords)
    kmap.offsets = offsets
    plan = kmap.plan_buffers() if (want_plan and USE_PLAN and k <= 32 and n_out > 0) else None
    subm = all(s == 1 for s in stride) and k % 2 == 1
    ws_bytes = L.lk_kmap_build_ws_bytes(n_in, n_out)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _capi.timed('lk_kmap_query', n_out * (16 + 4 * k)):
        _capi.check(L.lk_kmap_build(coords.data_ptr(), n_in, out_coords.data_ptr(), n_out, offsets.data_ptr(),
                                    k, 1 if subm else 0, tab.table.data_ptr(), tab.capacity,
                                    1 if build_table else 0, nbr.data_ptr(),
                                    plan[0].data_ptr() if plan else None,
                                    plan[1].data_ptr() if plan else None, ws.data_ptr(), ws_bytes,
                                    _capi.stream()), 'lk_kmap_build')
    if plan:
        kmap._plan = plan
    kmap.subm = bool(subm)
    return kmap


# dense / sparse GEMMs on tcgen05 (3xTF32, fp32-level accuracy); '0' selects the FFMA kernel
USE_TENSOR_CORES = os.environ.get('LINKB200_TENSOR_CORES', '1') != '0'
# tile-skipping plan (lk_conv_plan) for the tensor-core conv; '0' runs every (offset, tile) step
USE_PLAN = os.environ.get('LINKB200_CONV_PLAN', '1') != '0'
# weight gradient on tcgen05 (lk_conv_wgrad_tc); '0' keeps the FFMA kernel (lk_conv_bwd_weight)
USE_TC_WGRAD = os.environ.get('LINKB200_TC_WGRAD', '1') != '0'
WGRAD_SLOTS = int(os.environ.get('LINKB200_WGRAD_SLOTS', '0'))
# Arithmetic of the tensor-core convs: 'fp32' (default) = 3xTF32, fp32-level accuracy (the parity
# configuration); 'tf32' = single-pass TF32 -- operands truncated to tf32 by the tensor core, ~1e-3
# relative error per product, fp32 accumulation: the reduced-precision mode for training / throughput
# (the reference runs its conv in fp16 under autocast, nn/functional/conv.py:19).
PRECISION = os.environ.get('LINKB200_CONV_PRECISION', 'fp32')


def set_precision(p: str) -> None:
    global PRECISION
    if p not in ('fp32', 'tf32'):
        raise ValueError("precision must be 'fp32' or 'tf32'")
    PRECISION = p


def precision_code() -> int:
    return 1 if PRECISION == 'tf32' else 0


_tc_supported = {}


def _pack_ex(w: torch.Tensor, layout: int, ci: int, co: int, flip_k: bool = False) -> torch.Tensor:
    """Packed tensor-core image (ci x co channels, zero padded) straight from `w`: [K, Cin, Cout] (layout 1,
    the module parameter) or [K, Cout, Cin] (layout 0), optionally with the offsets reversed -- the
    transposition, the padding and the flip happen inside lk_conv_tc_pack_weights_ex, not as torch copies."""
    w = w.detach().contiguous().float()
    k = w.shape[0]
    src_ci, src_co = (w.shape[1], w.shape[2]) if layout else (w.shape[2], w.shape[1])
    img = torch.empty(k * 2 * ci * co, dtype=torch.float32, device=w.device)
    _capi.check(_capi.lib().lk_conv_tc_pack_weights_ex(_capi.ptr(w), k, ci, co, src_ci, src_co, layout,
                                                       1 if flip_k else 0, _capi.ptr(img), _capi.stream()),
                'lk_conv_tc_pack_weights_ex')
    return img


def _tc_image(weight: torch.Tensor, c_pad: int = 0, c_pad_out: int = 0, cache_on=None) -> torch.Tensor:
    """Packed tensor-core image of a [K, Cin, Cout] kernel (input / output channels zero-padded to
    `c_pad` / `c_pad_out` if given).  The image is cached ON the parameter object it derives from
    (`weight` itself when it is an nn.Parameter, else `cache_on`), keyed by its version and storage
    address, so a module's weights are packed once per update; the cache lives and dies with the
    parameter, so a recycled address can never alias another tensor's entry."""
    holder = weight if isinstance(weight, torch.nn.Parameter) else cache_on
    if holder is not None:
        ver = (holder._version, holder.data_ptr(), c_pad, c_pad_out)
        hit = holder.__dict__.get('_lk_img')
        if hit is not None and hit[0] == ver:
            return hit[1]
    img = _pack_ex(weight, 1, max(c_pad, weight.shape[1]), max(c_pad_out, weight.shape[2]))
    if holder is not None:
        holder.__dict__['_lk_img'] = (ver, img)
    return img


def invalidate_caches(module: torch.nn.Module) -> None:
    """Drop every derived copy cached on a model's modules / parameters (packed weight images, folded
    BatchNorm affines, the block executor's argument templates).  The caches are keyed by the
    parameters' `_version` and storage address, so optimizer steps, `load_state_dict` and `copy_` refresh
    them by themselves; writes through `.data` (EMA `p.data.mul_()`, clamping) do NOT bump the version --
    call this after such an update."""
    for m in module.modules():
        for key in ('_lk_fold', '_lk_native_args', '_lk_kio', '_lk_refs', '_lk_dil', '_lk_enc_native', '_lk_native_static',
                    '_lk_elk_params'):
            m.__dict__.pop(key, None)
    for p in module.parameters():
        p.__dict__.pop('_lk_img', None)
        p.__dict__.pop('_lk_kio', None)


def _tc_ok(L, c_in, c_out) -> bool:
    ok = _tc_supported.get((c_in, c_out))
    if ok is None:
        ok = _tc_supported[(c_in, c_out)] = bool(L.lk_conv_tc_supported(c_in, c_out))
    return ok


def _pad_to_tc(c: int) -> int:
    """Smallest tensor-core channel count >= c (0 if none)."""
    for t in (32, 64, 128):
        if c <= t:
            return t
    return 0


# bf16 feature rows on the tensor-core conv itself (lk_conv_tc_fwd_bf16); '0': convert at the op boundary
BF16_NATIVE = os.environ.get('LINKB200_BF16_NATIVE', '1') != '0'


def _conv_fwd_bf16(feats, weight, nbr, n_out, weight_t, scale, shift, residual, relu, kmap, cache_on, k, c_in, c_out,
                   flip_k=False):
    """bf16 rows in / bf16 rows out on the tensor-core kernel (fp32 accumulation, tf32 weights)."""
    L = _capi.lib()
    ci, co = _pad_to_tc(c_in), _pad_to_tc(c_out)
    if ci != c_in:
        feats = torch.nn.functional.pad(feats, (0, ci - c_in))
    if weight_t is not None:
        img = _pack_ex(weight_t, 0, ci, co, flip_k)
    else:
        img = _tc_image(weight, ci if ci != c_in else 0, co if co != c_out else 0, cache_on=cache_on)
    fuse_tail = co == c_out
    ep = _capi.ConvEpilogue()
    if co != c_out:
        scale = torch.nn.functional.pad(scale, (0, co - c_out), value=1.0) if scale is not None else None
        shift = torch.nn.functional.pad(shift, (0, co - c_out)) if shift is not None else None
    if residual is not None and fuse_tail:
        residual = residual.to(torch.bfloat16).contiguous()
    ep.d_scale, ep.d_shift = _capi.ptr(scale), _capi.ptr(shift)
    ep.d_residual = _capi.ptr(residual) if (fuse_tail and residual is not None) else None
    ep.relu = 1 if (relu and (fuse_tail or residual is None)) else 0
    out = torch.empty(n_out, co, dtype=torch.bfloat16, device=feats.device)
    plan = kmap.plan() if kmap is not None else None
    nb = n_out * (4 * k + 2 * c_out) + feats.shape[0] * 2 * c_in + 4 * k * c_in * c_out
    with _capi.timed('lk_conv_fwd', nb):
        _capi.check(L.lk_conv_tc_fwd_bf16(_capi.ptr(feats, torch.bfloat16), _capi.ptr(img, torch.float32),
                                          _capi.ptr(nbr, torch.int32),
                                          _capi.ptr(plan[0]) if plan is not None else None,
                                          _capi.ptr(plan[1]) if plan is not None else None,
                                          n_out, k, ci, co, C.byref(ep), _capi.ptr(out), _capi.stream()),
                    'lk_conv_tc_fwd_bf16')
    if not fuse_tail:
        out = out[:, :c_out]
        if residual is not None:
            out = out + residual.to(out.dtype)
            if relu:
                out = torch.relu_(out)
        out = out.contiguous()
    return out


def _conv_fwd(feats, weight, nbr, n_out, weight_t=None, scale=None, shift=None, residual=None,
              relu=False, kmap=None, cache_on=None, flip_k=False):
    """out[o] = epilogue(sum_k feats[nbr[k, o]] @ weight[k]).  `weight` is [K, Cin, Cout] (may be
    None when its transpose `weight_t` [K, Cout, Cin] is given and the tensor-core kernel applies).
    epilogue: y = relu?(acc * scale + shift + residual), each part optional.  `cache_on`: the
    nn.Parameter that `weight` is a view of (its packed image is cached there).  The result has the
    dtype of `feats`: fp32 rows run in fp32 (3xTF32 / TF32), bf16 rows on the bf16 kernel where it
    covers the shape (c_out <= 64 after padding), else through fp32 at this boundary."""
    if weight is not None:
        k, c_in, c_out = weight.shape
    else:
        k, c_out, c_in = weight_t.shape
    if feats.shape[1] != c_in:
        raise ValueError('Input feature size and kernel size mismatch')   # convolution_cuda.cu:57
    _capi.check_device(feats)
    L = _capi.lib()
    in_dtype = feats.dtype
    if in_dtype != torch.float32:
        if (in_dtype == torch.bfloat16 and BF16_NATIVE and USE_TENSOR_CORES and k <= 32 and _pad_to_tc(c_in)
                and _pad_to_tc(c_out) in (32, 64)):
            return _conv_fwd_bf16(feats.contiguous(), weight, nbr, n_out, weight_t, scale, shift, residual, relu, kmap,
                                  cache_on, k, c_in, c_out, flip_k)
        out = _conv_fwd(feats.float(), weight.float() if weight is not None else None, nbr, n_out,
                        weight_t.float() if weight_t is not None else None, scale, shift,
                        residual.float() if residual is not None else None, relu, kmap, cache_on, flip_k)
        return out.to(in_dtype)
    # algorithmic bytes: kernel map + each input row once + output once + the weights
    nb = n_out * (4 * k + 4 * c_out) + feats.shape[0] * 4 * c_in + 4 * k * c_in * c_out
    if residual is not None:
        assert residual.shape == (n_out, c_out) and residual.dtype == torch.float32
    if USE_TENSOR_CORES and k <= 32 and _pad_to_tc(c_in) and _pad_to_tc(c_out):
        # tensor-core kernel; channel counts other than 32 / 64 / 128 (the 4-channel stem, the 5- and
        # 16-channel layers of the detection backbone) are zero-padded to the next supported size:
        # the padded K-columns / output columns cost tensor time only, and the FFMA kernel that
        # would serve them otherwise is ~5x slower at every size
        ci, co = _pad_to_tc(c_in), _pad_to_tc(c_out)
        if ci != c_in:
            feats = torch.nn.functional.pad(feats, (0, ci - c_in))
        if weight_t is not None:
            img = _pack_ex(weight_t, 0, ci, co, flip_k)
        else:
            img = _tc_image(weight, ci if ci != c_in else 0, co if co != c_out else 0, cache_on=cache_on)
        fuse_tail = co == c_out               # residual / ReLU stay in the epilogue unless C_out is padded
        ep = _capi.ConvEpilogue()
        if co != c_out:
            scale = torch.nn.functional.pad(scale, (0, co - c_out), value=1.0) if scale is not None else None
            shift = torch.nn.functional.pad(shift, (0, co - c_out)) if shift is not None else None
        ep.d_scale, ep.d_shift = _capi.ptr(scale), _capi.ptr(shift)
        ep.d_residual = _capi.ptr(residual) if fuse_tail else None
        ep.relu = 1 if (relu and (fuse_tail or residual is None)) else 0
        ep.precision = precision_code()
        out = torch.empty(n_out, co, dtype=torch.float32, device=feats.device)
        plan = kmap.plan() if kmap is not None else None      # only for the forward map (kmap.nbr)
        with _capi.timed('lk_conv_fwd', nb):
            _capi.check(L.lk_conv_tc_fwd_plan(_capi.ptr(feats, torch.float32), _capi.ptr(img, torch.float32),
                                              _capi.ptr(nbr, torch.int32),
                                              _capi.ptr(plan[0]) if plan is not None else None,
                                              _capi.ptr(plan[1]) if plan is not None else None,
                                              n_out, k, ci, co, C.byref(ep), _capi.ptr(out),
                                              _capi.stream()), 'lk_conv_tc_fwd_plan')
        if not fuse_tail:
            out = out[:, :c_out]
            if residual is not None:
                out = out + residual
                if relu:
                    out = torch.relu_(out)
            out = out.contiguous()
        return out
    out = torch.empty(n_out, c_out, dtype=torch.float32, device=feats.device)
    ep = _capi.ConvEpilogue()
    ep.d_scale, ep.d_shift = _capi.ptr(scale), _capi.ptr(shift)
    ep.d_residual = _capi.ptr(residual)
    ep.relu = 1 if relu else 0
    if weight is None:
        weight = (weight_t.flip(0) if flip_k else weight_t).transpose(1, 2).contiguous()
    with _capi.timed('lk_conv_fwd', nb):
        _capi.check(L.lk_conv_fwd_ex(_capi.ptr(feats, torch.float32), _capi.ptr(weight, torch.float32),
                                     _capi.ptr(nbr, torch.int32), n_out, k, c_in, c_out, C.byref(ep),
                                     _capi.ptr(out), _capi.stream()), 'lk_conv_fwd')
    return out


class ConvolutionFunction(Function):
    """Autograd wrapper (reference: ConvolutionFunction, conv.py:16-80).  forward and both
    backward products run on liblinkb200 kernels; fp32 compute."""

    @staticmethod
    def forward(ctx, feats, weight, kmap: KernelMap, transposed: bool = False):
        in_dtype = feats.dtype
        feats = feats.contiguous()
        if in_dtype != torch.bfloat16:          # bf16 rows stay bf16 (native kernel or boundary conversion in _conv_fwd)
            feats = feats.float()
        weight = weight.contiguous().float()
        if not transposed:
            out = _conv_fwd(feats, weight, kmap.nbr, kmap.n_out, kmap=kmap)
        else:
            out = _conv_fwd(feats, weight, kmap.inv, kmap.n_in)
        ctx.save_for_backward(feats, weight)
        ctx.kmap, ctx.transposed, ctx.in_dtype = kmap, transposed, in_dtype
        return out.to(in_dtype)

    @staticmethod
    def backward(ctx, grad_output):
        feats, weight = ctx.saved_tensors
        kmap, transposed = ctx.kmap, ctx.transposed
        g = grad_output.contiguous()
        if not (g.dtype == torch.bfloat16 and feats.dtype == torch.bfloat16):
            g, feats = g.float(), feats.float()
        k, c_in, c_out = weight.shape
        grad_feats = grad_weight = None
        # the relation seen from the forward INPUT rows / from the forward OUTPUT rows
        to_out = kmap.nbr if not transposed else kmap.inv
        n_in_rows = feats.shape[0]
        if ctx.needs_input_grad[0]:
            # dX = sum_k dY[to_in[k]] @ W[k]^T: the forward weight IS the transposed operand
            if kmap.subm and not transposed:
                # submanifold map: input i reaches output o through offset k exactly when o reaches i
                # through -offset[k] = offset[K-1-k], so inv[k] == nbr[K-1-k]: the gradient runs on the
                # FORWARD map with the offsets of W reversed -- no inverted map is built, and the forward
                # tile-skipping plan applies (the unplanned dgrad ran all 27 x tiles steps: 173 vs 73 us
                # at N = 119k, C = 64)
                grad_feats = _conv_fwd(g, None, kmap.nbr, n_in_rows, weight_t=weight, kmap=kmap, flip_k=True).to(ctx.in_dtype)
            else:
                to_in = kmap.inv if not transposed else kmap.nbr
                grad_feats = _conv_fwd(g, None, to_in, n_in_rows, weight_t=weight).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            grad_weight = torch.empty_like(weight)
            L = _capi.lib()
            feats, g = feats.float(), g.float()          # the weight gradient contracts fp32 rows (3xTF32)
            ci, co = _pad_to_tc(c_in), _pad_to_tc(c_out)
            if USE_TENSOR_CORES and USE_TC_WGRAD and k <= 32 and ci and co and g.shape[0] * k < 2 ** 31:
                # channel counts other than 32 / 64 / 128 (the 5- and 16-channel layers of the detection
                # backbone) are zero-padded like in the forward: the FFMA kernel is ~4x slower there too
                nbrp, perm, masks = kmap.wgrad_relation(transposed)
                f_p = feats if ci == c_in else torch.nn.functional.pad(feats, (0, ci - c_in))
                g_p = g if co == c_out else torch.nn.functional.pad(g, (0, co - c_out))
                gw_p = grad_weight if (ci == c_in and co == c_out) else torch.empty(k, ci, co, dtype=torch.float32,
                                                                                    device=g.device)
                _capi.check(L.lk_conv_wgrad_tc(
                    _capi.ptr(f_p), _capi.ptr(g_p), _capi.ptr(nbrp), _capi.ptr(perm), _capi.ptr(masks),
                    g.shape[0], k, ci, co, _capi.ptr(gw_p), WGRAD_SLOTS, _capi.stream()),
                    'lk_conv_wgrad_tc')
                if gw_p is not grad_weight:
                    grad_weight = gw_p[:, :c_in, :c_out].contiguous()
            else:
                _capi.check(L.lk_conv_bwd_weight(
                    _capi.ptr(feats), _capi.ptr(g), _capi.ptr(to_out), g.shape[0], k, c_in, c_out,
                    _capi.ptr(grad_weight), _capi.stream()), 'lk_conv_bwd_weight')
        return grad_feats, grad_weight, None, None


def _conv_autograd(feats: torch.Tensor, weight: torch.Tensor, kmap: KernelMap) -> torch.Tensor:
    """Non-transposed conv with autograd.  Training on fp32 rows at a tensor-core width takes the C++
    autograd node (link_b200/_ext.py: same kernels, the glue and the backward node outside the
    interpreter); everything else the python ConvolutionFunction."""
    from link_b200 import _ext
    k, c_in, c_out = weight.shape
    if (torch.is_grad_enabled() and (feats.requires_grad or weight.requires_grad) and feats.is_cuda
            and feats.dtype == torch.float32 and weight.dtype == torch.float32 and USE_TENSOR_CORES and USE_TC_WGRAD
            and USE_PLAN and k <= 32 and c_in in (32, 64, 128) and c_out in (32, 64, 128) and kmap.n_out > 0
            and kmap.n_out * k < 2 ** 31 and feats.shape[1] == c_in):
        ext = _ext.module()
        plan = kmap.plan() if ext is not None else None
        if plan is not None:
            _capi.check_device(feats)
            wg = kmap.wgrad_relation(False) if weight.requires_grad else (None, None, None)
            inv = kmap.inv if (feats.requires_grad and not kmap.subm) else None
            return ext.conv(feats, weight, kmap.nbr, plan[0], plan[1], inv, wg[0], wg[1], wg[2], kmap.subm,
                            precision_code(), WGRAD_SLOTS)
    return ConvolutionFunction.apply(feats, weight, kmap, False)


def _folded_bn(bn):
    """Eval-mode BatchNorm as a per-channel affine (scale, shift), cached until a parameter or
    running statistic of the module changes."""
    ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
           bn.weight.data_ptr(), bn.running_mean.data_ptr())
    hit = bn.__dict__.get('_lk_fold')           # cached on the module object itself
    if hit is not None and hit[0] == ver:
        return hit[1], hit[2]
    with torch.no_grad():
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
        shift = (bn.bias - bn.running_mean * scale).float().contiguous()
    bn.__dict__['_lk_fold'] = (ver, scale, shift)
    return scale, shift


def fusable(conv, bn, x: SparseTensor) -> bool:
    """True when conv (+ eval-mode BatchNorm) can run as ONE kernel with a fused epilogue:
    inference only (no autograd, BN in eval mode), fp32, a real sparse kernel (volume > 1)."""
    return (not torch.is_grad_enabled() and conv.kernel_volume > 1 and conv.bias is None
            and x.feats.dtype in (torch.float32, torch.bfloat16) and x.feats.is_cuda
            and (bn is None or (not bn.training and bn.track_running_stats and bn.affine)))


def conv_bn_act(input: SparseTensor, conv, bn=None, relu: bool = False,
                residual: Optional[torch.Tensor] = None) -> SparseTensor:
    """Inference fast path for  relu?(BN_eval(conv(x)) + residual):  the reference's three or four
    separate ops (spnn.Conv3d -> spnn.BatchNorm -> [+ shortcut] -> spnn.ReLU, linkencoder.py:26-37,
    64-91) in one sparse-conv launch whose epilogue applies the folded BatchNorm affine, the
    residual and the ReLU before the single output store.  Same kernel-map caching as conv3d."""
    feats = input.feats
    if not feats.is_contiguous():
        feats = feats.contiguous()
    kernel_size, stride = conv.kernel_size, conv.stride
    dilation = conv.__dict__.get('_lk_dil')
    if dilation is None:
        dilation = conv.__dict__['_lk_dil'] = make_ntuple(conv.dilation, ndim=3)
    scale = shift = None
    if bn is not None:
        scale, shift = _folded_bn(bn)
    w = conv.kernel            # the Parameter itself: its transposed copy is cached on it
    if not conv.transposed:
        key = (input.stride, kernel_size, stride, dilation)
        kmap = input.kmaps.get(key)
        if kmap is None:
            kmap = build_kernel_map(input, kernel_size, stride, dilation,
                                    want_plan=USE_TENSOR_CORES and w.shape[0] <= 32 and w.shape[1] <= 128 and w.shape[2] <= 128)
            input.kmaps[key] = kmap
        out = _conv_fwd(feats, w, kmap.nbr, kmap.n_out, None, scale, shift, residual, relu, kmap=kmap)
        output = SparseTensor(coords=kmap.out_coords, feats=out,
                              stride=tuple(input.stride[k] * stride[k] for k in range(3)))
    else:
        tensor_stride = tuple(input.stride[k] // stride[k] for k in range(3))
        kmap = input.kmaps[(tensor_stride, kernel_size, stride, dilation)]
        out = _conv_fwd(feats, w, kmap.inv, kmap.n_in, None, scale, shift, residual, relu)
        output = SparseTensor(coords=input.cmaps[tensor_stride], feats=out, stride=tensor_stride)
    output.cmaps = input.cmaps
    output.cmaps.setdefault(output.stride, output.coords)
    output.kmaps = input.kmaps
    return output


def conv3d(input: SparseTensor, weight: torch.Tensor,
           kernel_size: Union[int, Tuple[int, ...]], bias: Optional[torch.Tensor] = None,
           stride: Union[int, Tuple[int, ...]] = 1, dilation: Union[int, Tuple[int, ...]] = 1,
           transposed: bool = False) -> SparseTensor:
    feats, coords = input.feats, input.coords
    kernel_size = make_ntuple(kernel_size, ndim=3)
    stride = make_ntuple(stride, ndim=3)
    dilation = make_ntuple(dilation, ndim=3)

    if kernel_size == (1, 1, 1) and stride == (1, 1, 1) and dilation == (1, 1, 1):
        feats = feats.matmul(weight)
        if bias is not None:
            feats += bias
        output = SparseTensor(coords=coords, feats=feats, stride=input.stride)
    elif not transposed:
        key = (input.stride, kernel_size, stride, dilation)
        kmap = input.kmaps.get(key)
        if kmap is None:
            kmap = build_kernel_map(input, kernel_size, stride, dilation)
            input.kmaps[key] = kmap
        feats = _conv_autograd(feats, weight, kmap)
        if bias is not None:
            feats += bias
        output = SparseTensor(coords=kmap.out_coords, feats=feats,
                              stride=tuple(input.stride[k] * stride[k] for k in range(3)))
    else:
        tensor_stride = tuple(input.stride[k] // stride[k] for k in range(3))
        kmap = input.kmaps[(tensor_stride, kernel_size, stride, dilation)]
        feats = ConvolutionFunction.apply(feats, weight, kmap, True)
        if bias is not None:
            feats += bias
        output = SparseTensor(coords=input.cmaps[tensor_stride], feats=feats, stride=tensor_stride)

    output.cmaps = input.cmaps
    output.cmaps.setdefault(output.stride, output.coords)
    output.kmaps = input.kmaps
    return output
