"""ELKEncoder (a.k.a. LinKEncoder): the LinK encoder-only segmentation backbone.

Same constructor kwargs (`num_classes, cr, baseop, r, s, groups, run_up`), sub-module names and
parameter shapes as the reference (segmentation/core/models/semantic_kitti/linkencoder.py:188-381)
so that `seg/core/builder.make_model` can instantiate it and published state dicts load; every
sparse op underneath runs on liblinkb200."""
import os

import torch
import torch.nn as nn

import link_b200.nn as spnn
import link_b200.nn.functional as F
import ctypes as C

from link_b200 import _capi
from link_b200.elk import ELKBlock, upsample_index, upsample_voxel
from link_b200.tensor import SparseTensor
from link_b200.utils import make_ntuple

# inference: run the LinK block of a level on a side stream next to the level's conv stage
# (LINKB200_BRANCH_OVERLAP=0: one stream, the reference's order).  Inside the native executor the fork /
# join are two event calls from C; the same schedule driven from python (LINKB200_PY_BRANCH_OVERLAP=1,
# per-layer path) measured SLOWER (2.8 vs 2.5 ms per scan): the coarse levels are paced by the
# interpreter, and the stream switches add host time exactly there.
BRANCH_OVERLAP = os.environ.get('LINKB200_BRANCH_OVERLAP', '1') != '0'
PY_BRANCH_OVERLAP = os.environ.get('LINKB200_PY_BRANCH_OVERLAP', '0') == '1'
_BRANCH_STREAMS = {}
# coordinate pyramid of plan_levels as four parallel chains on side streams (0: level by level on the current stream)
PARALLEL_PYRAMID = os.environ.get('LINKB200_PARALLEL_PYRAMID', '1') != '0'
_PYRAMID_STREAMS = {}


def _pyramid_streams(device):
    st = _PYRAMID_STREAMS.get(device)
    if st is None:
        st = _PYRAMID_STREAMS[device] = [torch.cuda.Stream(device) for _ in range(4)]
    return st

# inference: the whole backbone behind one library call (lk_elk_encoder_fwd); '0': one call per layer
NATIVE_ENCODER = os.environ.get('LINKB200_NATIVE_ENCODER', '1') != '0'


def _branch_stream(device) -> 'torch.cuda.Stream':
    st = _BRANCH_STREAMS.get(device)
    if st is None:
        st = _BRANCH_STREAMS[device] = torch.cuda.Stream(device)
    return st


__all__ = ['ELKEncoder', 'LinKEncoder', 'BasicConvolutionBlock', 'BasicDeconvolutionBlock',
           'ResidualBlock']


def _conv_bn_act(x: SparseTensor, conv, bn, relu: bool, residual=None) -> SparseTensor:
    """Conv3d -> BatchNorm [-> + residual] [-> ReLU] outside the inference fast path: the conv is its
    autograd Function, BatchNorm + add + ReLU ONE fused op (training mode: csrc/bn.cu; else PyTorch's)."""
    y = conv(x)
    y.F = F.batch_norm_act(y.F, bn, relu, residual)
    return y


class BasicConvolutionBlock(nn.Module):
    """Conv3d -> BatchNorm -> ReLU (linkencoder.py:23-39)."""

    def __init__(self, inc, outc, ks=3, stride=1, dilation=1):
        super().__init__()
        self.net = nn.Sequential(
            spnn.Conv3d(inc, outc, kernel_size=ks, dilation=dilation, stride=stride),
            spnn.BatchNorm(outc),
            spnn.ReLU(True),
        )

    def forward(self, x):
        if F.fusable(self.net[0], self.net[1], x):
            return F.conv_bn_act(x, self.net[0], self.net[1], relu=True)
        return _conv_bn_act(x, self.net[0], self.net[1], True)


class BasicDeconvolutionBlock(nn.Module):
    """Transposed Conv3d -> BatchNorm -> ReLU (linkencoder.py:42-58)."""

    def __init__(self, inc, outc, ks=3, stride=1):
        super().__init__()
        self.net = nn.Sequential(
            spnn.Conv3d(inc, outc, kernel_size=ks, stride=stride, transposed=True),
            spnn.BatchNorm(outc),
            spnn.ReLU(True),
        )

    def forward(self, x):
        if F.fusable(self.net[0], self.net[1], x):
            return F.conv_bn_act(x, self.net[0], self.net[1], relu=True)
        return _conv_bn_act(x, self.net[0], self.net[1], True)


class ResidualBlock(nn.Module):
    """Two 3^3 convs with an identity / 1x1 shortcut (linkencoder.py:61-91)."""

    def __init__(self, inc, outc, ks=3, stride=1, dilation=1):
        super().__init__()
        self.net = nn.Sequential(
            spnn.Conv3d(inc, outc, kernel_size=ks, dilation=dilation, stride=stride),
            spnn.BatchNorm(outc),
            spnn.ReLU(True),
            spnn.Conv3d(outc, outc, kernel_size=ks, dilation=dilation, stride=1),
            spnn.BatchNorm(outc),
        )
        if inc == outc and stride == 1:
            self.downsample = nn.Sequential()
        else:
            self.downsample = nn.Sequential(
                spnn.Conv3d(inc, outc, kernel_size=1, dilation=1, stride=stride),
                spnn.BatchNorm(outc),
            )
        self.relu = spnn.ReLU(True)

    def forward(self, x):
        if F.fusable(self.net[0], self.net[1], x) and F.fusable(self.net[3], self.net[4], x):
            # conv+BN+ReLU, then conv+BN+shortcut+ReLU: two launches instead of seven
            shortcut = self.downsample(x).F if len(self.downsample) else x.F
            y = F.conv_bn_act(x, self.net[0], self.net[1], relu=True)
            return F.conv_bn_act(y, self.net[3], self.net[4], relu=True, residual=shortcut.contiguous())
        shortcut = self.downsample(x).F if len(self.downsample) else x.F
        y = _conv_bn_act(x, self.net[0], self.net[1], True)
        return _conv_bn_act(y, self.net[3], self.net[4], True, shortcut)


class ConvBN(nn.Sequential):
    """Conv3d -> BatchNorm (`stageN_tail` / `elkN_tail`, linkencoder.py:216-224); same state-dict
    keys as the reference's nn.Sequential, one fused launch in eval mode.  `residual` / `relu`
    additionally fold the level's merge  relu(x_conv + x_lk)  (linkencoder.py:350) into it."""

    def forward(self, x, residual=None, relu=False):
        if F.fusable(self[0], self[1], x):
            return F.conv_bn_act(x, self[0], self[1], relu=relu, residual=residual)
        return _conv_bn_act(x, self[0], self[1], relu, residual)


def _tail(inc, outc):
    return ConvBN(spnn.Conv3d(inc, outc, kernel_size=3, stride=1), spnn.BatchNorm(outc))


class _ELKBackbone(nn.Module):
    """Stem + four (conv stage || LinK block) levels + the decoder branches, with the reference's
    sub-module names (linkencoder.py:188-320 == linkunet.py:188-320)."""

    block_variant = 'encoder'

    def __init__(self, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        cr = kwargs.get('cr', 1.0)
        baseop = kwargs.get('baseop')
        groups = kwargs.get('groups')
        cs = [int(cr * 64)] * 9
        self.cs = cs
        self.run_up = kwargs.get('run_up', True)

        self.stem = nn.Sequential(
            spnn.Conv3d(4, cs[0], kernel_size=3, stride=1), spnn.BatchNorm(cs[0]), spnn.ReLU(True),
            spnn.Conv3d(cs[0], cs[0], kernel_size=3, stride=1), spnn.BatchNorm(cs[0]), spnn.ReLU(True))

        for lv in (1, 2, 3, 4):
            cin, cout = cs[lv - 1], cs[lv]
            setattr(self, f'down{lv}', nn.Sequential(
                BasicConvolutionBlock(cin, cin, ks=2, stride=2, dilation=1)))
            setattr(self, f'stage{lv}', nn.Sequential(
                ResidualBlock(cin, cout, ks=3, stride=1, dilation=1),
                ResidualBlock(cout, cout, ks=3, stride=1, dilation=1)))
            setattr(self, f'stage{lv}_tail', _tail(cout, cout))
            setattr(self, f'elk{lv}', ELKBlock(cin, cin, groups, baseop=baseop,
                                               variant=self.block_variant))
            setattr(self, f'elk{lv}_tail', _tail(cin, cout))
            setattr(self, f'activate{lv}', nn.ReLU(True))

        # Decoder branches: used by ELKUNet; in ELKEncoder they exist in the reference's state dict
        # but are never run (linkencoder.py:289-320 vs 339-381) -- kept so checkpoints load strictly.
        for u, (cin, cskip, cout) in enumerate([(cs[4], cs[3], cs[5]), (cs[5], cs[2], cs[6]),
                                                (cs[6], cs[1], cs[7]), (cs[7], cs[0], cs[8])], 1):
            setattr(self, f'up{u}', nn.ModuleList([
                BasicDeconvolutionBlock(cin, cout, ks=2, stride=2),
                nn.Sequential(ResidualBlock(cout + cskip, cout, ks=3, stride=1, dilation=1),
                              ResidualBlock(cout, cout, ks=3, stride=1, dilation=1))]))

    def weight_initialization(self):
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.LayerNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _pyramid_parallel(self, x: SparseTensor):
        """Output sites of the four strided levels, or None when the fast path does not apply.  The sites
        of level l are the distinct values of floor(c0 / 2^l) 2^l, so every level derives from the INPUT
        coordinates: four independent pack -> sort -> unique -> unpack chains (lk_downsample) on four side
        streams that wait only for the coordinates, and the four size read-backs synchronise those side
        streams only.  In a training loop the current stream still holds the previous step's backward when
        the next forward starts; read-backs on it (the level-by-level order) drain the device at every
        step, these do not (SparseTensor.from_host(ahead=True) uploads the coordinates off-stream too)."""
        from link_b200.nn.functional import _index
        coords = x.coords
        if not (PARALLEL_PYRAMID and coords.is_cuda and coords.shape[0] > 0 and tuple(x.stride) == (1, 1, 1)):
            return None
        convs = [getattr(self, f'down{lv}')[0].net[0] for lv in (1, 2, 3, 4)]
        if not all(tuple(c.kernel_size) == (2, 2, 2) and tuple(c.stride) == (2, 2, 2) for c in convs):
            return None
        if ((1, 1, 1), (2, 2, 2), (2, 2, 2), make_ntuple(convs[0].dilation, ndim=3)) in x.kmaps:
            return None                                   # maps already built for this scan
        L = _capi.lib()
        coords = coords.contiguous()
        n0, dev = coords.shape[0], coords.device
        bounds = _index.coord_bounds(coords, x.kmaps)
        main = torch.cuda.current_stream(dev)
        ev0 = getattr(x, '_coords_ready', None)
        if ev0 is None:
            ev0 = torch.cuda.Event()
            ev0.record(main)
        side = _pyramid_streams(dev)
        ws_bytes = L.lk_downsample_ws_bytes(n0)
        outs, nums = [], []
        for l in (1, 2, 3, 4):
            ss = (2 ** l,) * 3
            spec, bits = _index.make_keyspec(bounds, ss, (3, 0, 1, 2), ss)
            with torch.cuda.stream(side[l - 1]):
                side[l - 1].wait_event(ev0)
                out = torch.empty(n0, 4, dtype=torch.int32, device=dev)
                num = torch.empty(1, dtype=torch.int32, device=dev)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
                _capi.check(L.lk_downsample(_capi.ptr(coords, torch.int32), n0, C.byref(spec), bits, out.data_ptr(),
                                            num.data_ptr(), ws.data_ptr(), ws_bytes, _capi.stream()), 'lk_downsample')
            outs.append(out)
            nums.append(num)
        sizes = []
        for l in (1, 2, 3, 4):
            with torch.cuda.stream(side[l - 1]):
                sizes.append(int(nums[l - 1].item()))     # synchronises this side stream only
        levels = []
        for l in (1, 2, 3, 4):
            main.wait_stream(side[l - 1])
            outs[l - 1].record_stream(main)
            c_l = outs[l - 1][:sizes[l - 1]]
            _index.register_floored_bounds(x.kmaps, c_l, bounds, (2 ** l,) * 3)
            levels.append(c_l)
        return levels

    def plan_levels(self, x: SparseTensor) -> None:
        """Index-only prologue: builds the coordinate pyramid (the four stride-2 kernel maps and
        their output coordinates) before any feature kernel is enqueued.  The output size of each
        strided conv is data dependent (one 4-byte read-back per level); doing those read-backs
        here, while only tiny index kernels are in flight, lets the whole feature pipeline that
        follows be enqueued without a single host<->device synchronisation on the current stream."""
        pyramid = self._pyramid_parallel(x)
        t = x
        for lv in (1, 2, 3, 4):
            conv = getattr(self, f'down{lv}')[0].net[0]
            dil = make_ntuple(conv.dilation, ndim=3)
            key = (t.stride, conv.kernel_size, conv.stride, dil)
            kmap = t.kmaps.get(key)
            if kmap is None:
                kmap = F.build_kernel_map(t, conv.kernel_size, conv.stride, dil,
                                          out_coords=pyramid[lv - 1] if pyramid is not None else None)
                t.kmaps[key] = kmap
            nxt = SparseTensor(t.feats, kmap.out_coords,
                               tuple(t.stride[k] * conv.stride[k] for k in range(3)))
            nxt.cmaps, nxt.kmaps = t.cmaps, t.kmaps
            nxt.cmaps.setdefault(nxt.stride, nxt.coords)
            t = nxt

    def _fused_refs(self):
        """Module references of the inference fast path, resolved once (no getattr / Sequential
        indexing / nn.Module.__call__ on the per-scan hot path)."""
        refs = self.__dict__.get('_lk_refs')
        if refs is None:
            lv_refs = []
            for lv in (1, 2, 3, 4):
                d = getattr(self, f'down{lv}')[0].net
                stage = [(rb.net[0], rb.net[1], rb.net[3], rb.net[4]) for rb in getattr(self, f'stage{lv}')]
                tl, et = getattr(self, f'stage{lv}_tail'), getattr(self, f'elk{lv}_tail')
                lv_refs.append(((d[0], d[1]), stage, (tl[0], tl[1]), getattr(self, f'elk{lv}'),
                                (et[0], et[1])))
            st = self.stem
            refs = ((st[0], st[1], st[3], st[4]), lv_refs)
            self.__dict__['_lk_refs'] = refs
        return refs

    def _forward_levels_fused(self, x: SparseTensor):
        """forward_levels for inference: every Conv3d -> BatchNorm(eval) [-> + shortcut] [-> ReLU]
        group is one fused sparse-conv launch, every LinK block one native-executor call.

        The two branches of a level -- the conv stage and the LinK block, both functions of x_in
        (linkencoder.py:346-349) -- are enqueued on two streams: the stage's first conv builds the
        level's 3^3 kernel map and tile plan on the main stream, then the block runs on a side stream
        while the main stream continues with the rest of the stage; the tail conv that merges them
        waits for the block.  On the coarse levels (<= 28k voxels) neither branch fills the 148 SMs
        (a sparse conv there is one 128-row tile per CTA on 29-84 CTAs, ~26 us of latency), so the
        branches overlap instead of queueing.  Off by default on this per-layer path (see
        PY_BRANCH_OVERLAP); the native executor does it from C."""
        s, r = self.kwargs.get('s'), self.kwargs.get('r')
        cba = F.conv_bn_act
        (c0, b0, c1, b1), lv_refs = self._fused_refs()
        x0 = cba(cba(x, c0, b0, True), c1, b1, True)
        feats = [x0]
        cur = x0
        overlap = PY_BRANCH_OVERLAP
        if overlap:
            main = torch.cuda.current_stream()
            side = _branch_stream(x0.feats.device)
        for (dc, db), stage, (tc_, tb), elk_mod, (ec, eb) in lv_refs:
            x_in = cba(cur, dc, db, True)
            y = x_in
            join = None
            for i, (ca, ba, cb, bb) in enumerate(stage):
                y1 = cba(y, ca, ba, True)
                if i == 0 and overlap:
                    # the level's kernel map + plan are now enqueued (main); the block reads them and x_in
                    fork = torch.cuda.Event()
                    fork.record(main)
                    x_br = SparseTensor(x_in.feats, x_in.coords, x_in.stride)
                    x_br.cmaps, x_br.kmaps = x_in.cmaps, x_in.kmaps
                    with torch.cuda.stream(side):
                        side.wait_event(fork)
                        x_lk = elk_mod.forward(x_br, x_in.stride[0] * s, r)
                        join = torch.cuda.Event()
                        join.record(side)
                    x_lk.feats.record_stream(main)      # allocated on the side stream, consumed on main
                y = cba(y1, cb, bb, True, y.feats)
            x_conv = cba(y, tc_, tb, False)
            if join is not None:
                main.wait_event(join)
            else:
                x_lk = elk_mod.forward(x_in, x_in.stride[0] * s, r)
            cur = cba(x_lk, ec, eb, True, x_conv.feats)
            feats.append(cur)
        return feats

    def forward_levels(self, x: SparseTensor):
        """Stem + the four (conv stage || LinK block) levels; returns [x0, x1, x2, x3, x4]."""
        s, r = self.kwargs.get('s'), self.kwargs.get('r')
        x.cmaps.setdefault(x.stride, x.coords)
        self.plan_levels(x)
        if (F.fusable(self.stem[0], self.stem[1], x) and not self.training
                and self.stem[0].in_channels == x.feats.shape[1]
                and all(len(rb.downsample) == 0 for lv in (1, 2, 3, 4)
                        for rb in getattr(self, f'stage{lv}'))):
            return self._forward_levels_fused(x)
        if F.fusable(self.stem[0], self.stem[1], x):
            x0 = F.conv_bn_act(F.conv_bn_act(x, self.stem[0], self.stem[1], relu=True),
                               self.stem[3], self.stem[4], relu=True)
        else:
            x0 = _conv_bn_act(_conv_bn_act(x, self.stem[0], self.stem[1], True), self.stem[3], self.stem[4], True)
        feats = [x0]
        cur = x0
        for lv in (1, 2, 3, 4):
            x_in = getattr(self, f'down{lv}')(cur)
            x_conv = getattr(self, f'stage{lv}_tail')(getattr(self, f'stage{lv}')(x_in))
            # NB: the block mutates x_in (it then carries the LinK output); the conv stage above
            # has already consumed it -- same ordering as linkencoder.py:348-349.
            # relu(x_conv + BN(conv(x_lk))) -- the merge is folded into the tail conv's epilogue
            x_lk = getattr(self, f'elk{lv}')(x_in, x_in.s[0] * s, r)
            x_conv = getattr(self, f'elk{lv}_tail')(x_lk, residual=x_conv.F.contiguous(), relu=True)
            feats.append(x_conv)
            cur = x_conv
        return feats


class ELKEncoder(_ELKBackbone):

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        cs = self.cs
        self.classifier = nn.Sequential(
            nn.Conv1d(in_channels=cs[8] * 5, out_channels=120, kernel_size=1, groups=5),
            nn.ReLU(True),
            nn.Conv1d(in_channels=120, out_channels=kwargs['num_classes'], kernel_size=1, groups=1))
        self.weight_initialization()

    def forward(self, x: SparseTensor) -> torch.Tensor:
        if self._native_ok(x):
            x0, x1, x2, x3, x4 = self._forward_levels_native(x)
            return self._classify_pushdown([x4, x3, x2, x1, x0])
        x0, x1, x2, x3, x4 = self.forward_levels(x)
        if not torch.is_grad_enabled() and x0.F.dtype == torch.float32 and x0.F.is_cuda:
            return self._classify_pushdown([x4, x3, x2, x1, x0])
        return self._classify_levels([x4, x3, x2, x1, x0])

    # ---- native executor (lk_elk_encoder_fwd): the whole backbone behind one library call ----

    def _native_ok(self, x: SparseTensor) -> bool:
        """Inference on fp32 CUDA rows, BatchNorm in eval mode, every layer at a tensor-core width
        (32 / 64 / 128 channels; the stem's input is zero padded), identity shortcuts, a LinK block the
        block executor serves, no per-kernel timers."""
        import link_b200.elk as elk_mod
        from link_b200.nn.functional import conv as cv
        f = x._feats                      # (not x.feats: that would join a pending upload here)
        if not (NATIVE_ENCODER and not torch.is_grad_enabled() and not self.training and f.is_cuda
                and f.dtype == torch.float32 and _capi.TIMERS is None and elk_mod.NATIVE_EXECUTOR
                and cv.USE_TENSOR_CORES and cv.USE_PLAN and f.shape[0] > 0
                and tuple(x.stride) == (1, 1, 1) and f.shape[1] == self.stem[0].in_channels):
            return False
        hit = self.__dict__.get('_lk_native_static')
        if hit is None:
            widths = set(self.cs[:5])
            elks = [getattr(self, f'elk{lv}') for lv in (1, 2, 3, 4)]
            ok = (all(w in (32, 64, 128) for w in widths) and self.stem[0].in_channels <= 32
                  and all(len(rb.downsample) == 0 for lv in (1, 2, 3, 4) for rb in getattr(self, f'stage{lv}'))
                  and all(e.baseop in ('cos', 'sin') or e.groups == 1 for e in elks)
                  and all(m.bias is None for m in self.modules() if isinstance(m, spnn.Conv3d))
                  and self.kwargs.get('r') in (2, 3))
            # the BatchNorms on the executor's path (the decoder branches up1..up4 are never run)
            (c0, b0, c1, b1), lv_refs = self._fused_refs()
            bns = [b0, b1]
            for (dc, db), stage, (tc_, tb), elk, (ec, eb) in lv_refs:
                bns += [db, tb, eb] + [b for ca, ba, cb, bb in stage for b in (ba, bb)]
            hit = self.__dict__['_lk_native_static'] = (ok, bns)
        return hit[0] and all(not m.training and m.track_running_stats and m.affine for m in hit[1])

    def _native_template(self, dev):
        """lk_elk_encoder_args_t with every parameter-only field filled (packed weight images, folded
        BatchNorm affines, the four block templates), cached until a parameter changes."""
        import link_b200.elk as elk_mod
        from link_b200.nn.functional import conv as cv
        (c0, b0, c1, b1), lv_refs = self._fused_refs()
        convs = [(c0, b0), (c1, b1)]
        for (dc, db), stage, (tc_, tb), elk, (ec, eb) in lv_refs:
            convs += [(dc, db)] + [p for ca, ba, cb, bb in stage for p in ((ca, ba), (cb, bb))] + [(tc_, tb), (ec, eb)]
        ver = tuple((c.kernel._version, c.kernel.data_ptr(), b.weight._version, b.bias._version,
                     b.running_mean._version, b.running_var._version, b.running_mean.data_ptr()) for c, b in convs)
        elk_params = self.__dict__.get('_lk_elk_params')
        if elk_params is None:
            elk_params = self.__dict__['_lk_elk_params'] = [p for _, _, _, elk, _ in lv_refs for p in elk.parameters()]
        ver += tuple(p._version for p in elk_params)
        ver += (cv.precision_code(), elk_mod.SINGLE_STREAM, elk_mod.ACCURATE_TRIG, str(dev))
        hit = self.__dict__.get('_lk_enc_native')
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        keep = []

        def layer(dst, conv, bn, relu, pad_in=0):
            img = cv._tc_image(conv.kernel, pad_in, 0)
            scale, shift = cv._folded_bn(bn)
            keep.extend((img, scale, shift))
            dst.d_wimg, dst.d_scale, dst.d_shift = _capi.ptr(img), _capi.ptr(scale), _capi.ptr(shift)
            dst.c_in, dst.c_out, dst.relu = max(pad_in, conv.kernel.shape[1]), conv.kernel.shape[2], 1 if relu else 0

        a = _capi.ElkEncoderArgs()
        a.levels = 4
        a.c_max = max(self.cs[:5] + [32])
        a.conv_precision = cv.precision_code()
        a.single_stream = 1 if elk_mod.SINGLE_STREAM else 0
        a.overlap_branches = 1 if BRANCH_OVERLAP else 0
        layer(a.stem[0], c0, b0, True, pad_in=cv._pad_to_tc(c0.kernel.shape[1]))
        layer(a.stem[1], c1, b1, True)
        r = self.kwargs.get('r')
        for i, ((dc, db), stage, (tc_, tb), elk, (ec, eb)) in enumerate(lv_refs):
            L = a.level[i]
            layer(L.down, dc, db, True)
            for j, (ca, ba, cb, bb) in enumerate(stage):
                layer(L.stage[2 * j], ca, ba, True)
                layer(L.stage[2 * j + 1], cb, bb, True)
            layer(L.tail, tc_, tb, False)
            layer(L.elk_tail, ec, eb, True)
            scale = float(2 ** (i + 1)) if (elk.baseop == 'cos_x' and elk.variant == 'encoder') else 1.0
            t = elk_mod._native_template(elk.baseop, elk.inc, elk.pre_mix, elk.local_mix[0], elk.pos_weight[0].weight,
                                         getattr(elk, 'alpha', None), scale, elk.norm, elk.norm_local, dev)
            keep.append(t)
            C.memmove(C.byref(L.elk), C.byref(t[1]), C.sizeof(_capi.ElkBlockArgs))
            L.elk.r3 = r ** 3
        self.__dict__['_lk_enc_native'] = (ver, a, keep)
        return a, keep

    def _forward_levels_native(self, x: SparseTensor):
        """[x0 .. x4] through lk_elk_encoder_fwd: the level outputs and coordinates are caller-owned
        buffers of n0 rows (a strided level never outgrows its input), sliced to the sizes the
        executor reports."""
        from link_b200.nn.functional import _index
        from link_b200.nn.utils import get_kernel_offsets
        L = _capi.lib()
        feats, ready = x.take_feats_event()
        _capi.check_device(feats)
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)
        dev = feats.device
        tmpl, keep = self._native_template(dev)
        a = _capi.ElkEncoderArgs.from_buffer_copy(tmpl)
        coords = x.coords.contiguous()
        n0 = coords.shape[0]
        ci = a.stem[0].c_in
        feats = feats.contiguous()
        if feats.shape[1] != ci:
            feats = torch.nn.functional.pad(feats, (0, ci - feats.shape[1]))
        x.cmaps.setdefault(x.stride, x.coords)
        bounds = _index.coord_bounds(coords, x.kmaps)
        s, r = self.kwargs.get('s'), self.kwargs.get('r')
        blk_off = get_kernel_offsets(r, 1, 1, device=dev)
        off3 = [get_kernel_offsets(3, stride=(2 ** l,) * 3, device=dev) for l in range(5)]
        a.n0, a.d_coords0, a.d_feats0 = n0, _capi.ptr(coords, torch.int32), _capi.ptr(feats, torch.float32)
        a.d_off3_0 = _capi.ptr(off3[0])
        outs = [torch.empty(n0, self.cs[l], dtype=torch.float32, device=dev) for l in range(5)]
        lv_coords = [coords] + [torch.empty(n0, 4, dtype=torch.int32, device=dev) for _ in range(4)]
        a.d_out0 = _capi.ptr(outs[0])
        keep_off = [blk_off, off3]
        b = bounds
        for l in range(1, 5):
            Lv = a.level[l - 1]
            ss = (2 ** l,) * 3
            # sites of level l straight from the input coordinates: floor(c0 / 2^l) 2^l, order (b, x, y, z)
            Lv.down_spec, Lv.down_bits = _index.make_keyspec(bounds, ss, (3, 0, 1, 2), ss)
            off2 = get_kernel_offsets(2, stride=(2 ** (l - 1),) * 3, device=dev)
            keep_off.append(off2)
            Lv.d_off2, Lv.d_off3 = _capi.ptr(off2), _capi.ptr(off3[l])
            Lv.d_out, Lv.d_coords = _capi.ptr(outs[l]), _capi.ptr(lv_coords[l])
            st4 = ss + (1,)
            b = (tuple((b[0][k] // st4[k]) * st4[k] for k in range(4)), tuple((b[1][k] // st4[k]) * st4[k] for k in range(4)))
            se = 2 ** l * s
            Lv.elk.keyspec, Lv.elk.key_bits = _index.make_keyspec(b, (se, se, se), (0, 1, 2, 3))
            Lv.elk.d_block_offsets = _capi.ptr(blk_off)
        ws_bytes = L.lk_elk_encoder_ws_bytes(n0, 4, a.c_max, a.level[0].elk.gen.op, r ** 3)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        a.d_ws, a.ws_bytes = _capi.ptr(ws), ws_bytes
        _capi.check(L.lk_elk_encoder_fwd(C.byref(a), _capi.stream()), 'lk_elk_encoder_fwd')
        levels = []
        for l in range(5):
            n_l = int(a.n_out[l])
            t = SparseTensor(outs[l][:n_l], lv_coords[l] if l == 0 else lv_coords[l][:n_l], (2 ** l,) * 3)
            t.cmaps, t.kmaps = x.cmaps, x.kmaps
            t.cmaps.setdefault(t.stride, t.coords)
            levels.append(t)
        return levels

    def _classify_levels(self, levels) -> torch.Tensor:
        """Differentiable head (training).  Same push-down as the inference head: group l of the
        grouped 1x1 conv (linkencoder.py:323-327, 376-379) is applied at level l's own size as a plain
        fp32 matmul and the 24-channel result is gathered to the N0 voxels (`index_select`, whose
        backward is an `index_add`), instead of gathering five 64-channel tensors, concatenating them
        to [N0, 320] and running the grouped conv there -- as an einsum that was a strided batched
        SIMT GEMM of 3.7 ms per backward product on 160k voxels."""
        c0, c2 = self.classifier[0], self.classifier[2]
        g = c0.groups
        w0 = c0.weight.view(g, c0.out_channels // g, -1)                 # [5, 24, C]
        x0 = levels[-1]
        hs = []
        for i, lv in enumerate(levels):
            z = torch.mm(lv.F, w0[i].t())
            if lv is not x0:
                z = z.index_select(0, upsample_index(lv, x0))
            hs.append(z)
        h = torch.relu(torch.cat(hs, dim=1) + c0.bias)
        return torch.addmm(c2.bias, h, c2.weight.view(c2.out_channels, -1).t())

    def _classify_pushdown(self, levels) -> torch.Tensor:
        """Inference head.  The reference upsamples every level to N0 rows x 64 channels,
        concatenates them ([N0, 320]) and applies the grouped 1x1 conv (one group per level).  A
        1x1 conv commutes with a row gather, so the group of level l is applied at that level's own
        (coarse) size, and one kernel gathers the five 24-channel results, adds the bias and
        applies the ReLU: same values, ~5x less gather traffic, no [N0, 320] temporary."""
        c0, c2 = self.classifier[0], self.classifier[2]
        g = c0.groups
        oc = c0.out_channels // g
        w0 = c0.weight.view(g, oc, -1)
        x0 = levels[-1]
        n0 = x0.F.shape[0]
        z = [torch.mm(lv.F, w0[i].t()) for i, lv in enumerate(levels)]        # [N_l, 24] each
        idx = [upsample_index(lv, x0) for lv in levels[:-1]] + [None]
        h = torch.empty(n0, g * oc, dtype=torch.float32, device=x0.F.device)
        src_arr = (C.c_void_p * g)(*[t.data_ptr() for t in z])
        idx_arr = (C.c_void_p * g)(*[(t.data_ptr() if t is not None else None) for t in idx])
        _capi.check(_capi.lib().lk_gather_concat(src_arr, idx_arr, g, oc, n0,
                                                 _capi.ptr(c0.bias.detach()), 1, _capi.ptr(h),
                                                 _capi.stream()), 'lk_gather_concat')
        return torch.addmm(c2.bias, h, c2.weight.view(c2.out_channels, -1).t())

    def _classify(self, f_cat: torch.Tensor) -> torch.Tensor:
        """The reference's grouped 1x1 Conv1d head (linkencoder.py:323-327, 376-379) evaluated as
        plain fp32 matmuls on the same parameters: identical math, but cuDNN's default TF32
        convolution path (1e-3 relative error) is not involved."""
        c0, c2 = self.classifier[0], self.classifier[2]
        n = f_cat.shape[0]
        g = c0.groups
        w0 = c0.weight.view(g, c0.out_channels // g, -1)                 # [5, 24, C]
        h = torch.einsum('ngc,goc->ngo', f_cat.view(n, g, -1), w0).reshape(n, -1) + c0.bias
        h = torch.relu(h)
        return torch.addmm(c2.bias, h, c2.weight.view(c2.out_channels, -1).t())


LinKEncoder = ELKEncoder
