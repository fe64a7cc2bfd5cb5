"""CenterPoint head stack behind the LinK detection backbone (BASELINE config 4):

    points/voxels -> VoxelFeatureExtractorV3 -> SpMiddleResNetFHDELKv3 (link_b200/scn.py)
                  -> RPN (dense BEV neck) -> CenterHead (+ FastFocalLoss / RegLoss)

The dense part is library code by design (cuDNN convolutions through plain `torch.nn`): the custom
CUDA of this path lives in the sparse backbone and the LinK blocks.  Classes keep the reference's
names, constructor arguments, forward signatures and state-dict keys, so a det3d checkpoint loads
with `strict=True`:

  * VoxelFeatureExtractorV3  detection/det3d/models/readers/voxel_encoder.py:9-24
  * RPN                      detection/det3d/models/necks/rpn.py:22-159
  * SepHead, CenterHead      detection/det3d/models/bbox_heads/center_head.py:67-293
  * FastFocalLoss, RegLoss   detection/det3d/models/losses/centernet_loss.py:6-54
  * VoxelNet                 detection/det3d/models/detectors/voxelnet.py:10-66 (extract_feat / forward)
  * CenterHead.predict, circle_nms   center_head.py:296-515, det3d/core/utils/circle_nms_jit.py:4-28

Parity: tests/test_centerpoint_cpu.py loads reference-generated weights + outputs
(tests/golden/centerpoint.npz, made by tests/golden/make_centerpoint_golden.py from the unmodified
reference classes) and compares forward outputs, losses, the input gradient and the decoded /
NMS-filtered detections.
"""
import copy
import math
from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as tF
from torch import nn

from link_b200.iou3d import rotate_nms_pcdet

__all__ = ['VoxelFeatureExtractorV3', 'RPN', 'SepHead', 'CenterHead', 'FastFocalLoss', 'RegLoss', 'circle_nms',
           'VoxelNet', 'NUSC_TASKS', 'NUSC_COMMON_HEADS', 'NUSC_CODE_WEIGHTS', 'NUSC_TEST_CFG',
           'build_nusc_centerpoint']


class VoxelFeatureExtractorV3(nn.Module):
    """Mean of the points of every voxel: `features [Nv, P, C]` (zero padded past `num_voxels[i]`
    points) -> `[Nv, C]`."""

    def __init__(self, num_input_features=4, norm_cfg=None, name='VoxelFeatureExtractorV3'):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors=None):
        assert self.num_input_features == features.shape[-1]
        total = features[:, :, :self.num_input_features].sum(dim=1)
        return (total / num_voxels.type_as(features).view(-1, 1)).contiguous()


def _bn2d(channels: int, norm_cfg: dict) -> nn.BatchNorm2d:
    cfg = dict(norm_cfg)
    kind = cfg.pop('type', 'BN')
    if kind != 'BN':
        raise KeyError(f'Unrecognized norm type {kind}')
    trainable = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    bn = nn.BatchNorm2d(channels, **cfg)
    for p in bn.parameters():
        p.requires_grad = trainable
    return bn


class RPN(nn.Module):
    """Dense BEV neck: per level a strided 3x3 conv + `layer_nums[i]` 3x3 convs (each conv-BN-ReLU),
    every level from `len(layer_nums) - len(us_layer_strides)` on is brought to a common resolution
    (transposed conv for stride > 1, strided conv for stride < 1) and the results are concatenated."""

    def __init__(self, layer_nums, ds_layer_strides, ds_num_filters, us_layer_strides, us_num_filters,
                 num_input_features, norm_cfg=None, name='rpn', logger=None, **kwargs):
        super().__init__()
        self._layer_nums = list(layer_nums)
        self._layer_strides = list(ds_layer_strides)
        self._num_filters = list(ds_num_filters)
        self._upsample_strides = list(us_layer_strides)
        self._num_upsample_filters = list(us_num_filters)
        self._num_input_features = num_input_features
        self._norm_cfg = norm_cfg if norm_cfg is not None else dict(type='BN', eps=1e-3, momentum=0.01)
        assert len(self._layer_strides) == len(self._layer_nums) == len(self._num_filters)
        assert len(self._num_upsample_filters) == len(self._upsample_strides)
        self._upsample_start_idx = first = len(self._layer_nums) - len(self._upsample_strides)
        # every upsampled level must land on the same resolution
        ratios = [self._upsample_strides[k] / math.prod(self._layer_strides[:k + first + 1])
                  for k in range(len(self._upsample_strides))]
        assert all(r == ratios[0] for r in ratios)

        blocks, deblocks = [], []
        c_in = num_input_features
        for lvl, (depth, stride, c_out) in enumerate(zip(self._layer_nums, self._layer_strides, self._num_filters)):
            layers = [nn.ZeroPad2d(1), nn.Conv2d(c_in, c_out, 3, stride=stride, bias=False),
                      _bn2d(c_out, self._norm_cfg), nn.ReLU()]
            for _ in range(depth):
                layers += [nn.Conv2d(c_out, c_out, 3, padding=1, bias=False), _bn2d(c_out, self._norm_cfg), nn.ReLU()]
            blocks.append(nn.Sequential(*layers))
            if lvl >= first:
                up, c_up = self._upsample_strides[lvl - first], self._num_upsample_filters[lvl - first]
                if up > 1:
                    resample = nn.ConvTranspose2d(c_out, c_up, up, stride=up, bias=False)
                else:
                    down = int(round(1 / up))
                    resample = nn.Conv2d(c_out, c_up, down, stride=down, bias=False)
                deblocks.append(nn.Sequential(resample, _bn2d(c_up, self._norm_cfg), nn.ReLU()))
            c_in = c_out
        self.blocks = nn.ModuleList(blocks)
        self.deblocks = nn.ModuleList(deblocks)

    @property
    def downsample_factor(self):
        factor = math.prod(self._layer_strides)
        if self._upsample_strides:
            factor /= self._upsample_strides[-1]
        return factor

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        ups = []
        for lvl, block in enumerate(self.blocks):
            x = block(x)
            if lvl >= self._upsample_start_idx:
                ups.append(self.deblocks[lvl - self._upsample_start_idx](x))
        return torch.cat(ups, dim=1) if ups else x


class SepHead(nn.Module):
    """One small conv tower per regression target: `heads = {name: (out_channels, num_conv)}`."""

    def __init__(self, in_channels, heads, head_conv=64, final_kernel=1, bn=False, init_bias=-2.19, **kwargs):
        super().__init__(**kwargs)
        self.heads = heads
        pad = final_kernel // 2
        for name, (out_channels, num_conv) in heads.items():
            layers = []
            for _ in range(num_conv - 1):
                layers.append(nn.Conv2d(in_channels, head_conv, final_kernel, stride=1, padding=pad, bias=True))
                if bn:
                    layers.append(nn.BatchNorm2d(head_conv))
                layers.append(nn.ReLU())
            layers.append(nn.Conv2d(head_conv, out_channels, final_kernel, stride=1, padding=pad, bias=True))
            tower = nn.Sequential(*layers)
            if 'hm' in name:
                tower[-1].bias.data.fill_(init_bias)      # heat-map logits start near sigmoid^-1(0.1)
            else:
                for m in tower.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.kaiming_normal_(m.weight, a=0, mode='fan_out', nonlinearity='relu')
                        nn.init.zeros_(m.bias)
            setattr(self, name, tower)

    def forward(self, x):
        return {name: getattr(self, name)(x) for name in self.heads}


def _gather_at(feat: torch.Tensor, ind: torch.Tensor) -> torch.Tensor:
    """`feat [B, C, H, W]`, flat BEV cell indices `ind [B, M]` -> `[B, M, C]`
    (det3d/core/utils/center_utils.py:65-79)."""
    b, c = feat.shape[:2]
    flat = feat.reshape(b, c, -1)
    return flat.gather(2, ind.unsqueeze(1).expand(b, c, ind.shape[1])).transpose(1, 2)


class RegLoss(nn.Module):
    """Masked L1 at the object centres, summed over batch and objects, one value per box code."""

    def forward(self, output, mask, ind, target):
        pred = _gather_at(output, ind)
        m = mask.float().unsqueeze(2)
        loss = tF.l1_loss(pred * m, target * m, reduction='none') / (m.sum() + 1e-4)
        return loss.sum(dim=(0, 1))


class FastFocalLoss(nn.Module):
    """CornerNet focal loss with the positive term evaluated only at the annotated peaks."""

    def forward(self, out, target, ind, mask, cat):
        mask = mask.float()
        neg = (torch.log(1 - out) * out.pow(2) * (1 - target).pow(4)).sum()
        peak = _gather_at(out, ind).gather(2, cat.unsqueeze(2))                 # [B, M, 1]
        pos = (torch.log(peak) * (1 - peak).pow(2) * mask.unsqueeze(2)).sum()
        n_pos = mask.sum()
        if n_pos == 0:
            return -neg
        return -(pos + neg) / n_pos


class _Cfg:
    """Attribute + `.get` access over a plain dict or a det3d Config (test_cfg is read both ways)."""

    def __init__(self, obj):
        self._o = obj

    def get(self, key, default=None):
        o = self._o
        if isinstance(o, dict):
            v = o.get(key, default)
        else:
            v = getattr(o, key, default)
        return _Cfg(v) if isinstance(v, dict) else v

    def __getattr__(self, key):
        v = self.get(key, _Cfg)
        if v is _Cfg:
            raise AttributeError(key)
        return v


def circle_nms(centers: torch.Tensor, scores: torch.Tensor, thresh: float, post_max_size: int = 83,
               block: int = 2048) -> torch.Tensor:
    """Greedy NMS by squared centre distance (det3d/core/utils/circle_nms_jit.py:4-28: a box is
    dropped when a kept, higher-scoring box lies within `thresh` in SQUARED distance), on the
    device without a sequential loop over boxes: with boxes sorted by score, `keep[j] = not
    any_{i<j}(keep[i] and close[i, j])` has a unique solution; starting from keep = all and
    re-evaluating the right-hand side reaches it after as many sweeps as the longest suppression
    chain (each sweep is one masked reduction over the pairwise matrix, built in row blocks).
    Returns indices into the input, highest score first, at most `post_max_size`."""
    n = centers.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.long, device=centers.device)
    order = torch.argsort(scores, descending=True)
    c = centers[order]
    if c.is_cuda and n <= 65536:
        # device kernels (csrc/iou3d.cu: distance bitmasks + a one-warp greedy scan); the masked-reduction
        # form below needs one host check per sweep (740 ms on a 180 x 180 head map against 22 ms)
        from link_b200 import _capi
        c2 = c[:, :2].float().contiguous()
        L = _capi.lib()
        ws = torch.empty(L.lk_nms_bev_ws_bytes(n) // 8 + 1, dtype=torch.int64, device=c.device)
        keep8 = torch.empty(n, dtype=torch.uint8, device=c.device)
        _capi.check(L.lk_nms_circle(_capi.ptr(c2), n, float(thresh), _capi.ptr(ws), ws.numel() * 8, _capi.ptr(keep8),
                                    _capi.stream()), 'lk_nms_circle')
        return order[keep8.bool()][:post_max_size]
    keep = torch.ones(n, dtype=torch.bool, device=centers.device)
    rank = torch.arange(n, device=centers.device)
    for _ in range(n):
        supp = torch.zeros_like(keep)
        for r0 in range(0, n, block):
            d = c[r0:r0 + block, None, :] - c[None, :, :]
            close = (d[..., 0] ** 2 + d[..., 1] ** 2) <= thresh                  # [rows, n]
            earlier = rank[r0:r0 + block, None] < rank[None, :]
            supp |= (close & earlier & keep[r0:r0 + block, None]).any(dim=0)
        new_keep = ~supp
        if bool((new_keep == keep).all()):
            break
        keep = new_keep
    return order[keep][:post_max_size]


class CenterHead(nn.Module):
    """Shared 3x3 conv + one SepHead per task group (heat map + reg / height / dim / rot / vel)."""

    def __init__(self, in_channels=(128,), tasks=(), dataset='nuscenes', weight=0.25, code_weights=(),
                 common_heads=None, logger=None, init_bias=-2.19, share_conv_channel=64, num_hm_conv=2,
                 dcn_head=False):
        super().__init__()
        if dcn_head:
            raise NotImplementedError('dcn_head=True (deformable conv heads) is outside the hot path')
        common_heads = dict(common_heads or {})
        self.num_classes = [len(t['class_names']) for t in tasks]
        self.class_names = [t['class_names'] for t in tasks]
        self.code_weights = list(code_weights)
        self.weight = weight
        self.dataset = dataset
        self.in_channels = in_channels
        self.crit = FastFocalLoss()
        self.crit_reg = RegLoss()
        self.box_n_dim = 9 if 'vel' in common_heads else 7
        self.use_direction_classifier = False
        self.shared_conv = nn.Sequential(
            nn.Conv2d(in_channels, share_conv_channel, kernel_size=3, padding=1, bias=True),
            nn.BatchNorm2d(share_conv_channel), nn.ReLU(inplace=True))
        self.tasks = nn.ModuleList()
        for n_cls in self.num_classes:
            heads = copy.deepcopy(common_heads)
            heads.update(dict(hm=(n_cls, num_hm_conv)))
            self.tasks.append(SepHead(share_conv_channel, heads, bn=True, init_bias=init_bias, final_kernel=3))

    def forward(self, x, *kwargs):
        x = self.shared_conv(x)
        return [task(x) for task in self.tasks], x

    @staticmethod
    def _sigmoid(x):
        return torch.clamp(x.sigmoid_(), min=1e-4, max=1 - 1e-4)

    def loss(self, example, preds_dicts, test_cfg=None, **kwargs):
        if self.dataset not in ('waymo', 'nuscenes'):
            raise NotImplementedError()
        merged = defaultdict(list)
        for t, preds in enumerate(preds_dicts):
            preds['hm'] = self._sigmoid(preds['hm'])
            hm_loss = self.crit(preds['hm'], example['hm'][t], example['ind'][t], example['mask'][t], example['cat'][t])
            target_box = example['anno_box'][t]
            if 'vel' in preds:
                parts = ('reg', 'height', 'dim', 'vel', 'rot')
            else:
                parts = ('reg', 'height', 'dim', 'rot')
                target_box = target_box[..., [0, 1, 2, 3, 4, 5, -2, -1]]        # drop the velocity target
            preds['anno_box'] = torch.cat([preds[k] for k in parts], dim=1)
            box_loss = self.crit_reg(preds['anno_box'], example['mask'][t], example['ind'][t], target_box)
            loc_loss = (box_loss * box_loss.new_tensor(self.code_weights)).sum()
            ret = {'loss': hm_loss + self.weight * loc_loss, 'hm_loss': hm_loss.detach().cpu(),
                   'loc_loss': loc_loss, 'loc_loss_elem': box_loss.detach().cpu(),
                   'num_positive': example['mask'][t].float().sum()}
            for k, v in ret.items():
                merged[k].append(v)
        return merged

    @torch.no_grad()
    def predict(self, example, preds_dicts, test_cfg, **kwargs):
        """Decode the head outputs into boxes `(x, y, z, w, l, h, [vx, vy], yaw)`, threshold, NMS, merge
        the task groups (center_head.py:296-515).  `test_cfg.circular_nms=True` uses the on-device
        `circle_nms`, otherwise the rotated-IoU NMS of link_b200/iou3d.py (`rotate_nms_pcdet`)."""
        cfg = _Cfg(test_cfg)
        double_flip = cfg.get('double_flip', False)
        hm0 = preds_dicts[0]['hm']
        center_range = cfg.post_center_limit_range
        if len(center_range) > 0:
            center_range = torch.tensor(center_range, dtype=hm0.dtype, device=hm0.device)
        rets, metas = [], []
        for task_id, preds in enumerate(preds_dicts):
            p = {k: v.permute(0, 2, 3, 1).contiguous() for k, v in preds.items()}          # N H W C
            batch = p['hm'].shape[0]
            if double_flip:
                # groups of 4: original, y -> -y, x -> -x, both; bring the maps back to the original frame
                assert batch % 4 == 0, batch
                batch //= 4
                for k, v in p.items():
                    _, h, w, ch = v.shape
                    v = v.reshape(batch, 4, h, w, ch).clone()
                    v[:, 1] = torch.flip(v[:, 1], dims=[1])
                    v[:, 2] = torch.flip(v[:, 2], dims=[2])
                    v[:, 3] = torch.flip(v[:, 3], dims=[1, 2])
                    p[k] = v
            if 'metadata' not in example or len(example['metadata']) == 0:
                meta = [None] * batch
            else:
                meta = example['metadata'][:4 * batch:4] if double_flip else example['metadata']
            hm, dim = torch.sigmoid(p['hm']), torch.exp(p['dim'])
            rot_s, rot_c = p['rot'][..., 0:1].clone(), p['rot'][..., 1:2].clone()
            reg, hei = p['reg'].clone(), p['height']
            vel = p['vel'].clone() if 'vel' in p else None
            if double_flip:
                hm, hei, dim = hm.mean(dim=1), hei.mean(dim=1), dim.mean(dim=1)
                reg[:, 1, ..., 1] = 1 - reg[:, 1, ..., 1]                 # y -> -y: offset_y -> 1 - offset_y
                reg[:, 2, ..., 0] = 1 - reg[:, 2, ..., 0]
                reg[:, 3, ..., 0] = 1 - reg[:, 3, ..., 0]
                reg[:, 3, ..., 1] = 1 - reg[:, 3, ..., 1]
                reg = reg.mean(dim=1)
                rot_c[:, 1] *= -1                                          # theta -> pi - theta
                rot_s[:, 2] *= -1                                          # theta -> -theta
                rot_s[:, 3] *= -1
                rot_c[:, 3] *= -1
                rot_c, rot_s = rot_c.mean(dim=1), rot_s.mean(dim=1)
                if vel is not None:
                    vel[:, 1, ..., 1] *= -1
                    vel[:, 2, ..., 0] *= -1
                    vel[:, 3] *= -1
                    vel = vel.mean(dim=1)
            rot = torch.atan2(rot_s, rot_c)
            b, h, w, n_cls = hm.shape
            ys, xs = torch.meshgrid(torch.arange(h, device=hm.device), torch.arange(w, device=hm.device), indexing='ij')
            xs = xs.reshape(1, -1, 1).to(hm) + reg.reshape(b, h * w, 2)[:, :, 0:1]
            ys = ys.reshape(1, -1, 1).to(hm) + reg.reshape(b, h * w, 2)[:, :, 1:2]
            xs = xs * cfg.out_size_factor * cfg.voxel_size[0] + cfg.pc_range[0]
            ys = ys * cfg.out_size_factor * cfg.voxel_size[1] + cfg.pc_range[1]
            parts = [xs, ys, hei.reshape(b, h * w, 1), dim.reshape(b, h * w, 3)]
            if vel is not None:
                parts.append(vel.reshape(b, h * w, 2))
            parts.append(rot.reshape(b, h * w, 1))
            metas.append(meta)
            if cfg.get('per_class_nms', False):
                raise NotImplementedError('per_class_nms')                 # the reference silently skips the task
            rets.append(self.post_processing(torch.cat(parts, dim=2), hm.reshape(b, h * w, n_cls), cfg,
                                             center_range, task_id))
        out = []
        for i in range(len(rets[0])):
            ret = {}
            for k in ('box3d_lidar', 'scores'):
                ret[k] = torch.cat([r[i][k] for r in rets])
            first = 0
            for j, n_cls in enumerate(self.num_classes):                  # task-local labels -> global class ids
                rets[j][i]['label_preds'] += first
                first += n_cls
            ret['label_preds'] = torch.cat([r[i]['label_preds'] for r in rets])
            ret['metadata'] = metas[0][i]
            out.append(ret)
        return out


    @torch.no_grad()
    def post_processing(self, batch_box_preds, batch_hm, test_cfg, post_center_range, task_id):
        cfg = test_cfg if isinstance(test_cfg, _Cfg) else _Cfg(test_cfg)
        if cfg.get('tt_rotation', 0) != 0:
            raise NotImplementedError('tt_rotation')
        results = []
        for box_preds, hm_preds in zip(batch_box_preds, batch_hm):
            scores, labels = torch.max(hm_preds, dim=-1)
            mask = scores > cfg.score_threshold
            mask &= (box_preds[..., :3] >= post_center_range[:3]).all(1) & (box_preds[..., :3] <= post_center_range[3:]).all(1)
            box_preds, scores, labels = box_preds[mask], scores[mask], labels[mask]
            if cfg.get('circular_nms', False):
                sel = circle_nms(box_preds[:, :2], scores, thresh=cfg.min_radius[task_id],
                                 post_max_size=cfg.nms.nms_post_max_size)
            else:
                sel = rotate_nms_pcdet(box_preds[:, [0, 1, 2, 3, 4, 5, -1]].float(), scores.float(),
                                       thresh=cfg.nms.nms_iou_threshold, pre_maxsize=cfg.nms.nms_pre_max_size,
                                       post_max_size=cfg.nms.nms_post_max_size)
            results.append({'box3d_lidar': box_preds[sel], 'scores': scores[sel], 'label_preds': labels[sel]})
        return results


class VoxelNet(nn.Module):
    """reader -> sparse backbone -> neck -> head, the single-stage CenterPoint detector of config 4.
    `example` carries `voxels [Nv, P, C]`, `num_points [Nv]`, `coordinates [Nv, 4] (b, z, y, x)`,
    `shape` (grid size x, y, z per sample) and, for the loss, the CenterPoint targets."""

    def __init__(self, reader, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__()
        self.reader, self.backbone, self.neck, self.bbox_head = reader, backbone, neck, bbox_head
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        if train_cfg is not None and train_cfg.get('freeze_bkbn', False):
            for p in self.backbone.parameters():
                p.requires_grad = False

    @property
    def with_neck(self):
        return self.neck is not None

    def extract_feat(self, data):
        feats = self.reader(data['voxels'], data['num_points'])
        batch_size = data['batch_size'] if 'batch_size' in data else len(data['points'])
        x, voxel_feature = self.backbone(feats, data['coordinates'], batch_size, data['shape'][0])
        if self.with_neck:
            x = self.neck(x)
        return x, voxel_feature

    def forward(self, example, return_loss=True, **kwargs):
        x, _ = self.extract_feat(example)
        preds, _ = self.bbox_head(x)
        if return_loss:
            return self.bbox_head.loss(example, preds, self.test_cfg)
        if self.test_cfg is None:                  # no post-processing configured: the raw head maps
            return preds
        return self.bbox_head.predict(example, preds, self.test_cfg)


# nuScenes task grouping and head layout of the reference config
# (detection/configs/nusc/voxelnet/nusc_centerpoint_voxelnet_0075voxel_fix_bn_z_elkv3.py:6-56)
NUSC_TASKS: List[Dict] = [
    dict(num_class=1, class_names=['car']),
    dict(num_class=2, class_names=['truck', 'construction_vehicle']),
    dict(num_class=2, class_names=['bus', 'trailer']),
    dict(num_class=1, class_names=['barrier']),
    dict(num_class=2, class_names=['motorcycle', 'bicycle']),
    dict(num_class=2, class_names=['pedestrian', 'traffic_cone']),
]
NUSC_COMMON_HEADS: Dict[str, Tuple[int, int]] = {'reg': (2, 2), 'height': (1, 2), 'dim': (3, 2), 'rot': (2, 2), 'vel': (2, 2)}
NUSC_CODE_WEIGHTS: Sequence[float] = (1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0)
# test_cfg of the same config (lines 68-82); out_size_factor = get_downsample_factor(model) = 8
NUSC_TEST_CFG: Dict = dict(
    post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], max_per_img=500,
    nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=83,
             nms_iou_threshold=0.2),
    score_threshold=0.1, pc_range=[-54, -54], out_size_factor=8, voxel_size=[0.075, 0.075])


def build_nusc_centerpoint(backbone=None, test_cfg=None) -> VoxelNet:
    """The model dict of the reference config as modules (random init).  Pass `test_cfg=NUSC_TEST_CFG`
    to get decoded, NMS-filtered detections from `forward(example, return_loss=False)`."""
    if backbone is None:
        from link_b200.scn import SpMiddleResNetFHDELKv3
        backbone = SpMiddleResNetFHDELKv3(num_input_features=5, ds_factor=8)
    return VoxelNet(
        reader=VoxelFeatureExtractorV3(num_input_features=5),
        backbone=backbone,
        neck=RPN(layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                 us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256),
        bbox_head=CenterHead(in_channels=512, tasks=NUSC_TASKS, dataset='nuscenes', weight=0.25,
                             code_weights=list(NUSC_CODE_WEIGHTS), common_heads=dict(NUSC_COMMON_HEADS),
                             share_conv_channel=64, dcn_head=False),
        test_cfg=test_cfg)
