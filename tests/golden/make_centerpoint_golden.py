"""Golden fixture for the CenterPoint head stack (tests/golden/centerpoint.npz): runs the reference's
own, unmodified classes

    det3d/models/readers/voxel_encoder.py   VoxelFeatureExtractorV3
    det3d/models/necks/rpn.py               RPN
    det3d/models/bbox_heads/center_head.py  CenterHead (SepHead) forward + loss
    det3d/models/losses/centernet_loss.py   FastFocalLoss, RegLoss

on seeded inputs with a SMALL configuration (so that the weights fit in the fixture) and stores the
reference state dicts, the inputs and the outputs.  The `det3d` package itself does not import here
(terminaltables, spconv, nuscenes-devkit ... are absent), so the source files are loaded by path
under their own module names, with empty stand-ins for the package modules they only reference
(registries, builder, checkpoint loader, NMS helpers).  Build container only."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
DET = '/root/reference/detection/det3d'


def _pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(DET, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    parent, _, child = name.rpartition('.')
    setattr(sys.modules[parent], child, m)
    spec.loader.exec_module(m)
    return m


class _Registry:
    def register_module(self, cls):
        return cls


def import_reference():
    for p in ('det3d', 'det3d.core', 'det3d.core.utils', 'det3d.core.bbox', 'det3d.torchie', 'det3d.models',
              'det3d.models.necks', 'det3d.models.bbox_heads', 'det3d.models.readers', 'det3d.models.losses',
              'det3d.utils', 'det3d.utils.dist'):
        _pkg(p)
    # modules the loaded files only reference
    for name in ('det3d.core.box_torch_ops', 'det3d.core.bbox.box_np_ops', 'det3d.models.builder',
                 'det3d.utils.dist.dist_common'):
        _pkg(name)
    trainer = _pkg('det3d.torchie.trainer')
    trainer.load_checkpoint = lambda *a, **k: None
    reg = _pkg('det3d.models.registry')
    reg.HEADS = reg.NECKS = reg.READERS = _Registry()
    # real sources
    _load('det3d.torchie.cnn', 'torchie/cnn/weight_init.py')
    _load('det3d.core.utils.circle_nms_jit', 'core/utils/circle_nms_jit.py')      # numba
    _load('det3d.core.utils.center_utils', 'core/utils/center_utils.py')
    misc = _load('det3d.models.utils_misc', 'models/utils/misc.py')
    norm = _load('det3d.models.utils_norm', 'models/utils/norm.py')
    utils = _pkg('det3d.models.utils')
    for src in (misc, norm):
        for k, v in vars(src).items():
            if not k.startswith('__'):
                setattr(utils, k, v)
    _load('det3d.models.losses.centernet_loss', 'models/losses/centernet_loss.py')
    reader = _load('det3d.models.readers.voxel_encoder', 'models/readers/voxel_encoder.py')
    rpn = _load('det3d.models.necks.rpn', 'models/necks/rpn.py')
    head = _load('det3d.models.bbox_heads.center_head', 'models/bbox_heads/center_head.py')
    return reader, rpn, head


TASKS = [dict(num_class=1, class_names=['car']), dict(num_class=2, class_names=['truck', 'bus'])]
COMMON = {'reg': (2, 2), 'height': (1, 2), 'dim': (3, 2), 'rot': (2, 2), 'vel': (2, 2)}
CODE_W = [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 1.0, 1.0]
RPN_CFG = dict(layer_nums=[2, 1], ds_layer_strides=[1, 2], ds_num_filters=[8, 16], us_layer_strides=[1, 2],
               us_num_filters=[8, 8], num_input_features=12)
TEST_CFG = dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], circular_nms=True, min_radius=[4, 12],
                nms=dict(nms_pre_max_size=1000, nms_post_max_size=83, nms_iou_threshold=0.2), score_threshold=0.1,
                pc_range=[-54, -54], out_size_factor=8, voxel_size=[0.075, 0.075])


class AttrDict(dict):
    """Attribute access like det3d's Config (test_cfg.nms.nms_post_max_size, test_cfg.get(...))."""
    def __getattr__(self, k):
        v = self[k]
        return AttrDict(v) if isinstance(v, dict) else v


HEAD_CFG = dict(in_channels=16, tasks=TASKS, dataset='nuscenes', weight=0.25, code_weights=CODE_W,
                common_heads=COMMON, share_conv_channel=8, dcn_head=False)


def inputs(seed=0, b=2, h=20, w=24, max_objs=12):
    g = torch.Generator().manual_seed(seed)
    bev = torch.randn(b, 12, h, w, generator=g)
    voxels = torch.randn(50, 10, 5, generator=g)
    num = torch.randint(1, 11, (50,), generator=g)
    voxels = voxels * (torch.arange(10)[None, :, None] < num[:, None, None])
    example = {'hm': [], 'ind': [], 'mask': [], 'cat': [], 'anno_box': []}
    for t in TASKS:
        n_cls = len(t['class_names'])
        hm = torch.rand(b, n_cls, h, w, generator=g) ** 4
        ind = torch.randint(0, h * w, (b, max_objs), generator=g)
        mask = (torch.rand(b, max_objs, generator=g) < 0.6).to(torch.uint8)
        cat = torch.randint(0, n_cls, (b, max_objs), generator=g)
        for bi in range(b):                        # annotated peaks carry heat 1 in their class map
            for oi in range(max_objs):
                if mask[bi, oi]:
                    hm[bi, cat[bi, oi]].view(-1)[ind[bi, oi]] = 1.0
        example['hm'].append(hm); example['ind'].append(ind); example['mask'].append(mask)
        example['cat'].append(cat); example['anno_box'].append(torch.randn(b, max_objs, 10, generator=g))
    return bev, voxels, num, example


def main():
    import logging
    reader, rpn, head = import_reference()
    bev, voxels, num, example = inputs()
    out = {'bev': bev.numpy(), 'voxels': voxels.numpy(), 'num_points': num.numpy()}
    for k, lst in example.items():
        for t, v in enumerate(lst):
            out[f'ex_{k}_{t}'] = v.numpy()
    out['vfe'] = reader.VoxelFeatureExtractorV3(num_input_features=5)(voxels, num).numpy()

    torch.manual_seed(0)
    neck = rpn.RPN(logger=logging.getLogger('RPN'), **RPN_CFG)
    neck.init_weights()
    ch = head.CenterHead(**HEAD_CFG)
    for m in list(neck.modules()) + list(ch.modules()):       # non-trivial BN statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    for k, v in neck.state_dict().items():
        out['rpn/' + k] = v.clone().numpy()          # clone: train mode updates the BN statistics in place
    for k, v in ch.state_dict().items():
        out['head/' + k] = v.clone().numpy()
    for mode in ('eval', 'train'):
        getattr(neck, mode)(); getattr(ch, mode)()
        x_in = bev.clone().requires_grad_(mode == 'train')
        with torch.set_grad_enabled(mode == 'train'):
            x = neck(x_in)
            preds, shared = ch(x)
            out[f'{mode}_rpn'] = x.detach().numpy()
            out[f'{mode}_shared'] = shared.detach().numpy()
            for t, p in enumerate(preds):
                for k, v in p.items():
                    out[f'{mode}_pred_{t}_{k}'] = v.detach().clone().numpy()
            losses = ch.loss(example, preds, None)               # NB: applies sigmoid_ to preds[t]['hm'] in place
            out[f'{mode}_loss'] = np.array([float(v) for v in losses['loss']])
            out[f'{mode}_hm_loss'] = np.array([float(v) for v in losses['hm_loss']])
            out[f'{mode}_loc_loss_elem'] = np.stack([v.numpy() for v in losses['loc_loss_elem']])
            if mode == 'train':                                  # backward check: d(sum of task losses)/d(BEV input)
                sum(losses['loss']).backward()
                out['train_grad_bev'] = x_in.grad.numpy()
    # ---- predict: decode + circle NMS (the reference's numba circle_nms), plain and double-flip ----
    # (the train-mode pass above updated the BN running statistics: back to the stored weights)
    neck.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in out.items() if k.startswith('rpn/')})
    ch.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in out.items() if k.startswith('head/')})
    neck.eval(); ch.eval()
    g = torch.Generator().manual_seed(11)
    bev4 = torch.randn(4, 12, 20, 24, generator=g)
    out['bev4'] = bev4.numpy()
    for tag, x_in, cfg in (('plain', bev4, dict(TEST_CFG)), ('flip', bev4, dict(TEST_CFG, double_flip=True))):
        with torch.no_grad():
            preds, _ = ch(neck(x_in))
            dets = ch.predict({}, preds, AttrDict(cfg))
        out[f'pred_{tag}_n'] = np.array(len(dets))
        for i, d in enumerate(dets):
            out[f'pred_{tag}_{i}_boxes'] = d['box3d_lidar'].numpy()
            out[f'pred_{tag}_{i}_scores'] = d['scores'].numpy()
            out[f'pred_{tag}_{i}_labels'] = d['label_preds'].numpy()
        print(tag, [len(d['scores']) for d in dets])
    np.savez_compressed(os.path.join(HERE, 'centerpoint.npz'), **out)
    print(len(out), 'arrays;', 'eval loss', out['eval_loss'], 'train loss', out['train_loss'])


if __name__ == '__main__':
    main()
