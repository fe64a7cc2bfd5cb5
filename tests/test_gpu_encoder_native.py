"""GPU tests of the native encoder executor (csrc/encoder.cu: lk_elk_encoder_fwd behind
ELKEncoder.forward in inference): one library call per scan against the per-layer python path of the
same model (same kernels: the results agree to the float-atomic noise of the LinK pre-aggregation), on
one- and two-frame batches, with the branch / index streams on and off, at the widths the executor
serves (64 and 32 channels) and with the fall-back for the ones it does not (16)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from link_b200 import _capi
    _capi.lib()
    return torch.device('cuda:0')


def _scan(n, seed, batch=1):
    from link_b200.utils.synthetic import kitti_like_voxels
    cs, fs = [], []
    for b in range(batch):
        c3, f4 = kitti_like_voxels(n, seed=seed + b)
        cs.append(np.concatenate([c3, np.full((len(c3), 1), b, np.int32)], 1).astype(np.int32))
        fs.append(f4.astype(np.float32))
    return np.concatenate(cs), np.concatenate(fs)


def _model(dev, cr, baseop, groups, seed=0):
    from link_b200.linkencoder import ELKEncoder
    torch.manual_seed(seed)
    net = ELKEncoder(num_classes=19, cr=cr, baseop=baseop, r=3, s=7, groups=groups).to(dev).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    return net


def _run(net, coords, feats, dev):
    from link_b200 import SparseTensor
    st = SparseTensor(torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev), 1)
    with torch.no_grad():
        out = net(st)
    return out, st


@pytest.mark.parametrize('cr,baseop,groups,n,batch', [(1.0, 'cos', 2, 30_000, 1), (1.0, 'sin', 1, 9_000, 2),
                                                      (0.5, 'cos', 2, 12_000, 1), (1.0, 'cos_x', 1, 8_000, 1)])
def test_native_encoder_equals_per_layer_path(dev, monkeypatch, cr, baseop, groups, n, batch):
    import link_b200.linkencoder as le
    coords, feats = _scan(n, seed=int(cr * 10) + n % 7, batch=batch)
    net = _model(dev, cr, baseop, groups)
    monkeypatch.setattr(le, 'NATIVE_ENCODER', False)
    want, st_w = _run(net, coords, feats, dev)
    monkeypatch.setattr(le, 'NATIVE_ENCODER', True)
    got, st_g = _run(net, coords, feats, dev)
    assert '_lk_enc_native' in net.__dict__, 'the native executor did not run'
    # level coordinates: identical arrays (the pyramid is derived from the input coordinates directly)
    assert sorted(st_g.cmaps) == sorted(st_w.cmaps)
    for k in st_w.cmaps:
        assert torch.equal(st_g.cmaps[k], st_w.cmaps[k]), f'coordinates of stride {k}'
    scale = float(want.abs().max())
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5 * max(1.0, scale))
    # one stream, and no branch overlap: same values
    monkeypatch.setattr(le, 'BRANCH_OVERLAP', False)
    net.__dict__.pop('_lk_enc_native')
    got2, _ = _run(net, coords, feats, dev)
    np.testing.assert_allclose(got2.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5 * max(1.0, scale))
    import link_b200.elk as elk
    monkeypatch.setattr(elk, 'SINGLE_STREAM', True)
    got3, _ = _run(net, coords, feats, dev)
    np.testing.assert_allclose(got3.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5 * max(1.0, scale))


def test_native_encoder_follows_weight_updates_and_falls_back(dev, monkeypatch):
    """The argument template is keyed by the parameter versions: an in-place weight update is seen by the
    next call; widths the executor does not serve (cr = 0.25: 16 channels), training mode and autograd
    take the per-layer path."""
    import link_b200.linkencoder as le
    coords, feats = _scan(6_000, seed=5)
    net = _model(dev, 1.0, 'cos', 2)
    a, _ = _run(net, coords, feats, dev)
    with torch.no_grad():
        net.stage2[0].net[0].kernel.mul_(1.5)
        net.elk3.norm.weight.add_(0.25)
        net.stem[1].running_var.mul_(2.0)
    b, _ = _run(net, coords, feats, dev)
    monkeypatch.setattr(le, 'NATIVE_ENCODER', False)
    want, _ = _run(net, coords, feats, dev)
    monkeypatch.setattr(le, 'NATIVE_ENCODER', True)
    assert float((a - b).abs().max()) > 1e-3
    np.testing.assert_allclose(b.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5 * max(1.0, float(want.abs().max())))
    small = _model(dev, 0.25, 'cos', 2)
    _run(small, coords, feats, dev)
    assert '_lk_enc_native' not in small.__dict__
    net.train()
    assert not net._native_ok(le.SparseTensor(torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev), 1))


def test_native_encoder_from_host_upload(dev):
    """A scan whose features are still crossing PCIe (SparseTensor.from_host) goes through the executor
    and gives the same logits as device-resident inputs."""
    from link_b200 import SparseTensor
    coords, feats = _scan(10_000, seed=9)
    net = _model(dev, 1.0, 'cos', 2)
    want, _ = _run(net, coords, feats, dev)
    st = SparseTensor.from_host(torch.from_numpy(feats).pin_memory(), torch.from_numpy(coords).pin_memory(), 1, device=dev)
    with torch.no_grad():
        got = net(st)
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5 * max(1.0, float(want.abs().max())))
