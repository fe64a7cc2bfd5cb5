"""GPU tests of the C++ autograd nodes (link_b200/_ext.py, csrc_ext/autograd_ops.cpp) against the python
autograd Functions they stand in for: same liblinkb200 kernels, so outputs and input gradients are
bit-identical and the weight gradients agree to the float-atomic noise of the tcgen05 wgrad flush; and of
the off-stream upload / parallel coordinate pyramid of the training path."""
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


def _toggle(monkeypatch, on):
    from link_b200 import _ext
    monkeypatch.setattr(_ext, 'ENABLED', on)
    monkeypatch.setattr(_ext, '_mod', None)
    monkeypatch.setattr(_ext, '_tried', False)


def test_cpp_conv_and_batchnorm_nodes_equal_python_functions(dev, monkeypatch):
    import link_b200.nn.functional as F
    from link_b200 import SparseTensor, _ext
    from link_b200.utils.synthetic import random_voxels
    if _ext._so_path() is None:
        pytest.skip('C++ autograd extension not built')
    coords = torch.from_numpy(random_voxels(5000, 30, seed=1, batch=2)).to(dev)
    res = {}
    for use in (True, False):
        _toggle(monkeypatch, use)
        torch.manual_seed(0)
        x = SparseTensor(torch.randn(len(coords), 64, device=dev, requires_grad=True), coords, 1)
        x.cmaps[x.stride] = x.coords
        w = torch.randn(27, 64, 32, device=dev, requires_grad=True)          # submanifold 3^3: dgrad on the forward map
        w2 = torch.randn(8, 32, 32, device=dev, requires_grad=True)          # strided 2^3: dgrad on the inverted map
        y = F.conv3d(x, w, 3)
        z = F.conv3d(y, w2, 2, stride=2)
        bn = torch.nn.BatchNorm1d(32).to(dev).train()
        r = torch.randn(z.F.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(2)).requires_grad_(True)
        o = F.batch_norm_act(z.F, bn, True, r)
        o.square().sum().backward()
        res[use] = (o.detach(), x.F.grad.clone(), r.grad.clone(), w.grad.clone(), w2.grad.clone(), bn.weight.grad.clone(),
                    bn.running_var.clone(), type(y.F.grad_fn).__name__)
    assert res[True][7] != res[False][7], 'the C++ node did not run'
    for i, name in enumerate(['out', 'd rows', 'd residual']):
        assert torch.equal(res[True][i], res[False][i]), name
    for i, name in ((3, 'd w'), (4, 'd w2'), (5, 'd gamma'), (6, 'running_var')):
        a, b = res[True][i], res[False][i]
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=2e-6 * float(b.abs().max()), err_msg=name)


def test_from_host_ahead_and_parallel_pyramid(dev, monkeypatch):
    """from_host(ahead=True) (off-stream upload into copy-stream buffers, also from write-combined pinned
    memory) and the parallel coordinate pyramid give the same training step (loss, gradients, level
    coordinates) as the plain upload and the level-by-level pyramid."""
    import link_b200.linkencoder as le
    from link_b200 import SparseTensor
    from link_b200.linkencoder import ELKEncoder
    from link_b200.tensor import pinned_empty
    from link_b200.utils.synthetic import kitti_like_voxels
    c3, f4 = kitti_like_voxels(6000, seed=8)
    coords_h = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)).pin_memory()
    feats_h = pinned_empty(f4.shape, torch.float32, write_combined=True)
    assert feats_h.is_pinned()
    feats_h.copy_(torch.from_numpy(f4.astype(np.float32)))
    target = torch.randint(0, 19, (coords_h.shape[0],), device=dev)
    res = []
    for ahead, par in ((True, True), (False, False)):
        monkeypatch.setattr(le, 'PARALLEL_PYRAMID', par)
        torch.manual_seed(0)
        net = ELKEncoder(num_classes=19, cr=0.5, baseop='cos', r=3, s=7, groups=2).to(dev).train()
        st = SparseTensor.from_host(feats_h, coords_h, 1, device=dev, ahead=ahead)
        loss = torch.nn.functional.cross_entropy(net(st), target)
        loss.backward()
        torch.cuda.synchronize()
        res.append((loss.detach(), net.stem[0].kernel.grad.clone(), {k: v.clone() for k, v in st.cmaps.items()}))
    np.testing.assert_allclose(float(res[0][0]), float(res[1][0]), rtol=1e-5)
    np.testing.assert_allclose(res[0][1].cpu().numpy(), res[1][1].cpu().numpy(), rtol=1e-3, atol=1e-3 * float(res[1][1].abs().max()))
    assert sorted(res[0][2]) == sorted(res[1][2])
    for k in res[1][2]:
        assert torch.equal(res[0][2][k], res[1][2][k]), f'coordinates of stride {k}'
