"""GPU tests of the fused training-mode BatchNorm (+ shortcut add + ReLU) over sparse feature rows
(csrc/bn.cu behind nn.functional.batch_norm_act) against torch.nn.BatchNorm1d -> add -> relu:
forward values, running statistics, num_batches_tracked, and every gradient (rows, gamma, beta,
residual).  Tolerances: the two implementations round the batch statistics differently (double
accumulators here, fp32 Welford in PyTorch): 2e-5 relative to the output / gradient scale."""
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


def _close(a, b, what, tol=2e-5):
    a, b = a.detach().double().cpu().numpy(), b.detach().double().cpu().numpy()
    scale = max(1.0, float(np.abs(b).max()))
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=tol * scale, err_msg=what)


@pytest.mark.parametrize('n,c,relu,res', [(2, 4, False, False), (7, 16, True, False), (1000, 32, True, True),
                                          (50_000, 64, True, True), (160_000, 64, False, False),
                                          (33_333, 128, True, False), (5000, 48, False, True), (4097, 5 * 4, True, True)])
def test_fused_bn_vs_torch(dev, n, c, relu, res):
    from link_b200.nn.functional.norm import batch_norm_act, BatchNormActFunction
    g = torch.Generator().manual_seed(n + c)
    x = (torch.randn(n, c, generator=g) * 3 + torch.randn(c, generator=g) * 5).to(dev)   # large means: E[x^2] - mean^2 cancels
    r = torch.randn(n, c, generator=g).to(dev) if res else None
    go = torch.randn(n, c, generator=g).to(dev)
    bns = []
    for _ in range(2):
        bn = torch.nn.BatchNorm1d(c, eps=1e-3, momentum=0.01).to(dev).train()
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.5, 0.5)
            bn.running_mean.uniform_(-1, 1)
            bn.running_var.uniform_(0.5, 2)
        bns.append(bn)
    bns[1].load_state_dict(bns[0].state_dict())
    if relu:
        # an element whose pre-activation is within round-off of 0 may land on either side of the ReLU in
        # the two implementations: it gets no incoming gradient, so the comparison is mask-independent
        with torch.no_grad():
            pre = torch.nn.functional.batch_norm(x, None, None, bns[0].weight, bns[0].bias, True, 0.0, 1e-3)
            pre = pre + r if res else pre
        go = go * (pre.abs() > 1e-4)
    outs = []
    for fused, bn in zip((True, False), bns):
        xi = x.clone().requires_grad_(True)
        ri = r.clone().requires_grad_(True) if res else None
        if fused:
            y = batch_norm_act(xi, bn, relu, ri)
            # the fused op's node: the python Function, or the C++ node (link_b200/_ext.py) -- not PyTorch's batch norm
            assert type(y.grad_fn).__name__ in ('BatchNormActFunctionBackward', 'CppFunction'), type(y.grad_fn).__name__
        else:
            y = bn(xi)
            if res:
                y = y + ri
            if relu:
                y = torch.relu(y)
        y.backward(go)
        outs.append((y, xi.grad, bn.weight.grad, bn.bias.grad, ri.grad if res else None, bn))
    (y0, dx0, dg0, db0, dr0, bn0), (y1, dx1, dg1, db1, dr1, bn1) = outs
    _close(y0, y1, 'forward')
    _close(dx0, dx1, 'd rows', tol=5e-5)
    _close(dg0, dg1, 'd gamma', tol=5e-5)
    _close(db0, db1, 'd beta', tol=5e-5)
    if res:
        _close(dr0, dr1, 'd residual')
    _close(bn0.running_mean, bn1.running_mean, 'running_mean', tol=1e-6)
    _close(bn0.running_var, bn1.running_var, 'running_var', tol=1e-6)
    assert int(bn0.num_batches_tracked) == int(bn1.num_batches_tracked) == 1


def test_spnn_batchnorm_module_and_fallbacks(dev):
    """spnn.BatchNorm on a SparseTensor takes the fused kernels in training mode and PyTorch's op in
    eval mode, on CPU and for dtypes / widths the kernels do not serve; the encoder's training step uses
    them (same loss as with LINKB200_FUSED_BN off)."""
    import link_b200.nn as spnn
    import link_b200.nn.functional.norm as fn
    from link_b200 import SparseTensor
    coords = torch.randint(0, 50, (3000, 4), dtype=torch.int32, device=dev)
    x = torch.randn(3000, 32, device=dev)
    bn = spnn.BatchNorm(32).to(dev).train()
    ref = torch.nn.BatchNorm1d(32).to(dev).train()
    y = bn(SparseTensor(x, coords, 1)).F
    _close(y, ref(x), 'spnn.BatchNorm train')
    _close(bn.running_var, ref.running_var, 'running_var', tol=1e-6)
    bn.eval(); ref.eval()
    _close(bn(SparseTensor(x, coords, 1)).F, ref(x), 'eval')
    assert not fn.bn_act_supported(bn, x)
    bn.train()
    assert fn.bn_act_supported(bn, x)
    assert not fn.bn_act_supported(bn, x.half())
    assert not fn.bn_act_supported(bn, x[:, :30].contiguous())
    assert not fn.bn_act_supported(bn, x.cpu())
    assert not fn.bn_act_supported(bn, x[:1])


def test_encoder_training_step_same_with_and_without_fused_bn(dev, monkeypatch):
    import link_b200.nn.functional.norm as fn
    from link_b200 import SparseTensor
    from link_b200.linkencoder import ELKEncoder
    from link_b200.utils.synthetic import kitti_like_voxels
    c3, f4 = kitti_like_voxels(8000, seed=4)
    coords = torch.from_numpy(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)).to(dev)
    feats = torch.from_numpy(f4.astype(np.float32)).to(dev)
    target = torch.randint(0, 19, (coords.shape[0],), device=dev)
    res = []
    for fused in (True, False):
        monkeypatch.setattr(fn, 'USE_FUSED_BN', fused)
        torch.manual_seed(0)
        net = ELKEncoder(num_classes=19, cr=0.5, baseop='cos', r=3, s=7, groups=2).to(dev).train()
        loss = torch.nn.functional.cross_entropy(net(SparseTensor(feats, coords, 1)), target)
        loss.backward()
        res.append((loss.detach(), net.stem[0].kernel.grad.clone(), net.stage2[1].net[4].weight.grad.clone(),
                    net.stage1[0].net[1].running_var.clone()))
    _close(res[0][0], res[1][0], 'loss', tol=1e-5)
    _close(res[0][1], res[1][1], 'd stem kernel', tol=2e-3)
    _close(res[0][2], res[1][2], 'd BatchNorm weight', tol=2e-3)
    _close(res[0][3], res[1][3], 'running_var', tol=1e-5)
