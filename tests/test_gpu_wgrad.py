"""GPU tests of the tensor-core weight gradient (lk_conv_wgrad_prepass + lk_conv_wgrad_tc) against a
float64 contraction of the same kernel map, and through ConvolutionFunction.backward."""
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


def _kmap(dev, n, ksize, stride, cin):
    from link_b200 import SparseTensor
    from link_b200.nn.functional.conv import build_kernel_map
    from link_b200.utils.synthetic import kitti_like_voxels, random_voxels
    if n >= 9_000:
        c3, _ = kitti_like_voxels(n, seed=3)
        coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    else:
        coords = random_voxels(n, 24, seed=n)
    st = SparseTensor(torch.zeros(len(coords), cin, device=dev), torch.from_numpy(coords).to(dev), 1)
    return build_kernel_map(st, (ksize,) * 3, (stride,) * 3, (1, 1, 1), want_plan=True)


def _want(x, g, rel):
    """float64: gw[k] = sum_{o: rel[k,o] >= 0} x[rel[k,o]]^T g[o]"""
    K = rel.shape[0]
    xn, gn = x.double().cpu().numpy(), g.double().cpu().numpy()
    out = np.zeros((K, xn.shape[1], gn.shape[1]))
    for k in range(K):
        hit = rel[k] >= 0
        out[k] = xn[rel[k][hit]].T @ gn[hit]
    return out


@pytest.mark.parametrize('n,cin,cout,ksize,stride', [(20_000, 64, 64, 3, 1), (3_000, 32, 64, 3, 1),
                                                     (130, 64, 32, 3, 1), (9_000, 32, 32, 2, 2),
                                                     (1, 64, 64, 3, 1), (9_000, 128, 128, 3, 1),
                                                     (2_000, 64, 128, 3, 1), (2_500, 128, 32, 2, 2),
                                                     (700, 128, 64, 3, 1), (300, 32, 128, 3, 1),
                                                     (40_000, 32, 32, 3, 1)])
@pytest.mark.parametrize('transposed', [False, True])
def test_wgrad_tc_vs_float64(dev, n, cin, cout, ksize, stride, transposed):
    """Both directions of the relation (forward map with the plan order; inverted map without), all
    nine channel combinations, maps with 8 and 27 offsets, a single row, tiles past the end."""
    from link_b200 import _capi
    km = _kmap(dev, n, ksize, stride, cin)
    rel = (km.inv if transposed else km.nbr)
    K, rows = rel.shape
    n_src = km.n_out if transposed else km.n_in
    gen = torch.Generator().manual_seed(n + cin)
    x = torch.randn(n_src, cin, generator=gen).to(dev)
    g = torch.randn(rows, cout, generator=gen).to(dev)
    nbrp, perm, masks = km.wgrad_relation(transposed)
    assert (perm is not None) == (not transposed)
    # the pre-pass: masks and the permuted relation
    rel_h = rel.cpu().numpy()
    perm_h = perm.cpu().numpy() if perm is not None else np.arange(rows)
    relp = rel_h[:, perm_h]
    assert np.array_equal(nbrp.cpu().numpy(), relp)
    pad = (-rows) % 64
    hit = np.concatenate([relp >= 0, np.zeros((K, pad), bool)], 1).reshape(K, -1, 64).any(2)
    want_masks = (hit * (1 << np.arange(K, dtype=np.int64))[:, None]).sum(0)
    assert np.array_equal(masks.cpu().numpy().view(np.uint32).astype(np.int64), want_masks)
    want = _want(x, g, rel_h)
    scale = max(1.0, float(np.abs(want).max()))
    for slots in (0, 1, 2):
        gw = torch.full((K, cin, cout), float('nan'), device=dev)
        _capi.check(_capi.lib().lk_conv_wgrad_tc(_capi.ptr(x), _capi.ptr(g), _capi.ptr(nbrp), _capi.ptr(perm),
                                                 _capi.ptr(masks), rows, K, cin, cout, _capi.ptr(gw), slots,
                                                 _capi.stream()), 'lk_conv_wgrad_tc')
        err = float(np.abs(gw.cpu().numpy() - want).max()) / scale
        assert err < 2e-5, (slots, err)          # 3xTF32: ~1e-6 relative to the largest entry


def test_wgrad_tc_through_autograd_matches_ffma(dev, monkeypatch):
    """ConvolutionFunction.backward: the tensor-core weight gradient equals the FFMA kernel's."""
    import link_b200.nn.functional as F
    import link_b200.nn.functional.conv as conv_mod
    from link_b200 import SparseTensor
    from link_b200.utils.synthetic import random_voxels
    coords = torch.from_numpy(random_voxels(6_000, 32, seed=9)).to(dev)
    gen = torch.Generator().manual_seed(3)
    f = torch.randn(len(coords), 32, generator=gen).to(dev)
    ws = [(torch.randn(27, 32, 64, generator=gen) * 0.1).to(dev), (torch.randn(8, 64, 128, generator=gen) * 0.1).to(dev),
          (torch.randn(8, 128, 32, generator=gen) * 0.1).to(dev)]
    grads = {}
    for flag in (True, False):
        monkeypatch.setattr(conv_mod, 'USE_TC_WGRAD', flag)
        x = SparseTensor(f.clone().requires_grad_(True), coords, 1)
        x.cmaps[x.stride] = x.coords
        w = [t.clone().requires_grad_(True) for t in ws]
        y = F.conv3d(F.conv3d(F.conv3d(x, w[0], 3), w[1], 2, stride=2), w[2], 2, stride=2, transposed=True)
        y.F.square().sum().backward()
        grads[flag] = [t.grad.clone() for t in w]
    for a, b in zip(grads[True], grads[False]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4 * float(b.abs().max()))


def test_conv_tf32_mode_tolerance(dev):
    """Reduced-precision mode of the tensor-core conv (set_precision('tf32'): one MMA on operands
    truncated to tf32, fp32 accumulation).  Stated tolerance: 2e-3 of the largest output (10-bit
    mantissas, truncation: <= 2^-10 relative per operand); the default mode stays at 1e-5."""
    from link_b200.nn.functional import conv as conv_mod
    km = _kmap(dev, 20_000, 3, 1, 64)
    K, rows = km.nbr.shape
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(rows, 64, generator=gen).to(dev)
    w = (torch.randn(K, 64, 64, generator=gen) / 16).to(dev)
    rel = km.nbr.cpu().numpy()
    want = np.zeros((rows, 64))
    xn, wn = x.double().cpu().numpy(), w.double().cpu().numpy()
    for k in range(K):
        hit = rel[k] >= 0
        want[hit] += xn[rel[k][hit]] @ wn[k]
    scale = np.abs(want).max()
    try:
        errs = {}
        for prec in ('fp32', 'tf32'):
            conv_mod.set_precision(prec)
            got = conv_mod._conv_fwd(x, w, km.nbr, rows, kmap=km).double().cpu().numpy()
            errs[prec] = np.abs(got - want).max() / scale
    finally:
        conv_mod.set_precision('fp32')
    assert errs['fp32'] < 1e-5 and 1e-5 < errs['tf32'] < 2e-3, errs


def test_encoder_under_bf16_autocast(dev):
    """BASELINE config 3 names bf16: ELKEncoder trains under torch.autocast(bfloat16) the way the reference
    trains under fp16 autocast (its three custom Functions cast at their boundary, nn/functional/conv.py:19,
    voxelize.py:13, devoxelize.py:54): dense layers run in bf16, the sparse convs / the LinK aggregation take
    bf16 activations, accumulate in fp32 on their own kernels and hand bf16 back.  Stated tolerance of the
    logits against the fp32 run: 5e-2 of their range (8-bit mantissas through ~40 layers); gradients finite."""
    from link_b200 import SparseTensor
    from link_b200.linkencoder import ELKEncoder
    from link_b200.utils.synthetic import random_voxels
    coords = torch.from_numpy(random_voxels(4000, 40, seed=6, batch=2)).to(dev)
    torch.manual_seed(6)
    feats = torch.randn(coords.shape[0], 4, device=dev)
    target = torch.randint(0, 19, (coords.shape[0],), device=dev)
    net = ELKEncoder(num_classes=19, cr=0.5, baseop='cos', r=3, s=7, groups=2).to(dev).train()
    ref = net(SparseTensor(feats.clone(), coords, 1)).detach()
    net.zero_grad(set_to_none=True)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        logits = net(SparseTensor(feats.clone(), coords, 1))
        loss = torch.nn.functional.cross_entropy(logits.float(), target)
    loss.backward()
    assert torch.isfinite(loss)
    for n, p in net.named_parameters():
        if not n.startswith('up'):
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    err = float((logits.float() - ref).abs().max())
    assert err <= 5e-2 * float(ref.abs().max() - ref.min()) + 5e-2, err


@pytest.mark.parametrize('n,cin,cout,ksize,stride', [(20_000, 64, 64, 3, 1), (3_000, 32, 64, 3, 1), (130, 64, 32, 3, 1),
                                                     (9_000, 32, 32, 2, 2), (2_500, 128, 32, 2, 2), (700, 128, 64, 3, 1),
                                                     (4_000, 16, 16, 3, 1), (1, 64, 64, 3, 1)])
def test_conv_bf16_rows_vs_float64(dev, n, cin, cout, ksize, stride):
    """lk_conv_tc_fwd_bf16: bf16 rows in / bf16 rows out, fp32 accumulation, tf32 weights, fused epilogue
    with a bf16 residual.  Against a float64 contraction of the SAME bf16-rounded inputs; stated
    tolerance: 1.5 * 2^-8 of the output range (the bf16 rounding of the result dominates) -- and the
    planned launch equals the unplanned one bit for bit."""
    from link_b200.nn.functional import conv as conv_mod
    km = _kmap(dev, n, ksize, stride, cin)
    K, rows = km.nbr.shape
    gen = torch.Generator().manual_seed(n + cout)
    x = torch.randn(km.n_in, cin, generator=gen).to(dev).to(torch.bfloat16)
    w = (torch.randn(K, cin, cout, generator=gen) / np.sqrt(cin * 4)).to(dev)
    res = torch.randn(rows, cout, generator=gen).to(dev).to(torch.bfloat16)
    scale, shift = (torch.rand(cout, generator=gen) + 0.5).to(dev), torch.randn(cout, generator=gen).to(dev)
    got = conv_mod._conv_fwd(x, w, km.nbr, rows, None, scale, shift, res, True, kmap=km)
    plain = conv_mod._conv_fwd(x, w, km.nbr, rows, None, scale, shift, res, True)
    assert got.dtype == torch.bfloat16 and torch.equal(got, plain)
    rel = km.nbr.cpu().numpy()
    acc = np.zeros((rows, cout))
    xn, wn = x.double().cpu().numpy(), w.double().cpu().numpy()
    for k in range(K):
        hit = rel[k] >= 0
        acc[hit] += xn[rel[k][hit]] @ wn[k]
    want = np.maximum(acc * scale.double().cpu().numpy() + shift.double().cpu().numpy() + res.double().cpu().numpy(), 0)
    err = np.abs(got.double().cpu().numpy() - want).max()
    assert err <= 1.5 * 2.0 ** -8 * max(1.0, np.abs(want).max()), err


def test_encoder_inference_bf16_activations(dev):
    """ELKEncoder forward on bf16 activations (bf16 conv kernel, LinK blocks through fp32 at their boundary):
    logits within 5e-2 of the fp32 run's range."""
    from link_b200 import SparseTensor
    from link_b200.linkencoder import ELKEncoder
    from link_b200.utils.synthetic import random_voxels
    coords = torch.from_numpy(random_voxels(6000, 48, seed=8, batch=2)).to(dev)
    torch.manual_seed(8)
    feats = torch.randn(coords.shape[0], 4, device=dev)
    net = ELKEncoder(num_classes=19, cr=1.0, baseop='cos', r=3, s=7, groups=2).to(dev).eval()
    with torch.no_grad():
        ref = net(SparseTensor(feats.clone(), coords, 1))
        with torch.autocast('cuda', dtype=torch.bfloat16):
            got = net(SparseTensor(feats.clone().to(torch.bfloat16), coords, 1))
    err = float((got.float() - ref).abs().max())
    assert torch.isfinite(got).all() and err <= 5e-2 * float(ref.max() - ref.min()) + 5e-2, err
