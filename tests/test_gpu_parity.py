"""GPU parity tests: every call goes through the C-ABI library (liblinkb200.so) and is compared
with (a) the reference-generated golden fixtures and (b) the CPU oracle on the same seeded
inputs.  Integer / index results must be bit-exact; fp32 features within the stated tolerance."""
import numpy as np
import pytest
import torch

from conftest import golden_sd, load_golden
from oracle import link_oracle as O

pytestmark = pytest.mark.gpu

BLOCKS = ['block_g1_cosx_2x3', 'block_cos_g2_2x3', 'block_sin_g1_2x5', 'block_cosx_stride2']
# fp32 tolerance for block / conv outputs (rtol 1e-4, atol 1e-5); LayerNorm'd block
# outputs carry the reference's own accumulation-order noise, measured ~2e-5 abs on the fixtures.
RTOL, ATOL = 1e-4, 2e-5


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from link_b200 import _capi
    _capi.lib()
    return torch.device('cuda:0')


def cu(x, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    return t if dtype is None else t.to(dtype)


# ------------------------------------------------------------------ integer ops
def test_hash_kat(dev):
    import link_b200.nn.functional as F
    g = load_golden('kat')
    assert np.array_equal(F.sphash(cu(g['coords'], dev)).cpu().numpy(), g['hash'])
    for off, key in [(g['off2'], 'khash2'), (g['off3'], 'khash3')]:
        got = F.sphash(cu(g['coords_b0'], dev), cu(off, dev))
        assert np.array_equal(got.cpu().numpy(), g[key])
    q = F.sphash(cu(np.asarray([[1, 2, 3, 0], [9, 9, 9, 0], [0, 0, 0, 0]], dtype=np.int32), dev))
    assert F.sphashquery(q, cu(g['hash'], dev)).tolist() == [1, -1, 0]
    assert F.spcount(cu(np.asarray([0, 0, 2, -1, 2, 2], dtype=np.int32), dev), 4).tolist() == [2, 0, 3, 0]
    # empty inputs
    assert F.sphash(torch.zeros(0, 4, dtype=torch.int, device=dev)).shape == (0,)
    assert F.sphashquery(torch.zeros(0, dtype=torch.long, device=dev), cu(g['hash'], dev)).shape == (0,)
    assert F.sphashquery(q, torch.zeros(0, dtype=torch.long, device=dev)).tolist() == [-1, -1, -1]
    assert F.spcount(torch.zeros(0, dtype=torch.int, device=dev), 3).tolist() == [0, 0, 0]


def test_hash_and_query_random_vs_oracle(dev):
    import link_b200.nn.functional as F
    rng = np.random.default_rng(1)
    c = rng.integers(-5000, 5000, size=(200_000, 4)).astype(np.int32)
    c[:, 3] = rng.integers(0, 4, size=len(c))
    h = F.sphash(cu(c, dev)).cpu().numpy()
    assert np.array_equal(h, O.sphash(c))
    off = O.get_kernel_offsets(3, 2)
    kh = F.sphash(cu(c[:5000], dev), cu(off, dev)).cpu().numpy()
    assert np.array_equal(kh, O.sphash(c[:5000], off))
    # duplicates in the references: lowest index wins; misses -> -1
    ref = np.concatenate([h[:50_000], h[:10_000]])
    qry = np.concatenate([h[40_000:60_000], rng.integers(0, 2 ** 59, size=1000)])
    got = F.sphashquery(cu(qry, dev), cu(ref, dev)).cpu().numpy()
    assert np.array_equal(got, O.sphashquery(qry, ref))
    # 2-D query shape is preserved
    got2 = F.sphashquery(cu(kh, dev), cu(h[:5000], dev))
    assert got2.shape == kh.shape and np.array_equal(got2.cpu().numpy(), O.sphashquery(kh, h[:5000]))


@pytest.mark.parametrize('n,bits', [(1, 1), (2, 3), (2047, 8), (2048, 9), (2049, 16), (10_000, 5),
                                    (100_003, 24), (300_000, 40), (65_536, 64),
                                    (2_200_000, 21)])   # > 1024 tiles: unfused scan path
def test_sort_unique_vs_numpy(dev, n, bits):
    from link_b200.nn.functional import _index
    rng = np.random.default_rng(n + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi, size=n, dtype=np.uint64, endpoint=True) if bits < 64 else \
        rng.integers(0, 2 ** 63 - 1, size=n, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    if n > 10:
        keys[rng.integers(0, n, size=n // 3)] = keys[rng.integers(0, n, size=n // 3)]  # duplicates
    su = _index.sort_unique(cu(keys.view(np.int64), dev), bits, want_order=True)
    m = int(su.num.item())
    uq, inv, cnt = np.unique(keys, return_inverse=True, return_counts=True)
    assert m == len(uq)
    assert np.array_equal(su.unique[:m].cpu().numpy().view(np.uint64), uq)
    assert np.array_equal(su.inverse.cpu().numpy(), inv.reshape(-1))
    assert np.array_equal(su.counts[:m].cpu().numpy(), cnt)
    assert np.array_equal(su.order.cpu().numpy(), np.argsort(keys, kind='stable'))
    seg = su.seg[:m + 1].cpu().numpy()
    assert seg[0] == 0 and seg[-1] == n and np.array_equal(np.diff(seg), cnt)
    assert np.array_equal(su.sorted_rank.cpu().numpy(), inv.reshape(-1)[np.argsort(keys, kind='stable')])


def test_sort_unique_empty(dev):
    from link_b200.nn.functional import _index
    su = _index.sort_unique(torch.zeros(0, dtype=torch.int64, device=dev), 8, want_order=True)
    assert int(su.num.item()) == 0


@pytest.mark.parametrize('name', BLOCKS)
def test_block_index_maps_bit_exact(dev, name):
    import link_b200.nn.functional as F
    from link_b200 import SparseTensor
    from link_b200.elk import block_index
    g = load_golden(name)
    s, r = int(g['s']), int(g['r'])
    coords = cu(g['coords'], dev)
    assert np.array_equal(F.sphash(coords).cpu().numpy(), g['hash'])
    st = SparseTensor(cu(g['feats'], dev), coords, int(g['tstride']))
    bi = block_index(st, s)
    assert bi.m == len(g['small_C'])
    assert np.array_equal(bi.small_C.cpu().numpy(), g['small_C'])
    assert np.array_equal(bi.idx_query.cpu().numpy(), g['idx_query'])
    assert np.array_equal(bi.counts[:bi.m].cpu().numpy(), g['counts'])
    assert np.array_equal(bi.neighbors(r)[:bi.m].cpu().numpy(), g['nbr_idx'])


def test_block_index_negative_coords_batches_and_r3(dev):
    from link_b200 import SparseTensor
    from link_b200.elk import block_index
    rng = np.random.default_rng(3)
    c = np.unique(rng.integers(-300, 500, size=(60_000, 3)), axis=0).astype(np.int32)
    c = c[rng.permutation(len(c))]
    c = np.concatenate([c, rng.integers(0, 3, size=(len(c), 1)).astype(np.int32)], 1)
    for s, r in [(7, 3), (14, 3), (3, 2), (1, 2)]:
        st = SparseTensor(torch.zeros(len(c), 4, device=dev), cu(c, dev), 1)
        bi = block_index(st, s)
        small_C, idx, counts = O.block_index(c, s)
        assert bi.m == len(small_C)
        assert np.array_equal(bi.small_C.cpu().numpy(), small_C)
        assert np.array_equal(bi.idx_query.cpu().numpy(), idx)
        assert np.array_equal(bi.counts[:bi.m].cpu().numpy(), counts)
        assert np.array_equal(bi.neighbors(r)[:bi.m].cpu().numpy(), O.block_neighbors(small_C, r))


# ------------------------------------------------------------------ float ops
def test_voxelize_devoxelize_vs_golden_and_grads(dev):
    import link_b200.nn.functional as F
    g = load_golden('ops')
    feats = cu(g['feats'], dev).requires_grad_(True)
    idx, counts = cu(g['idx'], dev), cu(g['counts'], dev)
    vox = F.spvoxelize(feats, idx, counts)
    np.testing.assert_allclose(vox.detach().cpu().numpy(), g['vox'], rtol=1e-5, atol=1e-6)
    nb, w = cu(g['nb'], dev), cu(g['w'], dev)
    dev_out = F.spdevoxelize(vox, nb, w, 2)
    np.testing.assert_allclose(dev_out.detach().cpu().numpy(), g['dev'], rtol=1e-5, atol=2e-6)
    # gradients against the oracle's autograd
    go = torch.randn(dev_out.shape, generator=torch.Generator().manual_seed(0))
    dev_out.backward(go.to(dev))
    f_cpu = torch.from_numpy(g['feats']).requires_grad_(True)
    o = O.spdevoxelize(O.spvoxelize(f_cpu, g['idx'], g['counts']), g['nb'], torch.from_numpy(g['w']))
    o.backward(go)
    np.testing.assert_allclose(feats.grad.cpu().numpy(), f_cpu.grad.numpy(), rtol=1e-4, atol=1e-6)
    # odd channel count -> scalar path; r^3 = 27
    rng = np.random.default_rng(0)
    f2 = rng.standard_normal((3000, 7)).astype(np.float32)
    i2 = rng.integers(-1, 400, size=3000).astype(np.int32)
    c2 = O.spcount(i2, 400)
    got = F.spvoxelize(cu(f2, dev), cu(i2, dev), cu(c2, dev)).cpu().numpy()
    np.testing.assert_allclose(got, O.spvoxelize(torch.from_numpy(f2), i2, c2).numpy(), rtol=1e-5, atol=1e-6)
    nb3 = rng.integers(-1, 400, size=(500, 27)).astype(np.int32)
    w3 = rng.random((500, 27)).astype(np.float32)
    got = F.spdevoxelize(cu(got, dev), cu(nb3, dev), cu(w3, dev), 3).cpu().numpy()
    want = O.spdevoxelize(O.spvoxelize(torch.from_numpy(f2), i2, c2), nb3, torch.from_numpy(w3)).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)


def test_aux_batch2_reference_api(dev):
    from link_b200 import SparseTensor
    from link_b200.elk import voxel_to_aux
    g = load_golden('aux_batch2')
    st = SparseTensor(cu(g['feats'], dev), cu(g['coords'], dev), 1)
    aux, idx, counts = voxel_to_aux(st, int(g['s']))
    assert np.array_equal(aux.C.cpu().numpy(), g['aux_C'])
    assert np.array_equal(idx.cpu().numpy(), g['idx']) and idx.dtype == torch.int64
    assert np.array_equal(counts.cpu().numpy(), g['counts'])
    np.testing.assert_allclose(aux.F.cpu().numpy(), g['aux_F'], rtol=1e-5, atol=1e-6)
    assert aux.s == (int(g['s']),) * 3 and aux.kmaps is st.kmaps


def test_window_mean_semantics_r3(dev):
    from link_b200 import SparseTensor
    from link_b200.elk import aux_to_voxel, voxel_to_aux
    rng = np.random.default_rng(0)
    c = np.unique(rng.integers(-9, 14, size=(400, 3)), axis=0).astype(np.int32)
    c = np.concatenate([c, rng.integers(0, 2, size=(len(c), 1)).astype(np.int32)], 1)
    f = torch.randn(len(c), 8, generator=torch.Generator().manual_seed(0))
    for s, r in [(3, 2), (2, 3), (5, 3)]:
        st = SparseTensor(f.to(dev), cu(c, dev), 1)
        aux, idx, counts = voxel_to_aux(st, s)
        out = aux_to_voxel(aux, st, idx, counts, r)
        assert out is st
        np.testing.assert_allclose(out.F.cpu().numpy(), O.window_mean_bruteforce(f, c, s, r).numpy(),
                                   rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ conv + kernel maps
def test_conv_chain_kmaps_and_outputs(dev):
    import link_b200.nn as spnn
    import link_b200.nn.functional as F
    from link_b200 import SparseTensor
    g = load_golden('conv')
    coords = cu(g['coords'], dev)
    assert np.array_equal(F.spdownsample(coords, 2, 3, 1).cpu().numpy(), g['ds_k3s2'])
    assert np.array_equal(F.spdownsample(cu(g['y2_C'], dev), 2, 2, 2).cpu().numpy(), g['ds_k2s2_t2'])
    x = SparseTensor(cu(g['feats'], dev), coords, 1)
    x.cmaps[x.stride] = x.coords
    mods = [spnn.Conv3d(12, 20, 3), spnn.Conv3d(20, 24, 2, stride=2), spnn.Conv3d(24, 8, 3),
            spnn.Conv3d(8, 6, 2, stride=2, transposed=True), spnn.Conv3d(6, 5, 1)]
    ys = []
    with torch.no_grad():
        for m, wk in zip(mods, ['w1', 'w2', 'w3', 'w4', 'w5']):
            m.kernel.copy_(torch.from_numpy(g[wk]))
            m.to(dev)
            x = m(x)
            ys.append(x)
    assert np.array_equal(ys[1].C.cpu().numpy(), g['y2_C']) and ys[1].s == (2, 2, 2)
    assert np.array_equal(ys[3].C.cpu().numpy(), g['y4_C']) and ys[3].s == (1, 1, 1)
    for key, km in x.kmaps.items():
        if not (isinstance(key, tuple) and len(key) == 4 and isinstance(key[0], tuple)):
            continue
        tag = 'kmap_s%d_k%d_st%d' % (key[0][0], key[1][0], key[2][0])
        assert np.array_equal(km[0].cpu().numpy(), g[tag + '_nbmaps']), tag
        assert np.array_equal(km[1].cpu().numpy(), g[tag + '_nbsizes']), tag
    for y, k in zip(ys, ['y1', 'y2', 'y3', 'y4', 'y5']):
        np.testing.assert_allclose(y.F.cpu().numpy(), g[k], rtol=RTOL, atol=1e-5, err_msg=k)


@pytest.mark.parametrize('n,cin,cout,ksize,stride', [(20_000, 64, 64, 3, 1), (3_000, 32, 64, 3, 1),
                                                     (130, 64, 32, 3, 1), (9_000, 32, 32, 2, 2),
                                                     (1, 64, 64, 3, 1), (9_000, 128, 128, 3, 1),
                                                     (2_000, 64, 128, 3, 1), (2_500, 128, 32, 2, 2),
                                                     (700, 128, 64, 3, 1), (300, 32, 128, 3, 1),
                                                     (4_000, 16, 16, 3, 1), (3_000, 5, 16, 3, 1),
                                                     (2_000, 16, 32, 3, 1), (1_500, 48, 24, 2, 2)])   # padded to 32/64
def test_conv_plan_tile_skipping(dev, n, cin, cout, ksize, stride):
    """lk_conv_plan: perm is a permutation grouped by class, tile_mask is consistent with
    the kernel map; the planned tensor-core conv equals the unplanned one bit for bit (the plan only
    moves rows between tiles) and the float64 numpy contraction within fp32 tolerance."""
    import ctypes as Ct
    import link_b200.nn.functional as F
    from link_b200 import SparseTensor, _capi
    from link_b200.nn.functional.conv import build_kernel_map, _conv_fwd
    from link_b200.utils.synthetic import kitti_like_voxels, random_voxels
    if n >= 9_000:
        c3, _ = kitti_like_voxels(n, seed=3)
        coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    else:
        coords = random_voxels(n, 24, seed=n)
    st = SparseTensor(torch.zeros(len(coords), cin, device=dev), cu(coords, dev), 1)
    km = build_kernel_map(st, (ksize,) * 3, (stride,) * 3, (1, 1, 1))
    K, n_out = km.nbr.shape
    perm, tmask = km.plan()
    nbr = km.nbr.cpu().numpy()
    perm_h = perm.cpu().numpy()
    tmask_h = tmask.cpu().numpy().view(np.uint32)
    assert np.array_equal(np.sort(perm_h), np.arange(n_out))
    rowmask = ((nbr >= 0) * (1 << np.arange(K, dtype=np.int64))[:, None]).sum(0)
    pm = rowmask[perm_h]
    pad = (-n_out) % 128
    want_tm = np.bitwise_or.reduce(np.concatenate([pm, np.zeros(pad, np.int64)]).reshape(-1, 128), axis=1)
    assert np.array_equal(tmask_h.astype(np.int64), want_tm)
    if K <= 8:
        cls = pm & 255
    else:
        off = km.offsets.cpu().numpy()
        code = np.zeros(K, np.int64)
        for a in range(3):
            code |= (off[:, a] < 0).astype(np.int64) << (2 * a)
            code |= (off[:, a] > 0).astype(np.int64) << (2 * a + 1)
        cls = np.bitwise_or.reduce(np.where((nbr[:, perm_h] >= 0), code[:, None], 0), axis=0)
    assert np.all(np.diff(cls) >= 0), 'rows must be grouped by class'
    g = torch.Generator().manual_seed(n)
    x = torch.randn(km.n_in, cin, generator=g)
    w = torch.randn(K, cin, cout, generator=g) / np.sqrt(cin * 4)
    res = torch.randn(n_out, cout, generator=g)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    xd, wd, rd, sd, hd = (t.to(dev) for t in (x, w, res, scale, shift))
    planned = _conv_fwd(xd, wd, km.nbr, n_out, None, sd, hd, rd, True, kmap=km)
    plain = _conv_fwd(xd, wd, km.nbr, n_out, None, sd, hd, rd, True)
    assert torch.equal(planned, plain)
    acc = np.zeros((n_out, cout))
    xn, wn = x.numpy().astype(np.float64), w.numpy().astype(np.float64)
    for k in range(K):
        hit = nbr[k] >= 0
        acc[hit] += xn[nbr[k][hit]] @ wn[k]
    want = np.maximum(acc * scale.numpy() + shift.numpy() + res.numpy(), 0)
    np.testing.assert_allclose(planned.cpu().numpy(), want, rtol=RTOL, atol=ATOL)


def test_conv_backward_vs_oracle_autograd(dev):
    import link_b200.nn.functional as F
    from link_b200 import SparseTensor
    rng = np.random.default_rng(5)
    c = np.unique(rng.integers(0, 24, size=(3000, 3)), axis=0).astype(np.int32)
    c = np.concatenate([c, np.zeros((len(c), 1), np.int32)], 1)
    f = rng.standard_normal((len(c), 8)).astype(np.float32)
    w1 = (rng.standard_normal((27, 8, 12)) * 0.2).astype(np.float32)
    w2 = (rng.standard_normal((8, 12, 16)) * 0.2).astype(np.float32)
    w3 = (rng.standard_normal((8, 16, 4)) * 0.2).astype(np.float32)

    xg = SparseTensor(cu(f, dev).requires_grad_(True), cu(c, dev), 1)
    xg.cmaps[xg.stride] = xg.coords
    wg = [cu(w, dev).requires_grad_(True) for w in (w1, w2, w3)]
    yg = F.conv3d(F.conv3d(F.conv3d(xg, wg[0], 3), wg[1], 2, stride=2), wg[2], 2, stride=2,
                  transposed=True)
    xo = O.OTensor(torch.from_numpy(f.copy()).requires_grad_(True), c, 1)
    xo.cmaps[xo.s] = xo.C
    wo = [torch.from_numpy(w.copy()).requires_grad_(True) for w in (w1, w2, w3)]
    yo = O.conv3d(O.conv3d(O.conv3d(xo, wo[0], 3), wo[1], 2, stride=2), wo[2], 2, stride=2,
                  transposed=True)
    np.testing.assert_allclose(yg.F.detach().cpu().numpy(), yo.F.detach().numpy(), rtol=RTOL, atol=1e-5)
    go = torch.randn(yo.F.shape, generator=torch.Generator().manual_seed(1))
    yg.F.backward(go.to(dev))
    yo.F.backward(go)
    np.testing.assert_allclose(xg.F.grad.cpu().numpy(), xo.F.grad.numpy(), rtol=1e-3, atol=1e-5)
    for a, b in zip(wg, wo):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-3, atol=2e-5)


# ------------------------------------------------------------------ the LinK block
def _make_block(g, dev, variant='encoder'):
    from link_b200.elk import ELKBlock
    C = int(g['C'])
    blk = ELKBlock(C, C, groups=int(g['groups']), baseop=str(g['baseop']), variant=variant)
    missing = blk.load_state_dict(golden_sd(g), strict=True)
    return blk.to(dev)


@pytest.mark.parametrize('accurate', [False, True])
@pytest.mark.parametrize('name', BLOCKS)
def test_block_forward_fused_vs_golden(dev, name, accurate, monkeypatch):
    """accurate=True: libdevice sincosf, held to ATOL; accurate=False (default): SFU sin/cos after
    Cody-Waite reduction (abs error 2^-20.9 ~ 5e-7 per phase), which LayerNorm over as few as 8
    channels amplifies to at most ~3e-5 on the unit-scale outputs -> atol 4e-5."""
    import link_b200.elk as elk
    from link_b200 import SparseTensor
    monkeypatch.setattr(elk, 'ACCURATE_TRIG', accurate)
    atol = ATOL if accurate else 4e-5
    g = load_golden(name)
    blk = _make_block(g, dev).eval()
    st = SparseTensor(cu(g['feats'], dev), cu(g['coords'], dev), int(g['tstride']))
    with torch.no_grad():
        out = blk(st, int(g['s']), int(g['r']))
    assert out is st
    key = ((int(g['tstride']),) * 3, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    assert np.array_equal(st.kmaps[key][0].cpu().numpy(), g['nbmaps'])
    assert np.array_equal(st.kmaps[key][1].cpu().numpy(), g['nbsizes'])
    np.testing.assert_allclose(out.F.cpu().numpy(), g['out'], rtol=RTOL, atol=atol)


@pytest.mark.parametrize('name', BLOCKS)
def test_block_forward_composed_and_grads(dev, name):
    from link_b200 import SparseTensor
    g = load_golden(name)
    blk = _make_block(g, dev).train()
    feats = cu(g['feats'], dev).requires_grad_(True)
    st = SparseTensor(feats, cu(g['coords'], dev), int(g['tstride']))
    out = blk(st, int(g['s']), int(g['r']))
    np.testing.assert_allclose(out.F.detach().cpu().numpy(), g['out'], rtol=RTOL, atol=ATOL)
    go = torch.randn(out.F.shape, generator=torch.Generator().manual_seed(2))
    out.F.backward(go.to(dev))
    # oracle gradients
    p = {k: v.clone().requires_grad_(True) for k, v in golden_sd(g).items()}
    f_cpu = torch.from_numpy(g['feats']).requires_grad_(True)
    o = O.elk_block_forward(f_cpu, g['coords'], int(g['tstride']), p, int(g['s']), int(g['r']),
                            str(g['baseop']), int(g['groups']))
    o.backward(go)
    # float atomics in the backward scatter kernels: accumulation order varies run to run
    np.testing.assert_allclose(feats.grad.cpu().numpy(), f_cpu.grad.numpy(), rtol=2e-3, atol=1e-4)
    ours = dict(blk.named_parameters())
    for k in ['pre_mix.0.weight', 'pos_weight.0.weight', 'local_mix.0.kernel', 'norm.weight']:
        ref = p[k].grad.numpy()
        np.testing.assert_allclose(ours[k].grad.cpu().numpy(), ref, rtol=5e-3,
                                   atol=5e-5 * max(1.0, float(np.abs(ref).max())), err_msg=k)


@pytest.mark.parametrize('baseop,groups,C,s,r', [('cos', 2, 64, 7, 3), ('cos', 2, 32, 5, 3),
                                                 ('cos_x', 1, 16, 3, 2), ('sin', 4, 128, 7, 3),
                                                 ('cos', 1, 48, 7, 3), ('cos', 2, 8, 3, 3)])
def test_block_fused_vs_oracle_r3_multibatch(dev, baseop, groups, C, s, r):
    """Configurations the reference CPU path cannot run (r=3, batch>1): compare with the oracle."""
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock
    from link_b200.utils.synthetic import random_voxels
    coords = random_voxels(4000, 48, seed=C + s, batch=2)
    torch.manual_seed(C)
    blk = ELKBlock(C, C, groups=groups, baseop=baseop).eval()
    with torch.no_grad():
        for m in (blk.pre_mix[1], blk.norm, blk.norm_local):
            m.weight.uniform_(0.5, 1.5)
            m.bias.uniform_(-0.5, 0.5)
    feats = torch.randn(len(coords), C)
    want = O.elk_block_forward(feats, coords, 1, {k: v.detach() for k, v in blk.state_dict().items()},
                               s, r, baseop, groups)
    blk = blk.to(dev)
    with torch.no_grad():
        got = blk(SparseTensor(feats.to(dev), cu(coords, dev), 1), s, r).F
    np.testing.assert_allclose(got.cpu().numpy(), want.detach().numpy(), rtol=RTOL, atol=4e-5)


def test_encoder_forward_vs_golden(dev):
    from link_b200 import SparseTensor
    from link_b200.linkencoder import ELKEncoder
    g = load_golden('encoder_cosx_2x3')
    enc = ELKEncoder(num_classes=19, cr=0.25, baseop='cos_x', r=2, s=3, groups=1)
    res = enc.load_state_dict(golden_sd(g), strict=False)
    assert all(k.startswith('up') for k in res.missing_keys) and not res.unexpected_keys
    enc = enc.to(dev).eval()
    st = SparseTensor(cu(g['feats'], dev), cu(g['coords'], dev), 1)
    with torch.no_grad():
        logits = enc(st)
    sizes = [v.shape[0] for k, v in st.cmaps.items()]
    assert sizes == g['level_sizes'].tolist()
    np.testing.assert_allclose(logits.cpu().numpy(), g['logits'], rtol=1e-3, atol=1e-4)


# ------------------------------------------------------------------ BASELINE-size properties
def test_full_size_scan_properties(dev):
    """~120k-voxel SemanticKITTI-shaped scan, LinK cos (3x7)^3, C=64: index maps bit-exact vs the
    oracle; fused == composed; row-permutation equivariance; linearity in the features."""
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock, block_index
    from link_b200.utils.synthetic import kitti_like_voxels
    c3, _ = kitti_like_voxels(120_000, seed=0)
    coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    N = len(coords)
    assert abs(N - 120_000) < 2400
    s, r, C = 7, 3, 64
    st = SparseTensor(torch.zeros(N, C, device=dev), cu(coords, dev), 1)
    bi = block_index(st, s)
    small_C, idx, counts = O.block_index(coords, s)
    assert bi.m == len(small_C) and np.array_equal(bi.small_C.cpu().numpy(), small_C)
    assert np.array_equal(bi.idx_query.cpu().numpy(), idx)
    assert np.array_equal(bi.counts[:bi.m].cpu().numpy(), counts)
    assert np.array_equal(bi.neighbors(r)[:bi.m].cpu().numpy(), O.block_neighbors(small_C, r))
    assert int(bi.counts[:bi.m].sum()) == N

    torch.manual_seed(0)
    blk = ELKBlock(C, C, groups=2, baseop='cos').to(dev).eval()
    feats = torch.randn(N, C, device=dev)
    with torch.no_grad():
        fused = blk(SparseTensor(feats.clone(), cu(coords, dev), 1), s, r).F
    blk.train()
    composed = blk(SparseTensor(feats.clone().requires_grad_(True), cu(coords, dev), 1), s, r).F.detach()
    blk.eval()
    # phases reach |p| ~ 3e3 rad here: fp32 phase round-off (1 ulp of p ~ 2e-4) bounds agreement
    np.testing.assert_allclose(fused.cpu().numpy(), composed.cpu().numpy(), rtol=2e-3, atol=2e-3)
    # permutation equivariance (index build must not depend on row order)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(0)).to(dev)
    with torch.no_grad():
        fp = blk(SparseTensor(feats[perm].clone(), cu(coords, dev)[perm].contiguous(), 1), s, r).F
    np.testing.assert_allclose(fp.cpu().numpy(), fused[perm].cpu().numpy(), rtol=1e-3, atol=2e-4)
    # linearity of the aggregation in its input features
    from link_b200.elk import link_aggregate
    w = blk.pos_weight[0].weight
    f1, f2 = torch.randn(N, C, device=dev), torch.randn(N, C, device=dev)
    cc = cu(coords, dev)
    a = link_aggregate(f1, cc, bi, r, 'cos', w)
    b = link_aggregate(f2, cc, bi, r, 'cos', w)
    ab = link_aggregate(f1 + 2 * f2, cc, bi, r, 'cos', w)
    np.testing.assert_allclose(ab.cpu().numpy(), (a + 2 * b).cpu().numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('op,C,groups', [('cos', 64, 2), ('sin', 32, 1), ('cos_x', 16, 1), ('cos', 128, 4),
                                         ('cos', 48, 1), ('cos', 8, 1), ('cos_x', 4, 1)])
@pytest.mark.parametrize('n', [1, 31, 1000, 40_013])
def test_preagg_segmented_vs_storage_order_vs_numpy(dev, op, C, groups, n):
    """The two forms of pass 1 (storage order + atomics per LiDAR column; block order through the
    sort permutation) against a float64 numpy block sum of the weighted planes."""
    import ctypes as Ct
    from link_b200 import SparseTensor, _capi
    from link_b200.elk import block_index, _kernel_gen
    from link_b200.utils.synthetic import random_voxels
    coords = random_voxels(n, 40, seed=n + C)
    n = len(coords)
    st = SparseTensor(torch.zeros(n, C, device=dev), cu(coords, dev), 1)
    bi = block_index(st, 3)
    m = bi.m
    g = torch.Generator().manual_seed(C)
    f = torch.randn(n, C, generator=g)
    w = torch.randn(C // groups, 3, generator=g) * 0.3
    alpha = torch.rand(C // groups, generator=g) + 0.5 if op == 'cos_x' else None
    k = 3 if op == 'cos_x' else 2
    w_d, a_d = w.to(dev), (alpha.to(dev) if alpha is not None else None)   # keep alive: gen holds raw pointers
    gen = _kernel_gen(op, C, w_d, a_d, 1.0)
    L, stream = _capi.lib(), _capi.stream()
    fd = f.to(dev)
    out = []
    for seg in (False, True):
        sums = torch.zeros(n, k * C, device=dev)
        if seg:
            _capi.check(L.lk_link_preagg_seg_fwd(_capi.ptr(fd), _capi.ptr(st.C), _capi.ptr(bi.order),
                                                 _capi.ptr(bi.sorted_rank), n, Ct.byref(gen),
                                                 _capi.ptr(sums), stream), 'seg')
        else:
            _capi.check(L.lk_link_preagg_fwd(_capi.ptr(fd), _capi.ptr(st.C), _capi.ptr(bi.idx_query), n,
                                             Ct.byref(gen), _capi.ptr(sums), stream), 'storage')
        out.append(sums[:m].cpu().numpy())
        assert float(sums[m:].abs().max()) == 0.0 if m < n else True
    pos = (coords[:, :3].astype(np.float32) @ w.numpy().T.astype(np.float32))
    if alpha is not None:
        pos = pos * alpha.numpy()
    pos = np.tile(pos, (1, groups)).astype(np.float64)
    fn = f.numpy().astype(np.float64)
    planes = {'cos': [fn * np.cos(pos), fn * np.sin(pos)], 'sin': [fn * np.sin(pos), fn * np.cos(pos)],
              'cos_x': [fn * np.cos(pos), fn * np.sin(pos), fn * pos]}[op]
    want = np.zeros((m, k * C))
    np.add.at(want, bi.idx_query.cpu().numpy(), np.concatenate(planes, 1))
    for got in out:
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(out[0], out[1], rtol=1e-4, atol=1e-4)    # float atomics: order varies


@pytest.mark.parametrize('op,C,groups,s', [('cos', 64, 2, 7), ('cos_x', 64, 1, 3), ('sin', 16, 1, 5),
                                           ('cos', 128, 2, 3), ('cos', 32, 2, 2)])
def test_preagg_ring_long_ranges(dev, op, C, groups, s):
    """Balanced ring kernel at sizes where every lane group walks several chunks (q > 8 sorted
    positions per group, last chunk partial) and block runs straddle chunk, group and warp
    boundaries: float64 numpy block sums, rows >= M untouched, and batch > 1."""
    import ctypes as Ct
    from link_b200 import SparseTensor, _capi
    from link_b200.elk import block_index, _kernel_gen
    from link_b200.utils.synthetic import random_voxels
    coords = random_voxels(90_001, 80, seed=C + s, batch=2)
    n = len(coords)
    st = SparseTensor(torch.zeros(n, C, device=dev), cu(coords, dev), 1)
    bi = block_index(st, s)
    m = bi.m
    g = torch.Generator().manual_seed(C)
    f = torch.randn(n, C, generator=g)
    w = torch.randn(C // groups, 3, generator=g) * 0.1
    alpha = torch.rand(C // groups, generator=g) + 0.5 if op == 'cos_x' else None
    k = 3 if op == 'cos_x' else 2
    w_d, a_d = w.to(dev), (alpha.to(dev) if alpha is not None else None)
    gen = _kernel_gen(op, C, w_d, a_d, 1.0)
    L, stream = _capi.lib(), _capi.stream()
    sums = torch.zeros(n, k * C, device=dev)
    _capi.check(L.lk_link_preagg_seg_fwd(_capi.ptr(f.to(dev)), _capi.ptr(st.C), _capi.ptr(bi.order),
                                         _capi.ptr(bi.sorted_rank), n, Ct.byref(gen), _capi.ptr(sums), stream), 'seg')
    assert m < n and float(sums[m:].abs().max()) == 0.0
    pos = (coords[:, :3].astype(np.float32) @ w.numpy().T.astype(np.float32))
    if alpha is not None:
        pos = pos * alpha.numpy()
    pos = np.tile(pos, (1, groups)).astype(np.float64)
    fn = f.numpy().astype(np.float64)
    planes = {'cos': [fn * np.cos(pos), fn * np.sin(pos)], 'sin': [fn * np.sin(pos), fn * np.cos(pos)],
              'cos_x': [fn * np.cos(pos), fn * np.sin(pos), fn * pos]}[op]
    want = np.zeros((m, k * C))
    np.add.at(want, bi.idx_query.cpu().numpy(), np.concatenate(planes, 1))
    np.testing.assert_allclose(sums[:m].cpu().numpy(), want, rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------ dense pre_mix kernels
@pytest.mark.parametrize('c,tcore', [(16, False), (32, False), (64, False), (128, False),
                                     (32, True), (64, True)])
@pytest.mark.parametrize('n', [1, 127, 128, 129, 20_011])
def test_linear_layernorm_kernels(dev, c, tcore, n):
    """Fused Linear+LayerNorm (FFMA and tcgen05/3xTF32 variants) vs torch fp32."""
    from link_b200 import _capi
    g = torch.Generator().manual_seed(c + n)
    x = (torch.randn(n, c, generator=g) * 3).to(dev)
    w = (torch.randn(c, c, generator=g) / c ** 0.5).to(dev)
    gam = (torch.rand(c, generator=g) + 0.5).to(dev)
    bet = torch.randn(c, generator=g).to(dev)
    out = torch.empty_like(x)
    L = _capi.lib()
    fn = L.lk_linear_ln_tc_fwd if tcore else L.lk_linear_ln_fwd
    _capi.check(fn(_capi.ptr(x), _capi.ptr(w), _capi.ptr(gam), _capi.ptr(bet), 1e-6, n, c,
                   _capi.ptr(out), _capi.stream()), 'linear_ln')
    want = torch.nn.functional.layer_norm(x.double() @ w.double().t(), (c,), gam.double(),
                                          bet.double(), 1e-6).float()
    np.testing.assert_allclose(out.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=1e-5)


def test_unet_forward_vs_golden(dev):
    from link_b200 import SparseTensor
    from link_b200.linkunet import ELKUNet
    g = load_golden('unet_cosx_2x3')
    net = ELKUNet(num_classes=19, cr=0.25, baseop='cos_x', r=2, s=3, groups=1)
    net.load_state_dict(golden_sd(g), strict=True)
    net = net.to(dev).eval()
    st = SparseTensor(cu(g['feats'], dev), cu(g['coords'], dev), 1)
    with torch.no_grad():
        logits = net(st)
    np.testing.assert_allclose(logits.cpu().numpy(), g['logits'], rtol=1e-3, atol=2e-4)


# ------------------------------------------------------------------ detection block (TSELKBlock)
@pytest.mark.parametrize('baseop,C', [('cos', 16), ('cos', 64), ('sin', 32), ('cos_sin', 16), ('x', 16)])
def test_tselk_block_vs_oracle(dev, baseop, C):
    """TSELKBlock through the spconv-style adapter ((b,z,y,x) indices) vs the oracle's 'det'
    variant (ts_elk.py:144-230): fused path for cos/sin, composed path for cos_sin/x."""
    from link_b200.ts_elk import SparseConvTensor, TSELKBlock
    from link_b200.utils.synthetic import random_voxels
    coords = random_voxels(3000, 40, seed=C, batch=2)              # (x, y, z, b)
    torch.manual_seed(C)
    blk = TSELKBlock(C, C, baseop=baseop).eval()
    with torch.no_grad():
        for m in (blk.pre_mix[1], blk.norm, blk.norm_local):
            m.weight.uniform_(0.5, 1.5)
            m.bias.uniform_(-0.5, 0.5)
    feats = torch.randn(len(coords), C)
    sd = {k: v.detach() for k, v in blk.state_dict().items()}
    if baseop in ('cos', 'sin'):
        want = O.elk_block_forward(feats, coords, 1, sd, 7, 3, baseop, 1, variant='det')
    else:   # restated inline from ts_elk.py:196-222
        want = _tselk_reference(feats, coords, sd, baseop, C)
    blk = blk.to(dev)
    idx_bzyx = cu(coords[:, [3, 2, 1, 0]].copy(), dev)
    sct = SparseConvTensor(feats.to(dev), idx_bzyx, [41, 41, 41], 2)
    with torch.no_grad():
        out = blk(sct, 7)
    assert isinstance(out, SparseConvTensor) and torch.equal(out.indices, idx_bzyx)
    np.testing.assert_allclose(out.features.cpu().numpy(), want.detach().numpy(), rtol=RTOL, atol=4e-5)


def _tselk_reference(feats, coords, p, baseop, C):
    import torch.nn.functional as TF
    F_input = TF.layer_norm(TF.linear(feats, p['pre_mix.0.weight']), (C,), p['pre_mix.1.weight'],
                            p['pre_mix.1.bias'], 1e-6)
    local = O.conv3d(O.OTensor(feats, coords, 1), p['local_mix.0.kernel'], 3).F
    pos = TF.linear(torch.from_numpy(coords[:, :3].astype(np.float32)), p['pos_weight.0.weight'])
    if baseop == 'x':
        pos = pos[:, :C // 2].repeat(1, 2)
        lin = F_input * pos
        aux, sc, idx, cnt = O.voxel_to_aux(lin.contiguous(), coords, 7)
        new = O.aux_to_voxel(aux, sc, idx, cnt, 3) - lin
    else:
        sin, cos = torch.sin(pos), torch.cos(pos)
        aux, sc, idx, cnt = O.voxel_to_aux(torch.cat([F_input * cos, F_input * sin], 1), coords, 7)
        vf = O.aux_to_voxel(aux, sc, idx, cnt, 3)
        new = (vf[:, :C] * cos + vf[:, C:] * sin) + (vf[:, C:] * cos - vf[:, :C] * sin)
    new = TF.layer_norm(new, (C,), p['norm.weight'], p['norm.bias'], 1e-6)
    loc = TF.layer_norm(local, (C,), p['norm_local.weight'], p['norm_local.bias'], 1e-6)
    return torch.relu(new + loc)


# ------------------------------------------------------------------ spconv-free detection layers
def _dense_from(sct, dev):
    return sct.dense()          # [B, C, D, H, W]


@pytest.mark.parametrize('kind', ['subm', 'k3s2p1', 'k3s2p011', 'k311s211'])
def test_spconv_layers_vs_dense_conv3d(dev, kind):
    """SubMConv3d / SparseConv3d restated without spconv: checked against dense
    torch.nn.functional.conv3d on the densified input (active-site rule and values)."""
    import torch.nn.functional as TF
    from link_b200.scn import SparseConv3d, SparseConvTensor, SubMConv3d
    rng = np.random.default_rng(7)
    B, D, H, W, Cin, Cout = 2, 11, 24, 20, 8, 12
    occ = rng.random((B, D, H, W)) < 0.08
    idx = np.argwhere(occ).astype(np.int32)                      # (b, z, y, x)
    idx = idx[rng.permutation(len(idx))]
    feats = rng.standard_normal((len(idx), Cin)).astype(np.float32)
    x = SparseConvTensor(cu(feats, dev), cu(idx, dev), [D, H, W], B)
    torch.manual_seed(0)
    if kind == 'subm':
        m, stride, pad = SubMConv3d(Cin, Cout, 3, padding=1, bias=True, indice_key='a'), (1, 1, 1), (1, 1, 1)
    elif kind == 'k3s2p1':
        m, stride, pad = SparseConv3d(Cin, Cout, 3, 2, padding=1, bias=False), (2, 2, 2), (1, 1, 1)
    elif kind == 'k3s2p011':
        m, stride, pad = SparseConv3d(Cin, Cout, 3, 2, padding=[0, 1, 1], bias=False), (2, 2, 2), (0, 1, 1)
    else:
        m, stride, pad = SparseConv3d(Cin, Cout, (3, 1, 1), (2, 1, 1), bias=False), (2, 1, 1), (0, 0, 0)
    m = m.to(dev)
    with torch.no_grad():
        y = m(x)
    # dense oracle: weight [Cout, kz, ky, kx, Cin] -> conv3d layout [Cout, Cin, kz, ky, kx]
    w = m.weight.detach().permute(0, 4, 1, 2, 3).contiguous().cpu()
    dense_in = x.dense().cpu()
    ref = TF.conv3d(dense_in, w, m.bias.detach().cpu() if m.bias is not None else None, stride=stride, padding=pad)
    occ_t = torch.from_numpy(occ[:, None].astype(np.float32))
    reach = TF.conv3d(occ_t, torch.ones(1, 1, *m.kernel_size), stride=stride, padding=pad)[:, 0] > 0
    yi = y.indices.cpu().long()
    if kind == 'subm':
        assert torch.equal(y.indices, x.indices)
    else:
        want_sites = torch.nonzero(reach)                                     # sorted (b,z,y,x)
        assert list(y.spatial_shape) == list(ref.shape[2:])
        assert torch.equal(yi, want_sites)
    got = y.features.cpu()
    want = ref[yi[:, 0], :, yi[:, 1], yi[:, 2], yi[:, 3]]
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=RTOL, atol=1e-5)


def test_det_backbone_forward(dev):
    """SpMiddleResNetFHDELKv3: nuScenes-style call signature and output shapes; the fused
    inference path (BN folded into conv epilogues, native LinK executor) agrees with the unfused
    module-by-module path."""
    from link_b200.scn import SpMiddleResNetFHDELKv3
    rng = np.random.default_rng(3)
    B, D, H, W = 2, 40, 96, 96                                   # input_shape is (x, y, z)
    occ = rng.random((B, D, H, W)) < 0.01
    occ[:, 20:] = False
    idx = np.argwhere(occ).astype(np.int32)
    idx = idx[rng.permutation(len(idx))]
    feats = rng.standard_normal((len(idx), 5)).astype(np.float32)
    torch.manual_seed(0)
    net = SpMiddleResNetFHDELKv3(num_input_features=5, ds_factor=8).to(dev).eval()
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.uniform_(-0.1, 0.1)
                mod.running_var.uniform_(0.8, 1.2)
        dense, multi = net(cu(feats, dev), cu(idx, dev), B, [W, H, D])
    assert dense.shape == (B, 128 * 2, H // 8, W // 8)
    assert set(multi) == {'conv1', 'conv2', 'conv3', 'conv4'}
    assert [multi[k].features.shape[1] for k in ('conv1', 'conv2', 'conv3', 'conv4')] == [16, 32, 64, 128]
    assert list(multi['conv4'].spatial_shape) == [5, H // 8, W // 8]
    assert torch.isfinite(dense).all()
    # unfused path: autograd enabled -> module-by-module execution, composed LinK blocks
    for p in net.parameters():
        p.requires_grad_(False)
    f2 = cu(feats, dev).requires_grad_(True)
    dense2, _ = net(f2, cu(idx, dev), B, [W, H, D])
    np.testing.assert_allclose(dense.cpu().numpy(), dense2.detach().cpu().numpy(), rtol=2e-3, atol=2e-4)
    dense2.square().mean().backward()
    assert torch.isfinite(f2.grad).all() and float(f2.grad.abs().sum()) > 0


def test_centerpoint_detector_training_step(dev):
    """BASELINE config 4 end to end at a small grid: voxels -> VoxelFeatureExtractorV3 ->
    SpMiddleResNetFHDELKv3 (sparse convs + LinK blocks on the CUDA library) -> RPN -> CenterHead ->
    focal + regression losses, forward + backward in training mode: finite losses for the six task
    groups, finite non-zero gradients in the head, the neck and the sparse backbone."""
    from link_b200.centerpoint import NUSC_TASKS, build_nusc_centerpoint
    rng = np.random.default_rng(5)
    B, D, H, W, P = 2, 40, 96, 96, 10
    occ = rng.random((B, D, H, W)) < 0.01
    occ[:, 20:] = False
    idx = np.argwhere(occ).astype(np.int32)
    idx = idx[rng.permutation(len(idx))]
    nv = len(idx)
    num = rng.integers(1, P + 1, nv).astype(np.int32)
    voxels = rng.standard_normal((nv, P, 5)).astype(np.float32) * (np.arange(P)[None, :, None] < num[:, None, None])
    torch.manual_seed(0)
    model = build_nusc_centerpoint().to(dev).train()
    hw, max_objs = (H // 8) * (W // 8), 16
    g = torch.Generator().manual_seed(1)
    example = {'voxels': cu(voxels.astype(np.float32), dev), 'num_points': cu(num, dev), 'coordinates': cu(idx, dev),
               'batch_size': B, 'shape': [np.array([W, H, D])] * B,
               'hm': [], 'ind': [], 'mask': [], 'cat': [], 'anno_box': []}
    for t in NUSC_TASKS:
        n_cls = len(t['class_names'])
        example['hm'].append((torch.rand(B, n_cls, H // 8, W // 8, generator=g) ** 4).to(dev))
        example['ind'].append(torch.randint(0, hw, (B, max_objs), generator=g).to(dev))
        example['mask'].append((torch.rand(B, max_objs, generator=g) < 0.5).to(torch.uint8).to(dev))
        example['cat'].append(torch.randint(0, n_cls, (B, max_objs), generator=g).to(dev))
        example['anno_box'].append(torch.randn(B, max_objs, 10, generator=g).to(dev))
    losses = model(example, return_loss=True)
    assert len(losses['loss']) == len(NUSC_TASKS)
    total = sum(losses['loss'])
    assert torch.isfinite(total)
    total.backward()
    for name in ('bbox_head.tasks.0.hm.0.weight', 'bbox_head.shared_conv.0.weight', 'neck.blocks.0.1.weight'):
        gr = dict(model.named_parameters())[name].grad
        assert gr is not None and torch.isfinite(gr).all() and float(gr.abs().sum()) > 0, name
    bk = [p.grad for p in model.backbone.parameters() if p.grad is not None]
    assert len(bk) > 10 and all(torch.isfinite(x).all() for x in bk) and sum(float(x.abs().sum()) for x in bk) > 0
    with torch.no_grad():
        preds = model.eval()(example, return_loss=False)
    assert len(preds) == len(NUSC_TASKS) and preds[0]['hm'].shape == (B, 1, H // 8, W // 8)


def test_encoder_training_step(dev):
    """fwd + bwd + SGD through ELKEncoder in training mode (BatchNorm batch statistics, composed
    LinK blocks, sparse-conv backward kernels): finite gradients on every parameter the forward
    uses, decoder branches untouched (they are unused by ELKEncoder.forward), loss goes down."""
    from link_b200 import SparseTensor
    from link_b200.linkencoder import ELKEncoder
    from link_b200.utils.synthetic import random_voxels
    coords = cu(random_voxels(3000, 40, seed=4, batch=2), dev)
    torch.manual_seed(4)
    feats = torch.randn(coords.shape[0], 4, device=dev)
    target = torch.randint(0, 19, (coords.shape[0],), device=dev)
    net = ELKEncoder(num_classes=19, cr=0.25, baseop='cos', r=3, s=3, groups=2).to(dev).train()
    opt = torch.optim.SGD([p for n, p in net.named_parameters() if not n.startswith('up')], lr=0.05)
    losses = []
    for it in range(4):
        opt.zero_grad(set_to_none=True)
        logits = net(SparseTensor(feats.clone(), coords, 1))
        loss = torch.nn.functional.cross_entropy(logits, target)
        loss.backward()
        if it == 0:
            for n, p in net.named_parameters():
                if n.startswith('up'):
                    assert p.grad is None, n
                else:
                    assert p.grad is not None and torch.isfinite(p.grad).all(), n
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses


def test_reference_cuda_kernels_vs_oracle_and_ours(dev):
    """The reference GPU arm of bench.py (the reference's own CUDA kernels from oracle/_ref/backend_cuda.so
    under the restated python glue) computes the same ELKBlock forward as the CPU oracle and as ours."""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip('oracle/_ref/backend_cuda.so not built (needs /root/reference at build time)')
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock
    from link_b200.utils.synthetic import random_voxels
    for baseop, groups, C, s, r in [('cos', 2, 64, 7, 3), ('cos_x', 1, 16, 3, 2)]:
        coords = random_voxels(6000, 48, seed=7)
        torch.manual_seed(1)
        blk = ELKBlock(C, C, groups=groups, baseop=baseop).eval()
        feats = torch.randn(len(coords), C)
        sd = {k: v.detach() for k, v in blk.state_dict().items()}
        want = O.elk_block_forward(feats, coords, 1, sd, s, r, baseop, groups).numpy()
        blk = blk.to(dev)
        with torch.no_grad():
            ref = ref_gpu.elk_block_forward(feats.to(dev), cu(coords, dev), 1,
                                            {k: v.detach() for k, v in blk.state_dict().items()}, s, r, baseop, groups)
            ours = blk(SparseTensor(feats.to(dev), cu(coords, dev), 1), s, r).F
        np.testing.assert_allclose(ref.cpu().numpy(), want, rtol=1e-3, atol=2e-4)
        np.testing.assert_allclose(ours.cpu().numpy(), ref.cpu().numpy(), rtol=1e-3, atol=2e-4)


def test_from_host_async_upload_matches_resident_inputs(dev):
    """SparseTensor.from_host (coordinates on the current stream, features on a copy stream, joined on
    the device after the index-only kernels) gives the same block and encoder outputs (up to the
    run-to-run order of the float reductions that join block sums spanning two lane groups)."""
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock
    from link_b200.linkencoder import ELKEncoder
    from link_b200.utils.synthetic import kitti_like_voxels
    c3, f4 = kitti_like_voxels(30_000, seed=5)
    coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    ch = torch.from_numpy(coords).pin_memory()
    torch.manual_seed(0)
    blk = ELKBlock(64, 64, groups=2, baseop='cos').to(dev).eval()
    fh = torch.randn(len(coords), 64).pin_memory()
    enc = ELKEncoder(num_classes=19, cr=1.0, baseop='cos', r=3, s=7, groups=2).to(dev).eval()
    f4h = torch.from_numpy(f4.astype(np.float32)).pin_memory()
    with torch.no_grad():
        for _ in range(3):       # repeated: the upload buffer recycles memory of earlier steps
            want = blk(SparseTensor(fh.to(dev), ch.to(dev), 1), 7, 3).F
            got = blk(SparseTensor.from_host(fh, ch, 1, device=dev), 7, 3).F
            np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-5)
        want = enc(SparseTensor(f4h.to(dev), ch.to(dev), 1))
        got = enc(SparseTensor.from_host(f4h, ch, 1, device=dev))
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=1e-5)


def test_from_host_back_to_back_uploads_do_not_race(dev):
    """Ten from_host + block steps issued without any host synchronisation, with DIFFERENT features
    per step and the results kept on the device: the upload buffers recycle memory of earlier
    steps while those may still be queued, so every step must equal the resident-input result of
    its own features."""
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock
    from link_b200.utils.synthetic import kitti_like_voxels
    c3, _ = kitti_like_voxels(30_000, seed=7)
    coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    ch = torch.from_numpy(coords).pin_memory()
    cd = ch.to(dev)
    torch.manual_seed(1)
    blk = ELKBlock(64, 64, groups=2, baseop='cos').to(dev).eval()
    g = torch.Generator().manual_seed(3)
    fhs = [torch.randn(len(coords), 64, generator=g).pin_memory() for _ in range(4)]
    with torch.no_grad():
        wants = [blk(SparseTensor(f.to(dev), cd.clone(), 1), 7, 3).F.clone() for f in fhs]
        torch.cuda.synchronize()
        gots = []
        for k in range(10):
            gots.append(blk(SparseTensor.from_host(fhs[k % 4], ch, 1, device=dev), 7, 3).F.sum(dim=0))
        torch.cuda.synchronize()
    for k, got in enumerate(gots):
        want = wants[k % 4].sum(dim=0)
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=1e-3)


def test_upload_ring_back_to_back(dev):
    """UploadRing: twelve scans of different sizes and features streamed through a depth-2 ring with
    no host synchronisation; every result equals the resident-input result of its own scan."""
    from link_b200 import SparseTensor
    from link_b200.tensor import UploadRing
    from link_b200.elk import ELKBlock
    from link_b200.utils.synthetic import kitti_like_voxels
    scans = []
    g = torch.Generator().manual_seed(3)
    for seed, nv in ((7, 30_000), (8, 22_000), (9, 27_000)):
        c3, _ = kitti_like_voxels(nv, seed=seed)
        coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
        scans.append((torch.from_numpy(coords).pin_memory(), torch.randn(len(coords), 64, generator=g).pin_memory()))
    torch.manual_seed(1)
    blk = ELKBlock(64, 64, groups=2, baseop='cos').to(dev).eval()
    ring = UploadRing(max(len(c) for c, _ in scans), 64, device=dev, depth=2)
    with torch.no_grad():
        wants = [blk(SparseTensor(f.to(dev), c.to(dev), 1), 7, 3).F.sum(dim=0) for c, f in scans]
        torch.cuda.synchronize()
        gots = []
        for k in range(12):
            c, f = scans[k % 3]
            gots.append(blk(ring.upload(f, c, 1), 7, 3).F.sum(dim=0))
        torch.cuda.synchronize()
    for k, got in enumerate(gots):
        np.testing.assert_allclose(got.cpu().numpy(), wants[k % 3].cpu().numpy(), rtol=1e-4, atol=1e-3)


def test_fused_index_entry_points_match_their_parts(dev):
    """lk_sort_unique_coords == lk_pack_keys + lk_sort_unique_ex, lk_table_build_coords == lk_hash +
    lk_table_build, lk_block_neighbors_zero / lk_link_window_mean_seg == their unfused forms."""
    import ctypes as Ct
    from link_b200 import _capi
    from link_b200.nn.functional import _index
    from link_b200.nn.functional.hash import sphash
    from link_b200.nn.functional.query import HashTable
    from link_b200.nn.utils import get_kernel_offsets
    from link_b200.utils.synthetic import random_voxels
    L, st = _capi.lib(), _capi.stream()
    for n in (1, 2049, 30_000):
        coords_h = random_voxels(n, 60, seed=n, batch=2)
        coords = cu(coords_h, dev)
        n2 = coords.shape[0]
        spec, bits = _index.make_keyspec(_index.coord_bounds(coords), (3, 3, 3), (0, 1, 2, 3))
        a = _index.sort_unique(_index.pack_keys(coords, spec), bits, want_order=True)
        b_unique = torch.empty(n2, dtype=torch.int64, device=dev)
        b_inv, b_ord, b_srank = (torch.empty(n2, dtype=torch.int32, device=dev) for _ in range(3))
        b_seg = torch.empty(n2 + 1, dtype=torch.int32, device=dev)
        b_num = torch.empty(1, dtype=torch.int32, device=dev)
        ws_bytes = L.lk_sort_unique_ws_bytes(n2)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _capi.check(L.lk_sort_unique_coords(_capi.ptr(coords), Ct.byref(spec), n2, bits, _capi.ptr(b_unique),
                                            _capi.ptr(b_inv), _capi.ptr(b_ord), _capi.ptr(b_seg), None,
                                            _capi.ptr(b_num), _capi.ptr(b_srank), _capi.ptr(ws), ws_bytes, st),
                    'lk_sort_unique_coords')
        m = int(a.num.item())
        assert int(b_num.item()) == m
        assert torch.equal(b_unique[:m], a.unique[:m]) and torch.equal(b_inv, a.inverse)
        assert torch.equal(b_ord, a.order) and torch.equal(b_srank, a.sorted_rank)
        assert torch.equal(b_seg[:m + 1], a.seg[:m + 1])
        # table from coordinates == table from hashes
        t_ref = HashTable(sphash(coords))
        tab = torch.empty(t_ref.capacity * 16, dtype=torch.uint8, device=dev)
        _capi.check(L.lk_table_build_coords(_capi.ptr(coords), n2, _capi.ptr(tab), t_ref.capacity, st), 'tbc')
        # (slot placement under collisions depends on insertion order: compare look-ups, not bytes)
        q = torch.cat([sphash(coords), sphash(coords + 1000)])
        got = torch.empty_like(q)
        _capi.check(L.lk_table_query(_capi.ptr(q), q.shape[0], _capi.ptr(tab), t_ref.capacity, _capi.ptr(got), st), 'tq')
        assert torch.equal(got, t_ref.query(q)) and torch.equal(got[:n2], torch.arange(n2, device=dev))
        # neighbours + zeroing, window mean from segment starts
        offs = get_kernel_offsets(3, 1, 1, device=dev)
        nbr_a = torch.full((n2, 27), -7, dtype=torch.int32, device=dev)
        nbr_b = torch.full((n2, 27), -7, dtype=torch.int32, device=dev)
        zbuf = torch.ones(n2, 8, device=dev)
        _capi.check(L.lk_block_neighbors(_capi.ptr(a.unique), _capi.ptr(a.num), n2, Ct.byref(spec), _capi.ptr(offs),
                                         27, _capi.ptr(nbr_a), st), 'bn')
        _capi.check(L.lk_block_neighbors_zero(_capi.ptr(a.unique), _capi.ptr(a.num), n2, Ct.byref(spec),
                                              _capi.ptr(offs), 27, _capi.ptr(nbr_b), _capi.ptr(zbuf), 8, st), 'bnz')
        assert torch.equal(nbr_a, nbr_b)
        assert float(zbuf[:m].abs().sum()) == 0.0 and (m == n2 or float(zbuf[m:].min()) == 1.0)
        sums = torch.randn(n2, 8, device=dev)
        mean_a, mean_b = torch.empty(n2, 8, device=dev), torch.empty(n2, 8, device=dev)
        _capi.check(L.lk_link_window_mean(_capi.ptr(sums), _capi.ptr(a.counts), _capi.ptr(nbr_a), _capi.ptr(a.num),
                                          n2, 27, 8, _capi.ptr(mean_a), st), 'wm')
        _capi.check(L.lk_link_window_mean_seg(_capi.ptr(sums), _capi.ptr(a.seg), _capi.ptr(nbr_a), _capi.ptr(a.num),
                                              n2, 27, 8, _capi.ptr(mean_b), st), 'wms')
        assert torch.equal(mean_a[:m], mean_b[:m])


def test_points_to_voxel_bit_exact(dev):
    """GPU points_to_voxel == the reference's numba loop (golden fixture) and == the oracle on a
    larger cloud with rejected points, the voxel cap and the per-voxel point cap all active."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_points_golden import cloud
    from link_b200.ops import points_to_voxel
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'points.npz'))
    pts = cloud(seed=7)
    for tag in ('a', 'b'):
        mp, mv = (int(x) for x in g[f'{tag}_cfg'])
        v, c, n = points_to_voxel(cu(pts, dev), [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], mp, True, mv)
        assert np.array_equal(c.cpu().numpy(), g[f'{tag}_coors']) and np.array_equal(n.cpu().numpy(), g[f'{tag}_num'])
        vh = v.cpu().numpy()
        assert np.array_equal(vh[:len(g[f'{tag}_voxels'])], g[f'{tag}_voxels'])
        assert np.array_equal(vh.astype(np.float64).sum(axis=(1, 2)), g[f'{tag}_voxel_sum'])
    big = cloud(seed=11, n=300_000)
    for mp, mv, rev in [(10, 120_000, True), (3, 500, False), (1, 1, True)]:
        want = O.points_to_voxel(big, [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], mp, rev, mv)
        got = points_to_voxel(cu(big, dev), [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], mp, rev, mv)
        for a, b in zip(got, want):
            assert np.array_equal(a.cpu().numpy(), b)
    # every point rejected / empty input
    far = np.full((100, 5), 1000.0, np.float32)
    v, c, n = points_to_voxel(cu(far, dev), [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, True, 100)
    assert v.shape[0] == 0 and c.shape[0] == 0 and n.shape[0] == 0
    v, c, n = points_to_voxel(torch.zeros(0, 5, device=dev), [0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3], 10, True, 100)
    assert v.shape[0] == 0


# ------------------------------------------------------------------ SURVEY 8f row 4 (last: first GPU run pending)
def test_rotated_iou_and_predict_on_device(dev):
    """SURVEY §8f row 4 on the GPU: rotated BEV IoU against the reference-generated fixture, rotated NMS
    (selection may differ from the CPU-generated one only through IoUs within float noise of the
    threshold), and CenterHead.predict end to end with both NMS flavours on device tensors."""
    import os
    from link_b200.centerpoint import CenterHead, NUSC_COMMON_HEADS, NUSC_TASKS, NUSC_CODE_WEIGHTS
    from link_b200.iou3d import boxes_iou_bev, rotate_nms_pcdet
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'iou3d.npz')))
    iou = boxes_iou_bev(cu(g['a'], dev), cu(g['b'], dev))
    np.testing.assert_allclose(iou.cpu().numpy(), g['iou'], rtol=1e-3, atol=1e-4)
    sel = rotate_nms_pcdet(cu(g['nms_boxes'], dev), cu(g['nms_scores'], dev), 0.2, pre_maxsize=300, post_max_size=83)
    want = set(g['nms_t02'].tolist())
    assert sel.device.type == 'cuda' and len(set(sel.tolist()) & want) >= 0.95 * len(want)
    torch.manual_seed(0)
    head = CenterHead(in_channels=32, tasks=NUSC_TASKS, code_weights=list(NUSC_CODE_WEIGHTS),
                      common_heads=dict(NUSC_COMMON_HEADS), share_conv_channel=16).to(dev).eval()
    cfg = dict(post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], min_radius=[4, 12, 10, 1, 0.85, 0.175],
               nms=dict(nms_pre_max_size=1000, nms_post_max_size=83, nms_iou_threshold=0.2), score_threshold=0.1,
               pc_range=[-54, -54], out_size_factor=8, voxel_size=[0.075, 0.075])
    with torch.no_grad():
        preds, _ = head(torch.randn(2, 32, 24, 24, device=dev))
        for circular in (True, False):
            dets = head.predict({}, [dict(p) for p in preds], dict(cfg, circular_nms=circular))
            assert len(dets) == 2
            for d in dets:
                assert d['box3d_lidar'].device.type == 'cuda' and d['box3d_lidar'].shape[1] == 9
                assert d['scores'].shape == d['label_preds'].shape and len(d['scores']) <= 83 * len(NUSC_TASKS)
                assert torch.isfinite(d['box3d_lidar']).all()
                assert len(d['label_preds']) == 0 or 0 <= int(d['label_preds'].min()) <= int(d['label_preds'].max()) <= 9
