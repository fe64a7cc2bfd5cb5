"""Edge cases of the hot path through the C ABI / the python surface, against the oracle: empty and
single-voxel inputs, ragged batches (frames of very different sizes), negative and large
coordinates, tile-boundary sizes of the tensor-core conv, duplicate query keys, and a maximum-size
scan (500k voxels) through size-independent properties."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import link_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 2e-4


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from link_b200 import _capi
    _capi.lib()
    return torch.device('cuda:0')


def cu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _block(C_, groups, op, seed=0):
    from link_b200.elk import ELKBlock
    torch.manual_seed(seed)
    blk = ELKBlock(C_, C_, groups=groups, baseop=op).eval()
    with torch.no_grad():
        for m in (blk.pre_mix[1], blk.norm, blk.norm_local):
            m.weight.uniform_(0.5, 1.5)
            m.bias.uniform_(-0.5, 0.5)
    return blk


def _block_vs_oracle(dev, coords, C_=32, groups=2, op='cos', s=5, r=3, atol=4e-5):
    from link_b200 import SparseTensor
    blk = _block(C_, groups, op, seed=len(coords))
    feats = torch.randn(len(coords), C_, generator=torch.Generator().manual_seed(1))
    want = O.elk_block_forward(feats, coords, 1, {k: v.detach() for k, v in blk.state_dict().items()}, s, r, op, groups)
    with torch.no_grad():
        got = blk.to(dev)(SparseTensor(feats.to(dev), cu(coords, dev), 1), s, r).F
    np.testing.assert_allclose(got.cpu().numpy(), want.detach().numpy(), rtol=RTOL, atol=atol)


def test_empty_inputs_through_the_c_abi(dev):
    """n = 0 is a no-op that returns LK_OK for the index and feature entry points."""
    from link_b200 import _capi
    L, st = _capi.lib(), _capi.stream()
    e32 = torch.empty(0, 4, dtype=torch.int32, device=dev)
    e64 = torch.empty(0, dtype=torch.int64, device=dev)
    ef = torch.empty(0, 32, device=dev)
    assert L.lk_hash(e32.data_ptr(), 0, e64.data_ptr(), st) == 0
    assert L.lk_count(e32.data_ptr(), 0, torch.zeros(4, dtype=torch.int32, device=dev).data_ptr(), 4, st) == 0
    ep = _capi.ConvEpilogue()
    assert L.lk_conv_tc_fwd_plan(ef.data_ptr(), ef.data_ptr(), e32.data_ptr(), None, None, 0, 27, 32, 32, C.byref(ep),
                                 ef.data_ptr(), st) == 0
    a = _capi.ElkBlockArgs()
    a.n = 0
    a.d_coords = a.d_feats = a.d_out = a.d_ws = a.d_kmap = 1      # non-null, never dereferenced
    assert L.lk_elk_block_fwd(C.byref(a), st) == 0
    enc = _capi.ElkEncoderArgs()
    enc.n0, enc.levels = 0, 4
    assert L.lk_elk_encoder_fwd(C.byref(enc), st) == 0
    torch.cuda.synchronize()


def test_empty_sparse_tensor_through_the_modules(dev):
    """An empty scan (a frame without points after range filtering) passes through conv and query ops."""
    import link_b200.nn.functional as F
    from link_b200 import SparseTensor
    coords = torch.empty(0, 4, dtype=torch.int32, device=dev)
    x = SparseTensor(torch.empty(0, 16, device=dev), coords, 1)
    y = F.conv3d(x, torch.randn(27, 16, 32, device=dev), 3)
    assert y.F.shape == (0, 32)
    assert F.sphash(coords).shape == (0,)
    got = F.sphashquery(F.sphash(coords), F.sphash(cu(np.array([[1, 2, 3, 0]], np.int32), dev)))
    assert got.shape == (0,)


@pytest.mark.parametrize('n', [1, 2, 31, 127, 128, 129, 255, 257])
def test_tiny_and_tile_boundary_scans(dev, n):
    """Single voxels and sizes around the 128-row tiles of the tensor-core kernels: whole block vs oracle."""
    from link_b200.utils.synthetic import random_voxels
    coords = random_voxels(n, 12, seed=n, batch=1)[:n]
    assert len(coords) == n
    _block_vs_oracle(dev, coords)


def test_ragged_batch(dev):
    """Frames of very different sizes in one batch (1, 40 and 6000 voxels): blocks, kernel maps and
    windows never cross frames (the batch index is the 4th hashed coordinate)."""
    from link_b200.utils.synthetic import random_voxels
    parts = []
    for b, n in enumerate((1, 40, 6000)):
        c = random_voxels(n, 30, seed=10 + b, batch=1)[:n].copy()
        c[:, 3] = b
        parts.append(c)
    coords = np.concatenate(parts).astype(np.int32)
    coords = coords[np.random.default_rng(0).permutation(len(coords))]     # frames interleaved in memory
    _block_vs_oracle(dev, coords, C_=64, groups=2, op='cos', s=7, r=3)
    # each frame alone gives the same rows as inside the batch
    from link_b200 import SparseTensor
    blk = _block(32, 2, 'cos', seed=3).to(dev)
    feats = torch.randn(len(coords), 32, device=dev)
    with torch.no_grad():
        full = blk(SparseTensor(feats.clone(), cu(coords, dev), 1), 5, 3).F
        for b in range(3):
            sel = torch.from_numpy(coords[:, 3] == b).to(dev)
            alone = blk(SparseTensor(feats[sel].clone(), cu(coords[coords[:, 3] == b], dev), 1), 5, 3).F
            np.testing.assert_allclose(alone.cpu().numpy(), full[sel].cpu().numpy(), rtol=RTOL, atol=4e-5)


@pytest.mark.parametrize('shift', [(-37, -5, -1024), (30000, -30000, 12), (-(2 ** 20), 2 ** 20, 0)])
def test_negative_and_large_coordinates(dev, shift):
    """Coordinates far from the origin and below zero (floor division of negative numbers in the block
    keys; the hash is over the raw int32 fields): hashes bit-exact, block vs oracle.  The phase
    argument grows with |coordinate|, so the float tolerance is scaled by the phase ulp."""
    import link_b200.nn.functional as F
    from link_b200.utils.synthetic import random_voxels
    coords = random_voxels(3000, 40, seed=7, batch=2).copy()
    coords[:, :3] += np.array(shift, np.int32)
    assert np.array_equal(F.sphash(cu(coords, dev)).cpu().numpy(), O.sphash(coords))
    big = float(np.abs(coords[:, :3]).max())
    _block_vs_oracle(dev, coords, C_=32, groups=1, op='cos', s=7, r=3, atol=max(4e-5, 2e-6 * big))


def test_duplicate_and_missing_query_keys(dev):
    """sphashquery: duplicate queries all resolve, keys absent from the reference set give -1 (the
    reference's cuckoo table returns 0 -> -1 after the decrement, query.py:20-25)."""
    import link_b200.nn.functional as F
    rng = np.random.default_rng(5)
    ref = np.unique(rng.integers(-50, 50, size=(5000, 4)).astype(np.int32), axis=0)
    q = np.concatenate([ref[rng.integers(0, len(ref), 4000)], ref[:100], ref[:100],
                        (ref[:500] + np.array([1000, 0, 0, 0], np.int32))]).astype(np.int32)
    got = F.sphashquery(F.sphash(cu(q, dev)), F.sphash(cu(ref, dev))).cpu().numpy()
    want = O.sphashquery(O.sphash(q), O.sphash(ref))
    assert np.array_equal(got, want)
    assert (got[-500:] == -1).all() and (got[:-500] >= 0).all()


def test_maximum_size_scan_properties(dev):
    """~480k voxels in one frame (4x the benchmark scan), C = 64, cos (3x7)^3: finite output, the submanifold kernel map
    is symmetric (nbr[k][o] = i  <=>  nbr[K-1-k][i] = o), every voxel finds itself at the centre offset,
    the block partition covers every voxel once, and two runs agree to the float-atomic noise."""
    import link_b200.elk as elk
    import link_b200.nn.functional as F
    from link_b200 import SparseTensor
    from link_b200.utils.synthetic import kitti_like_voxels
    c3, _ = kitti_like_voxels(120_000, seed=0)                       # the benchmark scan ...
    step = int(c3[:, 0].max() - c3[:, 0].min()) + 3                  # ... four times side by side (touching edges)
    c3 = np.concatenate([c3 + np.array([k * step, 0, 0], np.int32) for k in range(4)])
    coords = cu(np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32), dev)
    n = coords.shape[0]
    assert n > 400_000
    blk = _block(64, 2, 'cos').to(dev)
    feats = torch.randn(n, 64, device=dev)
    st = SparseTensor(feats.clone(), coords, 1)
    with torch.no_grad():
        a = blk(st, 7, 3).F
        b = blk(SparseTensor(feats.clone(), coords, 1), 7, 3).F
    assert torch.isfinite(a).all()
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=2e-6)
    km = st.kmaps[((1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))]
    nbr = km.nbr
    K = nbr.shape[0]
    assert torch.equal(nbr[K // 2], torch.arange(n, dtype=torch.int32, device=dev))
    for k in (0, 5, 12):
        o = torch.nonzero(nbr[k] >= 0).squeeze(1)
        i = nbr[k][o].long()
        assert torch.equal(nbr[K - 1 - k][i], o.int())
    bi = elk.block_index(SparseTensor(feats, coords, 1), 7)
    assert int(bi.counts[:bi.m].sum()) == n
    assert int(bi.idx_query.min()) == 0 and int(bi.idx_query.max()) == bi.m - 1
