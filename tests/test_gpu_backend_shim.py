"""The drop-in boundary, call for call: every CUDA entry point of the reference's pybind module
`torchsparse.backend` (pybind_cuda.cpp:18-39) exists in link_b200.backend with the same signature, and
returns what the reference's OWN compiled CUDA backend (oracle/_ref/backend_cuda.so, built from
/root/reference by oracle/build_ref.py) returns for the same arguments.  Integer results bit-exact,
float results within fp32 accumulation-order noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def ref():
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip('oracle/_ref/backend_cuda.so not built (needs /root/reference at build time)')
    return ref_gpu.backend()


def _coords(n, extent, seed, batch=2):
    from link_b200.utils.synthetic import random_voxels
    return torch.from_numpy(random_voxels(n, extent, seed=seed, batch=batch))


def test_pybind_table_is_complete():
    import link_b200.backend as B
    for name in ['convolution_forward_cuda', 'convolution_backward_cuda', 'voxelize_forward_cuda',
                 'voxelize_backward_cuda', 'devoxelize_forward_cuda', 'devoxelize_backward_cuda', 'hash_cuda',
                 'kernel_hash_cuda', 'hash_query_cuda', 'count_cuda']:
        assert callable(getattr(B, name)), name


def test_hash_query_count_bit_exact(dev, ref):
    import link_b200.backend as B
    from link_b200.nn.utils import get_kernel_offsets
    c = _coords(5000, 40, 1).to(dev)
    assert torch.equal(B.hash_cuda(c), ref.hash_cuda(c))
    off = get_kernel_offsets(3, 2, device=dev)
    assert torch.equal(B.kernel_hash_cuda(c, off), ref.kernel_hash_cuda(c, off))
    h = ref.hash_cuda(c)
    q = torch.cat([h[::3], h[:50] + 12345])                 # hits and (almost surely) misses
    idx = torch.arange(h.numel(), device=dev)
    assert torch.equal(B.hash_query_cuda(q, h, idx), ref.hash_query_cuda(q, h, idx))
    cnt_idx = torch.randint(-1, 37, (4000,), device=dev, dtype=torch.int32)
    assert torch.equal(B.count_cuda(cnt_idx, 37), ref.count_cuda(cnt_idx, 37))


def test_voxelize_devoxelize_fwd_bwd(dev, ref):
    import link_b200.backend as B
    g = torch.Generator().manual_seed(2)
    n, m, c = 3000, 400, 24
    idx = torch.randint(0, m, (n,), generator=g, dtype=torch.int32).to(dev)
    counts = ref.count_cuda(idx, m)
    x = torch.randn(n, c, generator=g).to(dev)
    torch.testing.assert_close(B.voxelize_forward_cuda(x, idx, counts), ref.voxelize_forward_cuda(x, idx, counts),
                               rtol=1e-5, atol=1e-5)
    top = torch.randn(m, c, generator=g).to(dev)
    torch.testing.assert_close(B.voxelize_backward_cuda(top, idx, counts, n), ref.voxelize_backward_cuda(top, idx, counts, n),
                               rtol=1e-5, atol=1e-6)
    for r in (2, 3):
        ind = torch.randint(-1, m, (n, r ** 3), generator=g, dtype=torch.int32).to(dev)
        w = torch.rand(n, r ** 3, generator=g).to(dev)
        feat = torch.randn(m, c, generator=g).to(dev)
        torch.testing.assert_close(B.devoxelize_forward_cuda(feat, ind, w, r), ref.devoxelize_forward_cuda(feat, ind, w, r),
                                   rtol=1e-5, atol=1e-5)
        tg = torch.randn(n, c, generator=g).to(dev)
        torch.testing.assert_close(B.devoxelize_backward_cuda(tg, ind, w, m, r), ref.devoxelize_backward_cuda(tg, ind, w, m, r),
                                   rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('ksize,stride,cin,cout', [(3, 1, 16, 24), (2, 2, 32, 64), (3, 1, 64, 64)])
def test_convolution_forward_backward_on_reference_pair_lists(dev, ref, ksize, stride, cin, cout):
    """(neighbor_map, neighbor_offset) exactly as the reference's conv.py builds them (pair list ordered by
    offset, counts on the host): forward, transposed forward, and both backward products."""
    import link_b200.backend as B
    from link_b200 import SparseTensor
    from link_b200.nn.functional.conv import build_kernel_map
    c = _coords(4000, 30, 3).to(dev)
    st = SparseTensor(torch.zeros(c.shape[0], 1, device=dev), c, 1)
    km = build_kernel_map(st, (ksize,) * 3, (stride,) * 3, (1, 1, 1))
    nbmaps, nbsizes, (n_in, n_out) = km[0], km[1], km[2]
    nbmaps, nbsizes = nbmaps.int().contiguous(), nbsizes.int().cpu()
    g = torch.Generator().manual_seed(4)
    w = (torch.randn(ksize ** 3, cin, cout, generator=g) * 0.1).to(dev)
    for transpose in (False, True):
        ni, no = (n_in, n_out) if not transpose else (n_out, n_in)
        wk = w
        x = torch.randn(ni, cin, generator=g).to(dev)
        if transpose and stride == 1:
            continue                                           # the reference only transposes strided maps
        out_a = torch.zeros(no, cout, device=dev)
        out_b = torch.zeros(no, cout, device=dev)
        B.convolution_forward_cuda(x, out_a, wk, nbmaps, nbsizes, transpose)
        ref.convolution_forward_cuda(x, out_b, wk, nbmaps, nbsizes, transpose)
        torch.testing.assert_close(out_a, out_b, rtol=1e-4, atol=1e-4)
        go = torch.randn(no, cout, generator=g).to(dev)
        gi_a, gi_b = torch.zeros_like(x), torch.zeros_like(x)
        gw_a, gw_b = torch.zeros_like(wk), torch.zeros_like(wk)
        B.convolution_backward_cuda(x, gi_a, go, wk, gw_a, nbmaps, nbsizes, transpose)
        ref.convolution_backward_cuda(x, gi_b, go, wk, gw_b, nbmaps, nbsizes, transpose)
        torch.testing.assert_close(gi_a, gi_b, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(gw_a, gw_b, rtol=1e-3, atol=1e-3 * float(gw_b.abs().max()))


def test_reference_functional_layer_runs_on_the_shim(dev):
    """compat.install() registers link_b200.backend as `torchsparse.backend`."""
    import sys
    from link_b200 import compat
    compat.install()
    try:
        import torchsparse.backend as tb
        assert tb is sys.modules['link_b200.backend']
        c = _coords(100, 10, 5).to(dev)
        assert tb.hash_cuda(c).shape == (c.shape[0],)
    finally:
        compat.uninstall()
