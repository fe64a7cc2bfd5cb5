"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol declared in include/linkb200.h; host logic (offset tables, key layout, quantise,
collate) matches the reference-generated fixtures.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'linkb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lk_[a-z0-9_]+)\s*\(', text)))


def test_library_loads_and_exports_every_declared_symbol():
    from link_b200 import _capi
    lib = _capi.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/linkb200.h but not exported'
        assert n in _capi.PROTOTYPES, f'{n} has no ctypes prototype'
    assert set(_capi.PROTOTYPES) == set(names)
    assert lib.lk_version() >= 100
    assert lib.lk_table_capacity(1000) == 2048
    assert lib.lk_sort_unique_ws_bytes(1 << 20) > (1 << 20) * 24


def test_missing_gpu_fails_loudly():
    import link_b200.nn.functional as F
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        F.sphash(torch.zeros(4, 4, dtype=torch.int))


def test_struct_layout_matches_header():
    from link_b200 import _capi
    assert ctypes.sizeof(_capi.KeySpec) == 4 * (3 + 3 + 4 + 4 + 4)
    assert ctypes.sizeof(_capi.KernelGen) == 40
    assert _capi.KernelGen.d_pos_weight.offset == 16 and _capi.KernelGen.d_alpha.offset == 24


def test_kernel_offsets_match_reference():
    from link_b200.nn.utils import get_kernel_offsets
    g = load_golden('kat')
    assert np.array_equal(get_kernel_offsets(2).numpy(), g['off2'])
    assert np.array_equal(get_kernel_offsets(3).numpy(), g['off3'])
    assert np.array_equal(get_kernel_offsets(3, 4).numpy(), g['off3_s4'])
    assert np.array_equal(get_kernel_offsets((3, 1, 1)).numpy(), g['off311'])
    assert get_kernel_offsets(2).dtype == torch.int32


def test_keyspec_order_and_width():
    from link_b200.nn.functional._index import make_keyspec
    spec, bits = make_keyspec(((-7, 0, 3, 0), (100, 63, 40, 1)), (7, 7, 7), (0, 1, 2, 3))
    assert list(spec.lo) == [-1, 0, 0, 0]
    assert list(spec.bits) == [4, 4, 3, 1] and bits == 12     # q ranges: 15, 9, 5, 1
    spec, bits = make_keyspec(((0, 0, 0, 0), (0, 0, 0, 0)), (2, 2, 2), (3, 0, 1, 2))
    assert bits == 1 and list(spec.bits) == [0, 0, 0, 0]
    with pytest.raises(RuntimeError):
        make_keyspec(((-2 ** 30,) * 4, (2 ** 30,) * 4), (1, 1, 1), (0, 1, 2, 3))


def test_sparse_quantize_and_collate():
    from link_b200 import SparseTensor
    from link_b200.utils.collate import sparse_collate, sparse_collate_fn
    from link_b200.utils.quantize import sparse_quantize
    g = load_golden('voxelize')
    qc, qi, qv = sparse_quantize(g['pts'].copy(), 0.5, return_index=True, return_inverse=True)
    assert np.array_equal(qc, g['q_coords']) and np.array_equal(qi, g['q_index'])
    assert np.array_equal(qv, g['q_inverse'])
    a = SparseTensor(torch.ones(3, 2), torch.zeros(3, 3, dtype=torch.int))
    b = SparseTensor(torch.ones(2, 2), torch.ones(2, 3, dtype=torch.int))
    c = sparse_collate([a, b])
    assert c.C.shape == (5, 4) and c.C[:, 3].tolist() == [0, 0, 0, 1, 1] and c.F.shape == (5, 2)
    out = sparse_collate_fn([{'x': a, 'n': np.zeros(2)}, {'x': b, 'n': np.ones(2)}])
    assert out['x'].C.shape == (5, 4) and out['n'].shape == (2, 2)


def test_module_surface_and_state_dict_compat():
    """Same parameter names/shapes as the reference encoder (fixture made from its state dict)."""
    from link_b200.linkencoder import ELKEncoder
    g = load_golden('encoder_cosx_2x3')
    enc = ELKEncoder(num_classes=19, cr=0.25, baseop='cos_x', r=2, s=3, groups=1)
    ours = {k: tuple(v.shape) for k, v in enc.state_dict().items()
            if not k.startswith('up') and not k.endswith('num_batches_tracked')}
    ref = {k[3:]: tuple(v.shape) for k, v in g.items() if k.startswith('sd.')}
    assert ours == ref
    assert sum(p.numel() for p in ELKEncoder(num_classes=19, cr=1.0, baseop='cos', r=3, s=7,
                                             groups=2).parameters()) > 5_000_000


def test_unet_state_dict_compat():
    from link_b200.linkunet import ELKUNet
    g = load_golden('unet_cosx_2x3')
    net = ELKUNet(num_classes=19, cr=0.25, baseop='cos_x', r=2, s=3, groups=1)
    ours = {k: tuple(v.shape) for k, v in net.state_dict().items()
            if not k.endswith('num_batches_tracked')}
    ref = {k[3:]: tuple(v.shape) for k, v in g.items() if k.startswith('sd.')}
    assert ours == ref


def test_compat_shim_lets_reference_style_imports_resolve():
    import link_b200.compat as compat
    compat.install()
    import torchsparse
    import torchsparse.nn as spnn
    import torchsparse.nn.functional as F
    from torchsparse import PointTensor, SparseTensor  # noqa: F401
    from torchsparse.nn.utils import get_kernel_offsets  # noqa: F401
    from torchsparse.utils import make_ntuple  # noqa: F401
    from torchsparse.utils.collate import sparse_collate_fn  # noqa: F401
    from torchsparse.utils.quantize import sparse_quantize  # noqa: F401
    assert spnn.Conv3d is not None and hasattr(F, 'sphashquery') and hasattr(F, 'spdevoxelize')
    assert torchsparse.__version__
    compat.uninstall()


@pytest.mark.parametrize('n', [0, 1, 3, 4, 5, 1001])
def test_host_coord_bounds(n):
    """lk_host_coord_bounds (host code, no device): column-wise min / max, vector body + scalar tail."""
    from link_b200 import _capi
    c = torch.randint(-1000, 1000, (n, 4), dtype=torch.int32)
    lo, hi = (ctypes.c_int32 * 4)(), (ctypes.c_int32 * 4)()
    _capi.check(_capi.lib().lk_host_coord_bounds(c.data_ptr() if n else None, n, lo, hi), 'bounds')
    if n:
        assert list(lo) == c.min(0).values.tolist() and list(hi) == c.max(0).values.tolist()
    else:
        assert list(lo) == [0] * 4 and list(hi) == [0] * 4


def test_bind_host_to_gpu_is_safe_without_a_device():
    """sharding.bind_host_to_gpu: cpulist parsing, and no exception / no affinity change when the
    device (or its sysfs node) is not there."""
    import os
    from link_b200.sharding import bind_host_to_gpu, parse_cpulist
    assert parse_cpulist('0-3,8,10-11\n') == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist('5') == [5] and parse_cpulist('') == []
    before = os.sched_getaffinity(0)
    info = bind_host_to_gpu(0, sysfs='/nonexistent')
    assert info['bound'] is False
    assert os.sched_getaffinity(0) == before


def test_autograd_extension_binds_the_library():
    """link_b200/_ext.py: when the C++ autograd extension has been built (by __graft_entry__.build()), it
    loads on a machine without a GPU and accepts the addresses of the library's entry points."""
    from link_b200 import _ext
    if _ext._so_path() is None:
        pytest.skip('C++ autograd extension not built')
    m = _ext.module()
    assert m is not None and m.ready()
    with pytest.raises(RuntimeError):
        m.bind('lk_no_such_entry_point', 0)
