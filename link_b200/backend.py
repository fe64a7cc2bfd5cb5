"""`torchsparse.backend` on liblinkb200: the reference's pybind table as a drop-in module.

The reference's Python layer reaches its kernels through 10 CUDA entry points of the pybind module
`torchsparse.backend` (torchsparse/backend/pybind_cuda.cpp:18-39).  This module defines those ten
functions with the SAME names, argument order, in/out conventions and return values, each a thin
binding of one or two C-ABI symbols of include/linkb200.h -- it is the file a maintainer drops in
as `torchsparse/backend.py`; `link_b200.compat.install()` registers it as
`torchsparse.backend`, so the reference's own nn/functional/*.py run on it unmodified.

The `*_cpu` variants do not exist: there is no CPU path (a CPU tensor raises).  fp32 only: the
reference's fp16 dispatch (under autocast) is served by converting at this boundary.
Tested call-for-call against the reference's own CUDA backend compiled from /root/reference
(tests/test_gpu_backend_shim.py)."""
import ctypes as C

import torch

from link_b200 import _capi

__all__ = ['hash_cuda', 'kernel_hash_cuda', 'hash_query_cuda', 'count_cuda', 'voxelize_forward_cuda',
           'voxelize_backward_cuda', 'devoxelize_forward_cuda', 'devoxelize_backward_cuda',
           'convolution_forward_cuda', 'convolution_backward_cuda']


def _f32(t):
    return t.contiguous().float()


def _i32(t):
    return t.contiguous().int()


def hash_cuda(idx):
    """hash_cuda(idx [N,4] int32) -> [N] int64 (hash_cuda.cu:57-73)."""
    idx = _i32(idx)
    out = torch.empty(idx.shape[0], dtype=torch.int64, device=idx.device)
    _capi.check(_capi.lib().lk_hash(_capi.ptr(idx), idx.shape[0], _capi.ptr(out), _capi.stream()), 'lk_hash')
    return out


def kernel_hash_cuda(idx, kernel_offset):
    """kernel_hash_cuda(idx [N,4], kernel_offset [K,3]) -> [K,N] int64 (hash_cuda.cu:75-93)."""
    idx, off = _i32(idx), _i32(kernel_offset)
    out = torch.empty(off.shape[0], idx.shape[0], dtype=torch.int64, device=idx.device)
    _capi.check(_capi.lib().lk_kernel_hash(_capi.ptr(idx), idx.shape[0], _capi.ptr(off), off.shape[0],
                                           _capi.ptr(out), _capi.stream()), 'lk_kernel_hash')
    return out


def hash_query_cuda(hash_query, hash_target, idx_target):
    """hash_query_cuda(query [Q], target [n], idx_target [n]) -> [Q] int64: idx_target of the matching
    target + 1, 0 for a miss (query_cuda.cu:9-58; query.py:32 subtracts the 1)."""
    q, ref = hash_query.contiguous(), hash_target.contiguous()
    L = _capi.lib()
    cap = int(L.lk_table_capacity(ref.numel()))
    table = torch.empty(cap * 16, dtype=torch.uint8, device=q.device)
    pos = torch.empty(q.numel(), dtype=torch.int64, device=q.device)
    _capi.check(L.lk_table_build(_capi.ptr(ref), ref.numel(), _capi.ptr(table), cap, _capi.stream()), 'lk_table_build')
    _capi.check(L.lk_table_query(_capi.ptr(q), q.numel(), _capi.ptr(table), cap, _capi.ptr(pos), _capi.stream()),
                'lk_table_query')
    # the table stores the POSITION of each target; the reference stores idx_target (an arange in every
    # caller, query.py:18-20) -- translate for the general case
    hit = pos >= 0
    vals = idx_target.contiguous().long()
    out = torch.where(hit, vals[pos.clamp(min=0)] + 1, torch.zeros_like(pos)) if vals.numel() else torch.zeros_like(pos)
    return out.view(hash_query.shape)


def count_cuda(idx, s):
    """count_cuda(idx [N] int32, s) -> [s] int32 histogram of the non-negative entries (count_cuda.cu:10-31)."""
    idx = _i32(idx)
    out = torch.empty(int(s), dtype=torch.int32, device=idx.device)
    _capi.check(_capi.lib().lk_count(_capi.ptr(idx), idx.numel(), _capi.ptr(out), int(s), _capi.stream()), 'lk_count')
    return out


def voxelize_forward_cuda(inputs, idx, counts):
    """voxelize_forward_cuda(inputs [N,c], idx [N] int32, counts [M] int32) -> [M,c] means (voxelize_cuda.cu:44-62)."""
    x, idx, counts = _f32(inputs), _i32(idx), _i32(counts)
    out = torch.empty(counts.shape[0], x.shape[1], dtype=torch.float32, device=x.device)
    _capi.check(_capi.lib().lk_voxelize_fwd(_capi.ptr(x), _capi.ptr(idx), _capi.ptr(counts), x.shape[0], counts.shape[0],
                                            x.shape[1], _capi.ptr(out), _capi.stream()), 'lk_voxelize_fwd')
    return out.to(inputs.dtype)


def voxelize_backward_cuda(top_grad, idx, counts, N):
    """voxelize_backward_cuda(top_grad [M,c], idx [N], counts [M], N) -> [N,c] (voxelize_cuda.cu:64-80)."""
    g, idx, counts = _f32(top_grad), _i32(idx), _i32(counts)
    out = torch.empty(int(N), g.shape[1], dtype=torch.float32, device=g.device)
    _capi.check(_capi.lib().lk_voxelize_bwd(_capi.ptr(g), _capi.ptr(idx), _capi.ptr(counts), int(N), counts.shape[0],
                                            g.shape[1], _capi.ptr(out), _capi.stream()), 'lk_voxelize_bwd')
    return out.to(top_grad.dtype)


def devoxelize_forward_cuda(feat, indices, weight, r):
    """devoxelize_forward_cuda(feat [n,c], indices [N,r^3] int32, weight [N,r^3], r) -> [N,c] (devoxelize_cuda.cu:61-80)."""
    f, ind, w = _f32(feat), _i32(indices), _f32(weight)
    assert ind.shape[1] == int(r) ** 3 and w.shape == ind.shape
    out = torch.empty(ind.shape[0], f.shape[1], dtype=torch.float32, device=f.device)
    _capi.check(_capi.lib().lk_devoxelize_fwd(_capi.ptr(f), _capi.ptr(ind), _capi.ptr(w), ind.shape[0], ind.shape[1],
                                              f.shape[1], _capi.ptr(out), _capi.stream()), 'lk_devoxelize_fwd')
    return out.to(feat.dtype)


def devoxelize_backward_cuda(top_grad, indices, weight, n, r):
    """devoxelize_backward_cuda(top_grad [N,c], indices [N,r^3], weight [N,r^3], n, r) -> [n,c] (devoxelize_cuda.cu:82-101)."""
    g, ind, w = _f32(top_grad), _i32(indices), _f32(weight)
    assert ind.shape[1] == int(r) ** 3
    out = torch.empty(int(n), g.shape[1], dtype=torch.float32, device=g.device)
    _capi.check(_capi.lib().lk_devoxelize_bwd(_capi.ptr(g), _capi.ptr(ind), _capi.ptr(w), ind.shape[0], ind.shape[1],
                                              g.shape[1], int(n), _capi.ptr(out), _capi.stream()), 'lk_devoxelize_bwd')
    return out.to(top_grad.dtype)


def _map_from_pairs(neighbor_map, neighbor_offset, n_rows, row_col, identity_mid):
    """[K, n_rows] output-stationary map from the reference's pair list (lk_kmap_from_pairs)."""
    pairs = _i32(neighbor_map)
    sizes = neighbor_offset.detach().cpu().int().contiguous()        # the reference holds it on the host too
    k = sizes.numel()
    nbr = torch.empty(k, n_rows, dtype=torch.int32, device=pairs.device)
    _capi.check(_capi.lib().lk_kmap_from_pairs(_capi.ptr(pairs) if pairs.numel() else None, sizes.data_ptr(), k, n_rows,
                                               row_col, 1 if identity_mid else 0, _capi.ptr(nbr), _capi.stream()),
                'lk_kmap_from_pairs')
    return nbr


def convolution_forward_cuda(in_feat, out_feat, kernel, neighbor_map, neighbor_offset, transpose):
    """convolution_forward_cuda(in_feat [Nin,Cin], out_feat [Nout,Cout] (zeroed, written in place), kernel
    [K,Cin,Cout], neighbor_map [P,2] int32, neighbor_offset [K] int32 on the host, transpose) -> None
    (convolution_cuda.cu:53-165): out_feat = sum_k gather(in_feat) @ kernel[k] scattered to the output rows."""
    from link_b200.nn.functional.conv import _conv_fwd
    if in_feat.shape[1] != kernel.shape[1]:
        raise ValueError('Input feature size and kernel size mismatch')          # convolution_cuda.cu:57-59
    k = kernel.shape[0]
    n_in, n_out = in_feat.shape[0], out_feat.shape[0]
    mid = (k % 2 == 1) and n_in == n_out                                         # precompute_mid, :74-88
    nbr = _map_from_pairs(neighbor_map, neighbor_offset, n_out, 0 if transpose else 1, mid)
    out = _conv_fwd(_f32(in_feat), _f32(kernel), nbr, n_out)
    out_feat.copy_(out)


def convolution_backward_cuda(in_feat, grad_in_feat, grad_out_feat, kernel, grad_kernel, neighbor_map, neighbor_offset,
                              transpose):
    """convolution_backward_cuda(in_feat, grad_in_feat (out), grad_out_feat, kernel, grad_kernel (out),
    neighbor_map, neighbor_offset (host), transpose) -> None (convolution_cuda.cu:167-278)."""
    from link_b200.nn.functional.conv import _conv_fwd
    k, c_in, c_out = kernel.shape
    n_in, n_out = in_feat.shape[0], grad_out_feat.shape[0]
    # NB: the reference's backward has no precompute_mid shortcut -- every offset goes through its pairs
    to_out = _map_from_pairs(neighbor_map, neighbor_offset, n_out, 0 if transpose else 1, False)   # rows: grad rows
    to_in = _map_from_pairs(neighbor_map, neighbor_offset, n_in, 1 if transpose else 0, False)     # rows: input rows
    x, g, w = _f32(in_feat), _f32(grad_out_feat), _f32(kernel)
    grad_in_feat.copy_(_conv_fwd(g, None, to_in, n_in, weight_t=w))
    gw = torch.empty(k, c_in, c_out, dtype=torch.float32, device=x.device)
    _capi.check(_capi.lib().lk_conv_bwd_weight(_capi.ptr(x), _capi.ptr(g), _capi.ptr(to_out), n_out, k, c_in, c_out,
                                               _capi.ptr(gw), _capi.stream()), 'lk_conv_bwd_weight')
    grad_kernel.copy_(gw)
