"""Reference GPU arm: the reference's OWN CUDA kernels (oracle/_ref/backend_cuda.so, built by
oracle/build_ref.py::build_cuda from the sources under /root/reference) driven by a restatement of
the reference's Python glue for one ELKBlock forward.

TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).  /root/reference is not
available on the GPU box, so the thin Python layer between the model and `torchsparse.backend` is
restated here op for op -- same backend calls, same torch ops in between (cat, unique, nonzero,
.cpu() syncs, fresh allocations), each function citing the lines it follows -- and the kernels that
run are the reference's own.  Used by bench.py (`reference_gpu` entry) and by a GPU test that
checks this arm against the CPU oracle, so the number it produces belongs to a correct forward.
"""
import numpy as np
import torch
import torch.nn.functional as TF

from . import build_ref

_backend = None


def backend():
    global _backend
    if _backend is None:
        _backend = build_ref.load_cuda()
    return _backend


def available() -> bool:
    import os
    return os.path.exists(build_ref.OUT_CUDA) and torch.cuda.is_available()


# -- torchsparse/nn/functional/hash.py:10-37
def sphash(coords, offsets=None):
    coords = coords.contiguous()
    if offsets is None:
        return backend().hash_cuda(coords)
    return backend().kernel_hash_cuda(coords, offsets.contiguous())


# -- torchsparse/nn/functional/query.py:8-33
def sphashquery(queries, references):
    queries, references = queries.contiguous(), references.contiguous()
    sizes = queries.size()
    queries = queries.view(-1)
    indices = torch.arange(len(references), device=queries.device, dtype=torch.long)
    output = backend().hash_query_cuda(queries, references, indices)
    return (output - 1).view(*sizes)


# -- torchsparse/nn/functional/count.py:8-16
def spcount(coords, num):
    return backend().count_cuda(coords.contiguous(), num)


# -- torchsparse/nn/functional/voxelize.py:13-31 (forward)
def spvoxelize(feats, coords, counts):
    return backend().voxelize_forward_cuda(feats.contiguous(), coords.contiguous().int(), counts)


# -- torchsparse/nn/functional/devoxelize.py:54-73 (forward)
def spdevoxelize(feats, coords, weights, r):
    return backend().devoxelize_forward_cuda(feats.contiguous(), coords.contiguous().int(),
                                             weights.contiguous(), r)


# -- torchsparse/nn/utils/kernel.py:11-32
def get_kernel_offsets(size, stride=1, dilation=1, device='cuda'):
    size = (size,) * 3 if isinstance(size, int) else tuple(size)
    stride = (stride,) * 3 if isinstance(stride, int) else tuple(stride)
    dilation = (dilation,) * 3 if isinstance(dilation, int) else tuple(dilation)
    offsets = [(np.arange(-size[k] // 2 + 1, size[k] // 2 + 1) * stride[k] * dilation[k]) for k in range(3)]
    if np.prod(size) % 2 == 1:
        offsets = [[x, y, z] for z in offsets[2] for y in offsets[1] for x in offsets[0]]
    else:
        offsets = [[x, y, z] for x in offsets[0] for y in offsets[1] for z in offsets[2]]
    return torch.tensor(np.array(offsets), dtype=torch.int, device=device)


# -- torchsparse/nn/functional/conv.py:83-147 (stride-1 branch) + ConvolutionFunction.forward 16-62
def conv3d_subm(feats, coords, weight, tensor_stride=1, kmaps=None):
    key = ('subm3', tensor_stride)
    kmap = kmaps.get(key) if kmaps is not None else None
    if kmap is None:
        offsets = get_kernel_offsets(3, stride=tensor_stride, device=feats.device)
        references = sphash(coords)
        queries = sphash(coords, offsets)
        results = sphashquery(queries, references)
        nbsizes = torch.sum(results != -1, dim=1)
        nbmaps = torch.nonzero(results != -1)
        nbmaps[:, 0] = results.view(-1)[nbmaps[:, 0] * results.size(1) + nbmaps[:, 1]]
        kmap = [nbmaps, nbsizes, (feats.shape[0], coords.shape[0])]
        if kmaps is not None:
            kmaps[key] = kmap
    nbmaps, nbsizes, sizes = kmap
    output = torch.zeros(sizes[1], weight.size(-1), dtype=feats.dtype, device=feats.device)
    backend().convolution_forward_cuda(feats.contiguous(), output, weight.contiguous(),
                                       nbmaps.int().contiguous(), nbsizes.int().contiguous().cpu(), False)
    return output


# -- segmentation/core/models/utils.py:44-58
def voxel_to_aux(F_large, C_large, s):
    x_C = torch.cat([torch.div(C_large[:, :3], s, rounding_mode='floor').int(), C_large[:, 3:]], dim=1)
    large_x_hash = sphash(x_C)
    small_x_C = torch.unique(x_C, dim=0)
    small_x_hash = sphash(small_x_C)
    idx_query = sphashquery(large_x_hash, small_x_hash)
    counts = spcount(idx_query.int(), len(small_x_hash))
    inserted_feat = spvoxelize(F_large, idx_query, counts)
    return inserted_feat, small_x_C, idx_query, counts


# -- segmentation/core/models/utils.py:61-84
def aux_to_voxel(F_small, C_small, idx, counts, r=2):
    offsets = get_kernel_offsets(r, 1, 1, device=F_small.device)
    neighbor_hash = sphash(C_small, offsets)
    small_hash = sphash(C_small)
    idx_query = sphashquery(neighbor_hash, small_hash)
    idx_query = idx_query.transpose(0, 1).contiguous()
    f = torch.cat([F_small, torch.ones_like(F_small[:, :1])], dim=1)
    f = f * counts.unsqueeze(dim=-1)
    weights = torch.ones(F_small.shape[0], r ** 3, device=F_small.device).float()
    weights[idx_query == -1] = 0
    new_feat = spdevoxelize(f, idx_query, weights, r)
    new_feat = new_feat[:, :-1] / new_feat[:, -1:]
    return new_feat[idx]


# -- segmentation/core/models/semantic_kitti/linkencoder.py:124-185 ('cos' / 'sin' / 'cos_x')
def elk_block_forward(feats, coords, tensor_stride, p, s, r, baseop='cos', groups=1):
    C = feats.shape[1]
    F_input = TF.layer_norm(TF.linear(feats, p['pre_mix.0.weight']), (C,),
                            p['pre_mix.1.weight'], p['pre_mix.1.bias'], 1e-6)
    local = conv3d_subm(feats, coords, p['local_mix.0.kernel'], tensor_stride)
    xyz = coords[:, :3].float()
    if baseop == 'cos_x':
        pos = TF.linear(xyz / tensor_stride, p['pos_weight.0.weight']) * p['alpha']
    else:
        pos = TF.linear(xyz, p['pos_weight.0.weight']).repeat([1, groups])
    sin, cos = torch.sin(pos), torch.cos(pos)
    if baseop == 'sin':
        planes = torch.cat([F_input * sin, F_input * cos], dim=1)
    elif baseop == 'cos':
        planes = torch.cat([F_input * cos, F_input * sin], dim=1)
    else:
        lin = F_input * pos
        planes = torch.cat([F_input * cos, F_input * sin, lin], dim=1)
    aux_F, small_C, idx, counts = voxel_to_aux(planes, coords, s)
    vf = aux_to_voxel(aux_F, small_C, idx, counts, r)
    if baseop == 'sin':
        new = vf[:, :C] * cos - vf[:, C:] * sin
    elif baseop == 'cos':
        new = vf[:, :C] * cos + vf[:, C:] * sin
    else:
        new = vf[:, :C] * cos + vf[:, C:2 * C] * sin + (vf[:, 2 * C:] - lin)
    new = TF.layer_norm(new, (C,), p['norm.weight'], p['norm.bias'], 1e-6)
    loc = TF.layer_norm(local, (C,), p['norm_local.weight'], p['norm_local.bias'], 1e-6)
    return torch.relu(new + loc)
