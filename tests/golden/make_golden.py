"""Generate tests/golden/*.npz by executing the UNMODIFIED reference.

Runs only in the build container: imports the reference python layer from
/root/reference (torchsparse-u + segmentation/core/models) on top of
oracle/_ref/backend.so (the reference's CPU backend compiled in place by
oracle/build_ref.py).  The fixtures stay inside the envelope in which the
reference CPU path is itself correct (r=2, batch index 0 everywhere for ops that
hash with offsets).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cloud(seed, n_draw, n_keep, extent, batch=1, lo=0):
    g = torch.Generator().manual_seed(seed)
    cs = []
    for b in range(batch):
        c = torch.randint(lo, lo + extent, (n_draw, 3), dtype=torch.int, generator=g)
        c = torch.unique(c, dim=0)
        c = c[torch.randperm(len(c), generator=g)[:n_keep]]
        cs.append(torch.cat([c, torch.full((len(c), 1), b, dtype=torch.int)], 1))
    return torch.cat(cs, 0).contiguous()


def sd_np(module):
    return {k: v.detach().numpy() for k, v in module.state_dict().items()}


def main():
    build_ref.build()
    ts, U, L = ref_import.import_reference()
    import torchsparse.nn.functional as F
    from torchsparse.nn.utils import get_kernel_offsets
    from torchsparse.utils.quantize import sparse_quantize
    torch.set_num_threads(1)  # deterministic accumulation order

    # ---- 1. known-answer vectors for the integer ops -------------------------------
    c4 = torch.tensor([[0, 0, 0, 0], [1, 2, 3, 0], [-1, 0, 5, 1], [1439, 1439, 40, 3]],
                      dtype=torch.int)
    c0 = torch.tensor([[0, 0, 0, 0], [1, 2, 3, 0], [-7, 9, 2, 0]], dtype=torch.int)
    kat = dict(
        coords=c4.numpy(), hash=F.sphash(c4).numpy(),
        coords_b0=c0.numpy(),
        khash2=F.sphash(c0, get_kernel_offsets(2, 1, 1)).numpy(),
        khash3=F.sphash(c0, get_kernel_offsets(3, 1, 1)).numpy(),
        khash3_s2=F.sphash(c0, get_kernel_offsets(3, 2, 1)).numpy(),
        off2=get_kernel_offsets(2, 1, 1).numpy(), off3=get_kernel_offsets(3, 1, 1).numpy(),
        off3_s4=get_kernel_offsets(3, 4, 1).numpy(),
        off311=get_kernel_offsets((3, 1, 1), 1, 1).numpy(),
        query=F.sphashquery(F.sphash(torch.tensor([[1, 2, 3, 0], [9, 9, 9, 0], [0, 0, 0, 0]],
                                                  dtype=torch.int)), F.sphash(c4)).numpy(),
        count=F.spcount(torch.tensor([0, 0, 2, -1, 2, 2], dtype=torch.int), 4).numpy(),
    )
    np.savez_compressed(os.path.join(OUT, 'kat.npz'), **kat)

    # ---- 2. ELKBlock, BASELINE config 1: 8k voxels, C=16, cos_x (2x3)^3 -----------
    def block_case(name, seed, n, C, baseop, groups, s, r, extent=64, tstride=1, lo=0):
        coords = cloud(seed, 2 * n, n, extent, lo=lo)
        if tstride > 1:
            coords[:, :3] *= tstride
            coords = torch.unique(coords, dim=0)[torch.randperm(len(coords),
                                                               generator=torch.Generator().manual_seed(seed))]
        torch.manual_seed(seed)
        feats = torch.randn(len(coords), C)
        blk = L.ELKBlock(C, C, groups=groups, baseop=baseop)
        with torch.no_grad():   # non-trivial affine params so LayerNorm parity is exercised
            for m in (blk.pre_mix[1], blk.norm, blk.norm_local):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.5, 0.5)
            if baseop == 'cos_x':
                blk.alpha.uniform_(0.5, 1.5)
        st = ts.SparseTensor(feats.clone(), coords.clone(), tstride)
        # integer maps exactly as voxel_to_aux / aux_to_voxel compute them
        x_C = torch.cat([torch.div(coords[:, :3], s, rounding_mode='floor').int(), coords[:, 3:]], 1)
        small_C = torch.unique(x_C, dim=0)
        idx_query = F.sphashquery(F.sphash(x_C), F.sphash(small_C))
        counts = F.spcount(idx_query.int(), len(small_C))
        nbr = F.sphashquery(F.sphash(small_C, get_kernel_offsets(r, 1, 1)),
                            F.sphash(small_C)).transpose(0, 1).contiguous()
        with torch.no_grad():
            out = blk(st, s, r)
        kmap = st.kmaps[((tstride,) * 3, (3, 3, 3), (1, 1, 1), (1, 1, 1))]
        # gradient fixture (r=2 CPU devoxelize backward is valid; conv backward on CPU is
        # NotImplemented in the reference, so the local_mix branch is detached)
        np.savez_compressed(
            os.path.join(OUT, name + '.npz'), coords=coords.numpy(), feats=feats.numpy(),
            s=s, r=r, C=C, groups=groups, baseop=baseop, tstride=tstride,
            hash=F.sphash(coords).numpy(), small_C=small_C.numpy(), idx_query=idx_query.numpy(),
            counts=counts.numpy(), nbr_idx=nbr.numpy(), nbmaps=kmap[0].numpy(),
            nbsizes=kmap[1].numpy(), out=out.F.numpy(),
            **{'sd.' + k: v for k, v in sd_np(blk).items()})
        print(name, len(coords), 'M', len(small_C), 'P', len(kmap[0]), 'out', float(out.F.abs().mean()))

    block_case('block_g1_cosx_2x3', 0, 8000, 16, 'cos_x', 1, 3, 2)
    block_case('block_cos_g2_2x3', 1, 3000, 32, 'cos', 2, 3, 2, extent=40, lo=-11)
    block_case('block_sin_g1_2x5', 2, 2000, 8, 'sin', 1, 5, 2, extent=48)
    block_case('block_cosx_stride2', 3, 2500, 16, 'cos_x', 1, 6, 2, extent=40, tstride=2)

    # ---- 3. functional ops on random data ------------------------------------------
    g = torch.Generator().manual_seed(7)
    N, M, c = 5000, 700, 24
    idx = torch.randint(0, M, (N,), generator=g)
    counts = F.spcount(idx.int(), M)
    feats = torch.randn(N, c, generator=g)
    vox = F.spvoxelize(feats, idx, counts)
    nb = torch.randint(-1, M, (900, 8), generator=g).int()
    w = torch.rand(900, 8, generator=g)
    dev = F.spdevoxelize(vox, nb, w, 2)
    np.savez_compressed(os.path.join(OUT, 'ops.npz'), idx=idx.numpy(), counts=counts.numpy(),
                        feats=feats.numpy(), vox=vox.numpy(), nb=nb.numpy(), w=w.numpy(),
                        dev=dev.numpy())

    # ---- 4. conv: SubM k3, strided k2/s2, transposed k2/s2; spdownsample ------------
    coords = cloud(11, 6000, 4000, 36)
    torch.manual_seed(11)
    feats = torch.randn(len(coords), 12)
    st = ts.SparseTensor(feats, coords, 1)
    st.cmaps[st.stride] = st.coords
    import torchsparse.nn as spnn
    c1 = spnn.Conv3d(12, 20, 3)
    c2 = spnn.Conv3d(20, 24, 2, stride=2)
    c3 = spnn.Conv3d(24, 8, 3)
    c4_ = spnn.Conv3d(8, 6, 2, stride=2, transposed=True)
    c5 = spnn.Conv3d(6, 5, 1)
    with torch.no_grad():
        y1 = c1(st); y2 = c2(y1); y3 = c3(y2); y4 = c4_(y3); y5 = c5(y4)
    km = st.kmaps
    save = dict(coords=coords.numpy(), feats=feats.numpy(),
                w1=c1.kernel.detach().numpy(), w2=c2.kernel.detach().numpy(),
                w3=c3.kernel.detach().numpy(), w4=c4_.kernel.detach().numpy(),
                w5=c5.kernel.detach().numpy(),
                y1=y1.F.numpy(), y2=y2.F.numpy(), y2_C=y2.C.numpy(), y3=y3.F.numpy(),
                y4=y4.F.numpy(), y4_C=y4.C.numpy(), y5=y5.F.numpy(),
                ds_k3s2=F.spdownsample(coords, 2, 3, 1).numpy(),
                ds_k2s2_t2=F.spdownsample(y2.C, 2, 2, 2).numpy())
    for key, v in km.items():
        tag = 'kmap_s%d_k%d_st%d' % (key[0][0], key[1][0], key[2][0])
        save[tag + '_nbmaps'] = v[0].numpy()
        save[tag + '_nbsizes'] = v[1].numpy()
    np.savez_compressed(os.path.join(OUT, 'conv.npz'), **save)
    print('conv', {k: tuple(v[2]) for k, v in km.items()})

    # ---- 5. voxel_to_aux with batch 2 (no offset hashing -> valid on CPU) ------------
    coords = cloud(5, 3000, 2000, 30, batch=2, lo=-7)
    torch.manual_seed(5)
    feats = torch.randn(len(coords), 10)
    st = ts.SparseTensor(feats, coords, 1)
    aux, idx, cnt = U.voxel_to_aux(st, 4)
    np.savez_compressed(os.path.join(OUT, 'aux_batch2.npz'), coords=coords.numpy(),
                        feats=feats.numpy(), s=4, aux_F=aux.F.numpy(), aux_C=aux.C.numpy(),
                        idx=idx.numpy(), counts=cnt.numpy())

    # ---- 6. ELKEncoder (cr=0.25 -> 16 channels), cos_x (2x3)^3, eval mode ------------
    coords = cloud(21, 9000, 6000, 48)
    torch.manual_seed(21)
    feats = torch.randn(len(coords), 4)
    enc = L.ELKEncoder(num_classes=19, cr=0.25, baseop='cos_x', r=2, s=3, groups=1).eval()
    with torch.no_grad():
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    st = ts.SparseTensor(feats.clone(), coords.clone(), 1)
    st.cmaps[st.stride] = st.coords
    with torch.no_grad():
        logits = enc(st)
    sd = {k: v for k, v in sd_np(enc).items() if not k.startswith('up')
          and not k.endswith('num_batches_tracked')}
    np.savez_compressed(os.path.join(OUT, 'encoder_cosx_2x3.npz'), coords=coords.numpy(),
                        feats=feats.numpy(), logits=logits.numpy(),
                        level_sizes=np.asarray([v.shape[0] for v in st.cmaps.values()]),
                        **{'sd.' + k: v for k, v in sd.items()})
    print('encoder', logits.shape, float(logits.abs().mean()), [v.shape[0] for v in st.cmaps.values()])

    # ---- 6b. ELKUNet (cr=0.25), cos_x (2x3)^3, eval mode: decoder with transposed convs ------
    import core.models.semantic_kitti.linkunet as linkunet
    coords = cloud(31, 7000, 5000, 44)
    torch.manual_seed(31)
    feats = torch.randn(len(coords), 4)
    unet = linkunet.ELKUNet(num_classes=19, cr=0.25, baseop='cos_x', r=2, s=3, groups=1).eval()
    with torch.no_grad():
        for m in unet.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.2, 0.2)
    st = ts.SparseTensor(feats.clone(), coords.clone(), 1)
    st.cmaps[st.stride] = st.coords
    with torch.no_grad():
        logits = unet(st)
    sd = {k: v for k, v in sd_np(unet).items() if not k.endswith('num_batches_tracked')}
    np.savez_compressed(os.path.join(OUT, 'unet_cosx_2x3.npz'), coords=coords.numpy(),
                        feats=feats.numpy(), logits=logits.numpy(),
                        **{'sd.' + k: v for k, v in sd.items()})
    print('unet', logits.shape, float(logits.abs().mean()))

    # ---- 7. voxelisation front-ends ---------------------------------------------------
    g = torch.Generator().manual_seed(3)
    pts = (torch.rand(4000, 3, generator=g) * 20 - 5).numpy()
    qc, qi, qv = sparse_quantize(pts.copy(), 0.5, return_index=True, return_inverse=True)
    pC = torch.cat([torch.rand(3000, 3, generator=g) * 12, torch.zeros(3000, 1)], 1)
    pF = torch.randn(3000, 4, generator=g)
    pt = ts.PointTensor(pF.clone(), pC.clone())
    vst = U.initial_voxelize(pt, 1.0, 0.5)
    np.savez_compressed(os.path.join(OUT, 'voxelize.npz'), pts=pts, q_coords=qc, q_index=qi,
                        q_inverse=qv, pC=pC.numpy(), pF=pF.numpy(), v_F=vst.F.numpy(),
                        v_C=vst.C.numpy(), v_idx=pt.additional_features['idx_query'][1].numpy(),
                        v_counts=pt.additional_features['counts'][1].numpy())
    sz = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT) if f.endswith('.npz'))
    print('total fixture bytes', sz)


if __name__ == '__main__':
    main()
