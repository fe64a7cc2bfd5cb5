"""GPU tests of the hand-written backward of the linear-kernel path (csrc/link_fused.cu:
lk_link_bwd_norm / lk_link_bwd_apply behind LinkAggregateFunction) against (a) the oracle's autograd on
the CPU and (b) the composed path (the reference's op sequence on differentiable voxelize / devoxelize
kernels)."""
import numpy as np
import pytest
import torch

from oracle import link_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from link_b200 import _capi
    _capi.lib()
    return torch.device('cuda:0')


def cu(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def _dense_cloud(n, extent, seed, batch=2):
    """Random voxels dense enough that blocks hold many voxels (runs of > 32 and > 64 sorted positions
    per block at s = 7: the permutation-chunk reload path of the fused kernels)."""
    from link_b200.utils.synthetic import random_voxels
    return random_voxels(n, extent, seed=seed, batch=batch)


def test_window_mean_saves_populations(dev):
    """lk_link_window_mean_tot: the saved window populations T[b] equal the sum of the neighbour blocks'
    voxel counts (what the backward pass divides by), and the means equal lk_link_window_mean's."""
    import link_b200.elk as elk
    from link_b200 import SparseTensor
    coords = _dense_cloud(9000, 22, seed=3)
    n, C, s, r = len(coords), 32, 5, 3
    st = SparseTensor(torch.zeros(n, C, device=dev), cu(coords, dev), 1)
    bi = elk.block_index(st, s)
    m = bi.m
    g = torch.Generator().manual_seed(C)
    f = torch.randn(n, C, generator=g).to(dev)
    w = (torch.randn(C, 3, generator=g) * 0.3).to(dev)
    mean = torch.zeros(n, 2 * C, device=dev)
    tot = torch.zeros(n, device=dev)
    got = elk.link_aggregate(f, st.C, bi, r, 'cos', w, save=(mean, tot))
    want = elk.link_aggregate(f, st.C, bi, r, 'cos', w)
    # two runs of the pre-aggregation join per-run partial sums with float atomics in a different order
    torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-6)
    nbr = bi.neighbors(r)[:m].long()
    cnt = bi.counts[:m]
    want_tot = torch.where(nbr >= 0, cnt[nbr.clamp(min=0)], 0).sum(1)
    assert torch.equal(tot[:m].long(), want_tot.long())


@pytest.mark.parametrize('op,C,groups,s,r,n,extent', [('cos', 64, 2, 7, 3, 6000, 40), ('sin', 32, 1, 5, 3, 4000, 30),
                                                      ('cos', 16, 1, 3, 2, 3000, 24), ('cos', 128, 4, 7, 3, 5000, 20),
                                                      ('sin', 64, 2, 3, 2, 4000, 30), ('cos', 32, 2, 2, 2, 2500, 16)])
def test_fused_backward_vs_oracle_autograd_and_composed(dev, op, C, groups, s, r, n, extent, monkeypatch):
    """Training-mode ELKBlock: fused forward + hand-written backward.  Output and every gradient
    (features, pre_mix, pos_weight, local_mix kernel, both LayerNorms) against the oracle's CPU autograd;
    and against the composed GPU path.  r = 2 exercises the transposed neighbour table ({-1,0}^3)."""
    import link_b200.elk as elk
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock
    coords = _dense_cloud(n, extent, seed=C + s + r)
    n = len(coords)
    torch.manual_seed(C + s)
    blk = ELKBlock(C, C, groups=groups, baseop=op)
    with torch.no_grad():
        for mod in (blk.pre_mix[1], blk.norm, blk.norm_local):
            mod.weight.uniform_(0.5, 1.5)
            mod.bias.uniform_(-0.5, 0.5)
    feats = torch.randn(n, C)
    go = torch.randn(n, C, generator=torch.Generator().manual_seed(2))
    # oracle (CPU autograd)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    f_cpu = feats.clone().requires_grad_(True)
    o, parts = O.elk_block_forward(f_cpu, coords, 1, p, s, r, op, groups, return_parts=True)
    # elements whose pre-activation is within float noise of 0 may land on either side of the ReLU in two
    # fp32 implementations (a flipped mask changes the gradient of the whole row and, through the 3^3 conv,
    # of its neighbours): they receive no incoming gradient, so the comparison is mask-independent
    go = go * (parts['pre_act'].detach().abs() > 1e-3)
    o.backward(go)
    blk = blk.to(dev).train()

    def run(fused):
        monkeypatch.setattr(elk, 'FUSED_BACKWARD', fused)
        blk.zero_grad(set_to_none=True)
        f = feats.to(dev).requires_grad_(True)
        out = blk(SparseTensor(f, cu(coords, dev), 1), s, r).F
        out.backward(go.to(dev))
        return out.detach().cpu().numpy(), f.grad.cpu().numpy(), {k: v.grad.cpu().numpy() for k, v in blk.named_parameters()}

    out_f, df_f, gp_f = run(True)
    out_c, df_c, gp_c = run(False)
    np.testing.assert_allclose(out_f, o.detach().numpy(), rtol=1e-4, atol=4e-5)
    np.testing.assert_allclose(out_f, out_c, rtol=1e-4, atol=4e-5)
    for name, ref_df, ref_gp in (('oracle', f_cpu.grad.numpy(), {k: v.grad.numpy() for k, v in p.items() if v.grad is not None}),
                                 ('composed', df_c, gp_c)):
        np.testing.assert_allclose(df_f, ref_df, rtol=2e-3, atol=1e-4, err_msg=f'd feats vs {name}')
        for k in ['pre_mix.0.weight', 'pre_mix.1.weight', 'pre_mix.1.bias', 'pos_weight.0.weight', 'local_mix.0.kernel',
                  'norm.weight', 'norm.bias', 'norm_local.weight', 'norm_local.bias']:
            ref = ref_gp[k]
            np.testing.assert_allclose(gp_f[k], ref, rtol=5e-3, atol=5e-5 * max(1.0, float(np.abs(ref).max())),
                                       err_msg=f'{k} vs {name}')


def test_fused_backward_is_deterministic(dev):
    """One warp owns a block: the backward block sums use no atomics, so dF / dlocal are bit-identical
    run to run (the parameter gradients join per-CTA partial sums with atomics and may differ in the
    last bits)."""
    from link_b200 import SparseTensor
    from link_b200.elk import ELKBlock
    coords = _dense_cloud(8000, 30, seed=11)
    torch.manual_seed(0)
    blk = ELKBlock(64, 64, groups=2, baseop='cos').to(dev).train()
    feats = torch.randn(len(coords), 64, device=dev)
    go = torch.randn(len(coords), 64, device=dev)
    grads = []
    for _ in range(2):
        f = feats.clone().requires_grad_(True)
        out = blk(SparseTensor(f, cu(coords, dev), 1), 7, 3).F
        out.backward(go)
        grads.append(f.grad.clone())
    # dF passes through pre_mix / the conv backward (atomics there): compare at round-off level
    np.testing.assert_allclose(grads[0].cpu().numpy(), grads[1].cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_tselk_block_fused_backward_vs_composed(dev, monkeypatch):
    """Detection block in training mode: fused backward == composed path (pos_weight gradient flows into
    the first inc/2 rows of the Linear(3, inc) only, ts_elk.py:168)."""
    import link_b200.elk as elk
    from link_b200 import SparseTensor
    from link_b200.ts_elk import TSELKBlock
    coords = _dense_cloud(4000, 30, seed=5)
    torch.manual_seed(1)
    blk = TSELKBlock(32, 32, baseop='cos').to(dev).train()
    feats = torch.randn(len(coords), 32)
    go = torch.randn(len(coords), 32, generator=torch.Generator().manual_seed(3)).to(dev)
    res = []
    for fused in (True, False):
        monkeypatch.setattr(elk, 'FUSED_BACKWARD', fused)
        blk.zero_grad(set_to_none=True)
        f = feats.to(dev).requires_grad_(True)
        out = blk.forward_(SparseTensor(f, cu(coords, dev), 1), 7).F
        out.backward(go)
        res.append((out.detach().cpu().numpy(), f.grad.cpu().numpy(), {k: v.grad.cpu().numpy() for k, v in blk.named_parameters()}))
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-4, atol=4e-5)
    np.testing.assert_allclose(res[0][1], res[1][1], rtol=2e-3, atol=1e-4)
    for k, ref in res[1][2].items():
        np.testing.assert_allclose(res[0][2][k], ref, rtol=5e-3, atol=5e-5 * max(1.0, float(np.abs(ref).max())), err_msg=k)
    assert float(np.abs(res[0][2]['pos_weight.0.weight'][16:]).max()) == 0.0
