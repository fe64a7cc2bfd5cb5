"""world_size-2 gloo test (CPU) of the N>1 path: frame sharding is a partition, per-rank seeds are
distinct, and the whole-job throughput bookkeeping is max-time / sum-units."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from link_b200.sharding import frame_seed, reduce_throughput, shard_frames


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard_frames(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ms, units = reduce_throughput(10.0 + 5.0 * rank, 1000.0 * (rank + 1))
    seeds = [None] * world
    dist.all_gather_object(seeds, [frame_seed(rank, s) for s in range(3)])
    if rank == 0:
        out.put((gathered, ms, units, seeds))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, ms, units, seeds = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(i for part in gathered for i in part)
    assert flat == list(range(11))                      # partition: disjoint and complete
    assert gathered[0] == [0, 2, 4, 6, 8, 10] and gathered[1] == [1, 3, 5, 7, 9]
    assert ms == 15.0 and units == 3000.0               # max time, summed units
    assert len({s for part in seeds for s in part}) == 6


def test_single_process_is_identity():
    assert reduce_throughput(3.5, 7.0) == (3.5, 7.0)
    assert shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]


def _ddp_worker(rank, world, port, out):
    """Training-side N>1 plumbing: frames shard over ranks, the ONLY collective is DDP's
    gradient all-reduce.  The dense CenterPoint head runs on CPU, so gloo can exercise it here."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    from link_b200.centerpoint import RPN, CenterHead
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)                                     # same initial weights on every rank
    tasks = [dict(num_class=1, class_names=['car'])]
    net = torch.nn.Sequential(
        RPN(layer_nums=[1], ds_layer_strides=[1], ds_num_filters=[8], us_layer_strides=[1], us_num_filters=[8],
            num_input_features=4))
    head = CenterHead(in_channels=8, tasks=tasks, code_weights=[1.0] * 8, common_heads={'reg': (2, 2), 'height': (1, 2),
                      'dim': (3, 2), 'rot': (2, 2)}, share_conv_channel=8)

    class Both(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net, self.head = net, head

        def forward(self, x):
            return self.head(self.net(x))[0]

    def loss_of(model, frame):
        g = torch.Generator().manual_seed(frame_seed(0, frame))
        x = torch.randn(1, 4, 8, 8, generator=g)
        ex = {'hm': [torch.rand(1, 1, 8, 8, generator=g) ** 4], 'ind': [torch.randint(0, 64, (1, 5), generator=g)],
              'mask': [torch.ones(1, 5, dtype=torch.uint8)], 'cat': [torch.zeros(1, 5, dtype=torch.long)],
              'anno_box': [torch.randn(1, 5, 10, generator=g)]}
        return sum(head.loss(ex, model(x), None)['loss'])

    both = Both()
    ddp = DDP(both)
    frames = shard_frames(2, rank, world)                    # one frame per rank
    loss_of(ddp, frames[0]).backward()
    grads = torch.cat([p.grad.flatten() for p in both.parameters()])
    # single-process reference: mean of the two frames' gradients from the same initial weights
    torch.manual_seed(0)
    ref = Both()
    ref.load_state_dict(both.state_dict())
    ref.zero_grad()
    ((loss_of(ref, 0) + loss_of(ref, 1)) / 2).backward()
    want = torch.cat([p.grad.flatten() for p in ref.parameters()])
    if rank == 0:
        out.put((float((grads - want).abs().max()), float(want.abs().max())))
    dist.destroy_process_group()


def test_two_rank_ddp_gradients_equal_mean_over_frames():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scale > 0 and err <= 1e-5 * max(1.0, scale)


def _wrap_worker(rank, world, port, out):
    """wrap_ddp: a model with a never-used branch named like the encoder's decoder (`up1`) trains under
    DDP WITHOUT find_unused_parameters -- the branch is declared ignored -- and the used parameters'
    gradients are the mean over the ranks."""
    from link_b200.sharding import wrap_ddp
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = torch.nn.Linear(4, 8)
            self.bn = torch.nn.BatchNorm1d(8)
            self.head = torch.nn.Linear(8, 3)
            self.up1 = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.BatchNorm1d(8))   # built, never used

        def forward(self, x):
            return self.head(torch.relu(self.bn(self.stem(x))))

    net = Net()
    ddp = wrap_ddp(net)
    xs = [torch.randn(5, 4, generator=torch.Generator().manual_seed(10 + r)) for r in range(world)]
    for _ in range(2):                                      # two steps: a reducer waiting for `up1` would raise on the 2nd
        net.zero_grad(set_to_none=True)
        ddp(xs[rank]).square().mean().backward()
    got = torch.cat([p.grad.flatten() for n, p in net.named_parameters() if not n.startswith('up1')])
    assert all(p.grad is None for n, p in net.named_parameters() if n.startswith('up1'))
    # wrap_ddp does not re-broadcast the buffers at every forward: the ranks' running statistics differ
    # until sync_buffers establishes rank 0's on every rank (what DDP's default does per step)
    from link_b200.sharding import sync_buffers
    mine = net.bn.running_mean.clone()
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert float((both[0] - both[1]).abs().max()) > 1e-4
    sync_buffers(ddp)
    assert torch.equal(net.bn.running_mean, both[0]) and int(net.bn.num_batches_tracked) == 2
    ref = Net()
    ref.load_state_dict(net.state_dict())
    (sum(ref(x).square().mean() for x in xs) / world).backward()
    want = torch.cat([p.grad.flatten() for n, p in ref.named_parameters() if not n.startswith('up1')])
    if rank == 0:
        out.put(float((got - want).abs().max()))
    dist.destroy_process_group()


def test_wrap_ddp_ignores_unused_decoder_branches():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wrap_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err <= 1e-6
