"""Multi-GPU plumbing of the hot path: frames (scans) are independent units -- the batch index is
the 4th hashed coordinate, so blocks, kernel maps and windows never cross frames (reference:
hash_cuda.cu:14-19, utils.py:45) -- hence the forward path shards over ranks with NO data-path
collective.  The only communication is the throughput bookkeeping below (and, for training,
PyTorch DDP's gradient all-reduce).  Works on any torch.distributed backend (nccl on GPUs, gloo in
the CPU tests)."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ['shard_frames', 'frame_seed', 'reduce_throughput', 'wrap_ddp']


def shard_frames(num_frames: int, rank: int, world_size: int) -> List[int]:
    """Frame indices owned by `rank`: round-robin, so consecutive (similarly sized) scans spread
    over the ranks.  Every frame is owned by exactly one rank."""
    assert 0 <= rank < world_size
    return list(range(rank, num_frames, world_size))


def frame_seed(rank: int, step: int) -> int:
    """Seed of the synthetic scan a rank processes at a step (rank*1000 + step)."""
    return rank * 1000 + step


def reduce_throughput(local_ms: float, local_units: float, device=None) -> Tuple[float, float]:
    """(max over ranks of the device time, sum over ranks of the processed units).  Whole-job
    throughput = units / max-time.  Single process: identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_ms), float(local_units)
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def wrap_ddp(model: torch.nn.Module, device_index=None, unused_prefixes: Sequence[str] = ('up1', 'up2', 'up3', 'up4'),
             **kwargs):
    """DistributedDataParallel around a LinK model (reference: segmentation/train.py:99-100 wraps with
    `find_unused_parameters=True` because ELKEncoder.forward never touches the decoder branches
    up1..up4 it constructs, linkencoder.py:289-320 vs 339-381).  Here those parameters are declared
    to DDP as ignored instead: the reducer neither waits for their gradients nor walks the autograd
    graph after every forward to find them (host time on a host-bound step), and the gradient
    buckets alias the .grad tensors."""
    from torch.nn.parallel import DistributedDataParallel as DDP
    ignored = [n for n, _ in model.named_parameters() if n.split('.')[0] in unused_prefixes]
    ignored += [n for n, _ in model.named_buffers() if n.split('.')[0] in unused_prefixes]
    if ignored:
        DDP._set_params_and_buffers_to_ignore_for_model(model, ignored)
    kw = dict(find_unused_parameters=False, gradient_as_bucket_view=True)
    kw.update(kwargs)
    if device_index is not None:
        kw['device_ids'] = [device_index]
    return DDP(model, **kw)
