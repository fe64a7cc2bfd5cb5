"""Multi-GPU plumbing of the hot path: frames (scans) are independent units -- the batch index is
the 4th hashed coordinate, so blocks, kernel maps and windows never cross frames (reference:
hash_cuda.cu:14-19, utils.py:45) -- hence the forward path shards over ranks with NO data-path
collective.  The only communication is the throughput bookkeeping below (and, for training,
PyTorch DDP's gradient all-reduce).  Works on any torch.distributed backend (nccl on GPUs, gloo in
the CPU tests)."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ['shard_frames', 'frame_seed', 'reduce_throughput', 'wrap_ddp', 'sync_buffers', 'bind_host_to_gpu', 'parse_cpulist']


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


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (the sysfs cpulist format) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus: List[int] = []
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int, sysfs: str = '/sys') -> dict:
    """One process per GPU: pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE
    the pinned staging buffers are allocated -- `cudaHostAlloc` places pages by the calling thread's
    policy (local node), so the per-step host->device copy (32 MB per scan, the e2e bound) then
    leaves memory that is one PCIe root away instead of crossing the socket interconnect, and eight
    ranks no longer draw from one node's memory controllers.  The device's sysfs node
    (`/sys/bus/pci/devices/<id>/local_cpulist`) names the CPUs; they are intersected with the
    affinity the process already has (container cpusets).  Returns what was done; never raises for a
    missing sysfs entry (single-node hosts report numa_node = -1: nothing to do)."""
    import os
    info = {'device': device_index, 'bound': False}
    try:
        import ctypes as C
        from link_b200 import _capi
        buf = C.create_string_buffer(32)
        _capi.check(_capi.lib().lk_device_pci_bus_id(device_index, buf, 32), 'lk_device_pci_bus_id')
        bus = buf.value.decode().lower()
        info['pci'] = bus
        base = os.path.join(sysfs, 'bus', 'pci', 'devices', bus)
        with open(os.path.join(base, 'numa_node')) as f:
            info['numa_node'] = int(f.read().strip())
        with open(os.path.join(base, 'local_cpulist')) as f:
            local = set(parse_cpulist(f.read()))
        have = os.sched_getaffinity(0)
        want = sorted(local & have)
        info['local_cpus'], info['allowed_cpus'] = len(local), len(have)
        if info['numa_node'] >= 0 and want and len(want) < len(have):
            os.sched_setaffinity(0, want)
            info['bound'] = True
            info['cpus'] = len(want)
    except (OSError, ValueError, RuntimeError) as e:
        info['skipped'] = f'{type(e).__name__}: {e}'
    return info


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
    # broadcast_buffers: DDP's default re-broadcasts every buffer (here: the running statistics of ~30
    # BatchNorms) from rank 0 at the start of EVERY forward -- a coalesce / broadcast / scatter sequence on
    # the critical path of a step whose only effect is that all ranks carry rank 0's running statistics.
    # The same end state is reached by broadcasting them once, before evaluation or a checkpoint
    # (`sync_buffers`), so it is off here.
    kw = dict(find_unused_parameters=False, gradient_as_bucket_view=True, broadcast_buffers=False)
    kw.update(kwargs)
    if device_index is not None:
        kw['device_ids'] = [device_index]
    return DDP(model, **kw)


def sync_buffers(model: torch.nn.Module, src: int = 0) -> None:
    """Broadcast every buffer (BatchNorm running statistics, counters) from rank `src`: the state DDP's
    default `broadcast_buffers=True` maintains at every step, established once -- call before evaluation
    or saving a checkpoint when the model was wrapped with `wrap_ddp` (which turns the per-step broadcast
    off).  Single process: no-op."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    m = model.module if hasattr(model, 'module') else model
    with torch.no_grad():
        for b in m.buffers():
            dist.broadcast(b, src)
