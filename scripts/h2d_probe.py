"""Raw pinned host -> device copy bandwidth for the e2e step's buffers (features 119 325 x 64 fp32 and
coordinates 119 325 x 4 int32): the ceiling of the e2e loop, which moves 32.46 MB per scan.  Three kinds
of pinned memory: torch's pin_memory(), cudaHostAlloc through liblinkb200, and write-combined."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from link_b200.tensor import pinned_empty
dev = torch.device('cuda:0')
n = 119_325


def alloc(kind, shape, dt):
    if kind == 'torch':
        return torch.empty(shape, dtype=dt).pin_memory()
    return pinned_empty(shape, dt, write_combined=(kind == 'wc'))


cases = [('feats fp32', (n, 64), torch.float32), ('feats bf16', (n, 64), torch.bfloat16), ('coords', (n, 4), torch.int32),
         ('256 MB', (64 << 20,), torch.float32)]
for rep in range(2):
    for name, shape, dt in cases:
        line = f'{name:12s}'
        for kind in ('torch', 'lk', 'wc'):
            h = [alloc(kind, shape, dt) for _ in range(2)]
            assert h[0].is_pinned()
            for t in h:
                t.view(torch.uint8).fill_(1)          # touch (write) every page
            d = [torch.empty(shape, dtype=dt, device=dev) for _ in range(2)]
            for i in range(4):
                d[i % 2].copy_(h[i % 2], non_blocking=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 40
            e0.record()
            for i in range(reps):
                d[i % 2].copy_(h[i % 2], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nbytes = h[0].numel() * h[0].element_size()
            line += f'   {kind}: {ms * 1e3:7.1f} us {nbytes / ms / 1e6:5.1f} GB/s'
            del h, d
        print(line, flush=True)
