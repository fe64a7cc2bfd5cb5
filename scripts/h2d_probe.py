"""Raw pinned host -> device copy bandwidth for the e2e step's buffers (features 119 325 x 64 fp32 and
coordinates 119 325 x 4 int32): the ceiling of the e2e loop, which moves 32.46 MB per scan."""
import torch, time
dev = torch.device('cuda:0')
n = 119_325
cases = [('feats fp32', (n, 64), torch.float32), ('feats bf16', (n, 64), torch.bfloat16), ('feats fp16', (n, 64), torch.float16),
         ('(n,32) fp32', (n, 32), torch.float32), ('coords', (n, 4), torch.int32), ('256 MB', (64 << 20,), torch.float32)]
cases += [(f'{mb} MB u8', (mb << 20,), torch.uint8) for mb in (4, 8, 12, 16, 20, 24, 32)]
for name, shape, dt in cases:
    h = [torch.empty(shape, dtype=dt).pin_memory() for _ in range(2)]
    d = [torch.empty(shape, dtype=dt, device=dev) for _ in range(2)]
    for i in range(4):
        d[i % 2].copy_(h[i % 2], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    t0 = time.perf_counter()
    e0.record()
    for i in range(reps):
        d[i % 2].copy_(h[i % 2], non_blocking=True)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = h[0].numel() * h[0].element_size()
    print(f'{name:12s} {nbytes / 1e6:8.2f} MB  {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:6.1f} GB/s   host enqueue {(t1 - t0) / reps * 1e6:.1f} us/copy')
