#!/usr/bin/env python
"""bench.py -- LinK hot-path throughput on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload block|encoder] [--voxels 120000]

A "step" is one pass of the hot path over one fresh synthetic SemanticKITTI-shaped scan:
  block   : ELKBlock(64,64, groups=2, 'cos'), (3x7)^3  (pre_mix, local_mix 3^3 sparse conv, block
            index build, pre-aggregation, outer-block reuse, LayerNorms) on ~120k active voxels
  encoder : full ELKEncoder cos:(3x7)^3 forward on the same scan (BASELINE config 2)
Every step starts from raw (coords, feats): kernel maps and block index maps are REBUILT each
step (nothing cached across steps).  Multi-GPU: one process per GPU, each rank runs its own scans
(frames are independent: no data-path collective, weak scaling); value = voxels of all ranks /
max-over-ranks time.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'LinK-block voxels/sec @120k active voxels'
UNIT = 'voxels/s'
C_BLOCK, GROUPS, BASEOP, S_BLK, R_BLK = 64, 2, 'cos', 7, 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='block', choices=['block', 'encoder', 'train_encoder'])
    ap.add_argument('--no-train', action='store_true', help='skip the extra DDP training measurement')
    ap.add_argument('--amp', default='none', choices=['none', 'bf16'], help='train_encoder: torch.autocast(bfloat16) + single-pass TF32 sparse convs')
    ap.add_argument('--voxels', type=int, default=120_000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-encoder', action='store_true', help='skip the extra encoder measurement')
    ap.add_argument('--e2e-ring', action='store_true', help='(kept for compatibility: the pipelined e2e loop always runs)')
    return ap.parse_args()


def make_scan(n_voxels, seed):
    from link_b200.utils.synthetic import kitti_like_voxels
    c3, f4 = kitti_like_voxels(n_voxels, seed=seed)
    coords = np.concatenate([c3, np.zeros((len(c3), 1), np.int32)], 1).astype(np.int32)
    return coords, f4.astype(np.float32)


def block_params(seed=0):
    """Random-init ELKBlock weights as a plain state dict (same init as the module)."""
    from link_b200.elk import ELKBlock
    torch.manual_seed(seed)
    blk = ELKBlock(C_BLOCK, C_BLOCK, groups=GROUPS, baseop=BASEOP)
    return blk


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the committed
# `ncu --set full` capture of THIS round (profiles/r02_ncu_full.md: 39.7 MB read, < 0.1 MB written, two
# launches); ncu cannot run inside the bench, so the line names its source next to the number.
NCU_TRAFFIC = {'link_preagg_ring_kernel': 39.7e6}
NCU_TRAFFIC_SOURCE = 'profiles/r02_ncu_full.md (ncu --set full --clock-control none, scripts/profile_step.py, same scan and kernel)'


def workload_name(args, n):
    return workload_name_for(args.workload, n)


def config_for(workload, n, world):
    """`config` of the JSON line -- the SAME keys and strings from both arms (--impl ours / reference):
    it names the workload; how each arm ran it is in the arm's own keys."""
    return {'workload': workload_name_for(workload, n),
            'l2': 'flushed between timed iterations (256 MiB memsets, outside the events), repeated until they '
                  'outlast the host enqueue of one step so that the timed region is device time, not launch waiting',
            'bounds': 'device-resident loop: the scan\'s coordinate bounds (8 ints, a loader-side property) are '
                      'pre-seeded; e2e loop: computed on the host from the pinned coordinate buffer inside '
                      'the timed region (SparseTensor.from_host)',
            'parallelism': f'{world} independent frame streams (no data-path collective)'}


def workload_name_for(workload, n):
    if workload == 'block':
        return (f'ELKBlock cos:(3x7)^3 C={C_BLOCK} groups={GROUPS} fwd, synthetic '
                f'SemanticKITTI-shaped scan, N={n} active voxels, index+kernel maps rebuilt per step')
    return (f'ELKEncoder cos:(3x7)^3 cr=1.0 fwd, synthetic SemanticKITTI-shaped scan, N={n} active '
            'voxels, all maps rebuilt per step')


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    def __init__(self, index=0):
        self.rows, self.index, self.stop = [], index, threading.Event()
        self.proc = None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [p.strip() for p in line.split(',')]))

    def finish(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_block_step(O, blk_sd, feats_t, coords):
    return O.elk_block_forward(feats_t, coords, 1, blk_sd, S_BLK, R_BLK, BASEOP, GROUPS)


def run_cpu(args, n_sample, steps, warmup):
    """Reference arm / cpu_baseline: the CPU restatement of the same path (oracle port; the
    reference's own CPU devoxelize is wrong for r=3 -- devoxelize_cpu.cpp:19-24 -- so
    oracle/_ref cannot run this configuration), all host threads torch gives us."""
    from oracle import link_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    coords, f4 = make_scan(n_sample, seed=0)
    n = len(coords)
    if args.workload == 'block':
        blk = block_params()
        sd = {k: v.detach() for k, v in blk.state_dict().items()}
        feats = torch.randn(n, C_BLOCK, generator=torch.Generator().manual_seed(1))
        fn = lambda: cpu_block_step(O, sd, feats, coords)
    else:
        from link_b200.linkencoder import ELKEncoder
        torch.manual_seed(0)
        enc = ELKEncoder(num_classes=19, cr=1.0, baseop=BASEOP, r=R_BLK, s=S_BLK, groups=GROUPS).eval()
        sd = {k: v.detach() for k, v in enc.state_dict().items()}
        feats = torch.from_numpy(f4)
        fn = lambda: O.elk_encoder_forward(sd, feats, coords, s=S_BLK, r=R_BLK, baseop=BASEOP,
                                           groups=GROUPS)
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / steps
    return n / dt, dt, n, cores


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # every step = the whole scan (0.7 s of CPU work at 120k voxels on 16 cores), so K steps + W warm-up
    # steps as asked stay within minutes for the driver's K, W; the encoder workload (minutes per scan
    # on the CPU) is sampled at 30k voxels
    n_sample = args.voxels if args.workload == 'block' else min(args.voxels, 30_000)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    v, dt, n, cores = run_cpu(args, n_sample, steps, warmup)
    sample = f'{steps} steps after {warmup} warm-up of the same workload at N={n} voxels'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_for(args.workload, n, args.gpus),
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


# ----------------------------------------------------------------------------- GPU arm
def run_workload(args, workload, steps, warmup, dev, rank, world, local, with_clocks):
    """Times `steps` steps of one workload on this rank.  Returns a dict of local measurements
    (device ms, e2e ms, voxels, launches, per-kernel timers, clocks)."""
    import torch.distributed as dist
    from link_b200 import SparseTensor, _capi
    from link_b200.nn.functional import _index
    from link_b200.sharding import frame_seed

    # per-rank scans (weak scaling): 2 distinct scans alternate so that no step can reuse state
    scans = [make_scan(args.voxels, seed=frame_seed(rank, i)) for i in range(2)]
    if workload == 'block':
        model = block_params().to(dev).eval()
        feats_host = [torch.randn(len(c), C_BLOCK, generator=torch.Generator().manual_seed(i)).pin_memory()
                      for i, (c, _) in enumerate(scans)]
    else:
        from link_b200.linkencoder import ELKEncoder
        torch.manual_seed(0)
        model = ELKEncoder(num_classes=19, cr=1.0, baseop=BASEOP, r=R_BLK, s=S_BLK, groups=GROUPS)
        model = model.to(dev).eval()
        feats_host = [torch.from_numpy(f).pin_memory() for _, f in scans]
    coords_host = [torch.from_numpy(c).pin_memory() for c, _ in scans]
    bounds = [(c.min(0), c.max(0)) for c, _ in scans]     # dataset property, known on the host
    coords_dev = [c.to(dev) for c in coords_host]
    feats_dev = [f.to(dev) for f in feats_host]
    n_vox = [len(c) for c, _ in scans]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(i, coords, feats):
        return step_st(i, SparseTensor(feats, coords, 1))

    def step_st(i, st, seed_bounds=True):
        if seed_bounds:
            _index.set_coord_bounds(st.kmaps, bounds[i][0], bounds[i][1])
        with torch.no_grad():
            if workload == 'block':
                return model(st, S_BLK, R_BLK).F
            return model(st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(torch.zeros(1))      # CPU tensor -> gloo: a barrier across ranks
        torch.cuda.synchronize()

    # ---- device-resident loop (value) ------------------------------------------------------
    # The step is enqueued from python (one FFI call per block, ~185 calls per encoder) and the host needs
    # longer per step than the device: if the first event is recorded while the stream is empty, the
    # event window contains the device WAITING for launches (measured: 0.25 ms per block step against
    # 0.197 ms when the launches are already queued, and 0.199 ms for a CUDA-graph replay of the same
    # step; scripts/step_graph_probe.py).  So the L2 flush in front of every timed step is repeated
    # until it outlasts the host's enqueue time of one step (measured during warm-up): the device
    # reaches the first event with the whole step queued behind it and `value` is device time.  The host
    # side is reported separately (`host_enqueue_ms_per_step`) and is inside `e2e`.
    host_us = []
    for w in range(warmup):
        f = feats_dev[w % 2].clone()
        t_h = time.perf_counter()
        step(w % 2, coords_dev[w % 2], f)
        host_us.append((time.perf_counter() - t_h) * 1e6)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        flush.zero_()
    e1.record()
    torch.cuda.synchronize()
    flush_us = 1e3 * e0.elapsed_time(e1) / 4
    n_flush = int(min(48, max(2, np.ceil(3.0 * float(np.median(host_us[-3:])) / flush_us) + 2)))

    def flush_l2():
        for _ in range(n_flush):
            flush.zero_()
    sampler = ClockSampler(local)
    if with_clocks:
        sampler.start()
    launches0 = _capi.launch_count()
    evs = []
    t_wall0 = time.time()
    barrier()
    prof = None
    if os.environ.get('LINKB200_PROFILE_HOST') == '1' and rank == 0:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t_cpu0 = time.perf_counter()
    for k in range(steps):
        i = k % 2
        f = feats_dev[i].clone()          # the block overwrites st.F; clone outside the timed region
        flush_l2()                        # L2 flush between timed iterations (outside the events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(i, coords_dev[i], f)
        e1.record()
        evs.append((e0, e1))
    host_enqueue_ms = (time.perf_counter() - t_cpu0) * 1e3 / steps   # python + launch cost
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats('cumulative').print_stats(45)
    barrier()
    t_wall1 = time.time()
    launches = _capi.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- instrumented pass: same steps with one python-level call per kernel, CUDA events around
    #      each library call on the launching stream (per-kernel times for `kernels`/`roofline`) ----
    _capi.TIMERS = {}
    for k in range(steps):
        i = k % 2
        f = feats_dev[i].clone()
        flush_l2()
        step(i, coords_dev[i], f)
    barrier()
    timers, _capi.TIMERS = _capi.TIMERS, None
    kern = {}
    for name, lst in timers.items():
        ms = [a.elapsed_time(b) for a, b, _ in lst]
        kern[name] = {'launches_per_step': len(ms) / steps, 'avg_us': 1e3 * float(np.mean(ms)),
                      'bytes': float(np.mean([x[2] for x in lst]))}
    # ---- roofline pass for the HBM-bound kernel BASELINE.json names (pre-aggregation): the kernel
    #      alone, back to back on its launching stream, over ROTATING feature buffers whose total
    #      size exceeds L2 (6 x 30 MB > 126 MB), so every launch streams its rows from HBM while
    #      launch latency is amortised; CUDA events around the batch ----
    roof = None
    if workload == 'block':
        roof = preagg_roofline(dev, coords_dev[0], bounds[0], model)
        roof['path'] = path_roofline(dev, coords_dev[0], bounds[0], model)
    # ---- end-to-end loop: pinned host buffers -> H2D -> step -> D2H of a per-channel checksum --
    h2d = d2h = 0
    e2e_evs = []
    out_host = torch.empty(C_BLOCK if workload == 'block' else 19, dtype=torch.float32).pin_memory()
    w_e2e = min(3, warmup)
    for k in range(w_e2e + steps):
        i = k % 2
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        # public API: SparseTensor.from_host uploads the coordinates on the current stream and the
        # features on a copy stream; the block's index build overlaps the feature upload
        st = SparseTensor.from_host(feats_host[i], coords_host[i], 1, device=dev)   # seeds the bounds itself
        out = step_st(i, st, seed_bounds=False)
        out_host.copy_(out.sum(dim=0), non_blocking=True)
        e1.record()
        if k >= w_e2e:
            e2e_evs.append((e0, e1))
            h2d = coords_host[i].numel() * 4 + feats_host[i].numel() * 4
            d2h = out_host.numel() * 4
    barrier()
    e2e_ms = sum(a.elapsed_time(b) for a, b in e2e_evs)
    # ---- the same end-to-end loop with the features crossing PCIe as bf16 (from_host(dtype=float32)
    #      widens them on the device): an opt-in of the upload API, reported next to the fp32 line ----
    e2e_bf16 = None
    if workload == 'block':
        feats_bf16 = [f.to(torch.bfloat16).pin_memory() for f in feats_host]
        evs16 = []
        for k in range(w_e2e + steps):
            i = k % 2
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st = SparseTensor.from_host(feats_bf16[i], coords_host[i], 1, device=dev, dtype=torch.float32)
            out = step_st(i, st, seed_bounds=False)
            out_host.copy_(out.sum(dim=0), non_blocking=True)
            e1.record()
            if k >= w_e2e:
                evs16.append((e0, e1))
        barrier()
        ms16 = sum(a.elapsed_time(b) for a, b in evs16)
        e2e_bf16 = {'ms_per_step': ms16 / steps, 'value': float(sum(n_vox[k % 2] for k in range(steps))) / (ms16 * 1e-3),
                    'unit': UNIT, 'h2d_bytes_per_step': coords_host[0].numel() * 4 + feats_bf16[0].numel() * 2,
                    'd2h_bytes_per_step': out_host.numel() * 4,
                    'note': 'host features stored as bf16 (inputs rounded to bf16), widened to fp32 on the device; '
                            'arithmetic unchanged (fp32); this rank only'}
    e2e_ring = None
    if workload in ('block', 'encoder'):
        # same calls through a caller-owned staging ring, issued back to back, ONE event pair around
        # the loop: the upload of scan i+1 overlaps the processing of scan i
        from link_b200.tensor import UploadRing
        ring = UploadRing(max(n_vox), feats_host[0].shape[1], device=dev, depth=int(os.environ.get('LINKB200_RING_DEPTH', '2')))
        for k in range(w_e2e):
            out_host.copy_(step_st(k % 2, ring.upload(feats_host[k % 2], coords_host[k % 2], 1), seed_bounds=False).sum(dim=0), non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_h = time.perf_counter()
        for k in range(steps):
            out_host.copy_(step_st(k % 2, ring.upload(feats_host[k % 2], coords_host[k % 2], 1), seed_bounds=False).sum(dim=0), non_blocking=True)
        host_ms = (time.perf_counter() - t_h) * 1e3 / steps
        e1.record()
        barrier()
        e2e_ring = {'ms': e0.elapsed_time(e1), 'ms_per_step': e0.elapsed_time(e1) / steps, 'host_enqueue_ms_per_step': host_ms}
    # clocks / throttle reasons sampled under load over all timed loops of this workload (device-resident,
    # instrumented, roofline and end-to-end passes): the device-resident loop alone lasts a few ms
    clocks = sampler.finish(t_wall0, time.time()) if with_clocks else None
    # ---- reference GPU leg (reported baseline, rank 0, block workload): the reference's own CUDA
    #      kernels (oracle/_ref/backend_cuda.so) under a restatement of its python glue ----
    ref_gpu = None
    if workload == 'block' and world == 1 and not args.no_cpu_baseline:
        ref_gpu = reference_gpu_leg(dev, coords_dev[0], feats_dev[0], model, flush)
    return {'dev_ms': dev_ms, 'e2e_ms': e2e_ms, 'voxels': float(sum(n_vox[k % 2] for k in range(steps))),
            'launches': int(launches), 'kernels': kern, 'clocks': clocks, 'h2d': h2d, 'd2h': d2h,
            'roof': roof, 'ref_gpu': ref_gpu, 'e2e_ring': e2e_ring, 'e2e_bf16': e2e_bf16,
            'host_enqueue_ms': host_enqueue_ms, 'n0': n_vox[0], 'steps': steps, 'warmup': warmup,
            'n_flush': n_flush, 'flush_us': flush_us}


def run_train(args, steps, warmup, dev, rank, world, local):
    """BASELINE config 3: one DDP training step of ELKEncoder (cr = 1.0, cos (3x7)^3) per rank on its own
    batch of 2 synthetic SemanticKITTI-shaped scans capped at 80 000 voxels each (reference:
    segmentation/train.py:82-100 -- DistributedSampler + DistributedDataParallel(find_unused_parameters=
    True); configs/semantic_kitti/default.yaml: batch_size 2, num_points 80000, SGD momentum 0.9 nesterov,
    weight decay 1e-4; lr 0.024 here -- the reference's 0.24 comes with a warm-up schedule).  Frames never cross ranks; the ONLY collective is DDP's gradient
    all-reduce over NCCL (5.8 M parameters = 23 MB fp32).  Timed region per step, end to end: H2D of
    the batch (coords, feats, labels) from pinned memory -> forward -> cross-entropy -> backward (+
    all-reduce) -> SGD step -> D2H of the loss.  The same steps under `no_sync()` give the step
    without the all-reduce; the difference is the exposed communication time."""
    import contextlib
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from link_b200 import SparseTensor, _capi
    from link_b200.linkencoder import ELKEncoder
    from link_b200.sharding import frame_seed

    per_scan = min(args.voxels, 80_000)
    amp = getattr(args, 'amp', 'none') == 'bf16'
    if amp:
        from link_b200.nn.functional import conv as conv_mod
        conv_mod.set_precision('tf32')
    batches = []
    for b in range(2):                                        # two alternating batches per rank
        cs, fs = [], []
        for j in range(2):
            c, f = make_scan(per_scan, seed=frame_seed(rank, 10 + 2 * b + j))
            c = c.copy()
            c[:, 3] = j
            cs.append(c)
            fs.append(f)
        c, f = np.concatenate(cs), np.concatenate(fs)
        y = np.random.default_rng(frame_seed(rank, b)).integers(0, 19, size=len(c))
        batches.append((torch.from_numpy(c).pin_memory(), torch.from_numpy(f).pin_memory(),
                        torch.from_numpy(y).pin_memory()))
    torch.manual_seed(0)                                      # identical initial weights on every rank
    net = ELKEncoder(num_classes=19, cr=1.0, baseop=BASEOP, r=R_BLK, s=S_BLK, groups=GROUPS).to(dev).train()
    model = net
    if world > 1:
        if os.environ.get('LINKB200_DDP_REFERENCE_STYLE') == '1':      # the reference's own wrapping, for A/B
            model = DDP(net, device_ids=[local], find_unused_parameters=True)
        else:
            from link_b200.sharding import wrap_ddp
            model = wrap_ddp(net, local)                               # unused decoder branches ignored, no graph walk
    opt = torch.optim.SGD([p for p in net.parameters()], lr=0.024, momentum=0.9, weight_decay=1e-4, nesterov=True)
    loss_host = torch.zeros(1).pin_memory()

    def step(i, sync=True):
        c_h, f_h, y_h = batches[i % 2]
        st = SparseTensor.from_host(f_h, c_h, 1, device=dev, ahead=True)   # uploads + coordinate pyramid overlap the previous step's tail
        y = y_h.to(dev, non_blocking=True)
        ctx = contextlib.nullcontext() if (sync or world == 1) else model.no_sync()
        with ctx:
            opt.zero_grad(set_to_none=True)
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
                logits = model(st)
            loss = torch.nn.functional.cross_entropy(logits.float(), y, ignore_index=0)
            loss.backward()
        opt.step()
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])                   # NCCL barrier: communicator exists from here on
        torch.cuda.synchronize()

    def timed_loop(sync):
        for k in range(warmup):
            step(k, sync)
        barrier()
        launches0 = _capi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for k in range(steps):
            step(k, sync)
        e1.record()
        host_ms = (time.perf_counter() - t0) * 1e3 / steps
        barrier()
        return e0.elapsed_time(e1), host_ms, _capi.launch_count() - launches0

    ms, host_ms, launches = timed_loop(True)
    ms_nosync = timed_loop(False)[0] if world > 1 else ms
    loss_val = float(loss_host.item())
    vox = float(sum(len(batches[k % 2][0]) for k in range(steps)))
    h2d = sum(t.numel() * t.element_size() for t in batches[0])
    nparam = sum(p.numel() for p in net.parameters())
    return {'ms': ms, 'ms_nosync': ms_nosync, 'voxels': vox, 'steps': steps, 'warmup': warmup, 'h2d': h2d, 'd2h': 4,
            'launches': int(launches), 'host_ms': host_ms, 'loss': loss_val, 'nparam': nparam,
            'n_batch': len(batches[0][0]), 'amp': amp}


def train_line(tr, world, dev):
    """Whole-job numbers of the training workload (max time over ranks, summed voxels)."""
    from link_b200.sharding import reduce_throughput
    ms, vox = reduce_throughput(tr['ms'], tr['voxels'], None)
    ms_ns, _ = reduce_throughput(tr['ms_nosync'], tr['voxels'], None)
    steps = tr['steps']
    return {'workload': (f"ELKEncoder cos:(3x7)^3 cr=1.0 DDP training step (fwd + cross-entropy + bwd + gradient "
                         f"all-reduce + SGD), 2 synthetic scans per rank, {tr['n_batch']} voxels per batch, "
                         + ('bf16 autocast (dense layers bf16, sparse convs single-pass TF32, fp32 accumulation and master weights)'
                            if tr.get('amp') else 'fp32')),
            'value': vox / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms / steps, 'steps': steps, 'warmup': tr['warmup'],
            'scaling': 'weak', 'collective': f'NCCL all-reduce of {tr["nparam"] * 4 / 1e6:.1f} MB fp32 gradients '
                                             f'per step (torch DDP, {world} ranks)' if world > 1 else 'none (1 rank)',
            'ms_per_step_no_allreduce': ms_ns / steps,
            'allreduce_exposed_ms_per_step': max(0.0, (ms - ms_ns) / steps),
            'timed_region': 'H2D of the batch from pinned memory, forward, loss, backward, all-reduce, SGD step, '
                            'D2H of the loss; CUDA events around the K steps, max over ranks',
            'h2d_bytes_per_step': tr['h2d'], 'd2h_bytes_per_step': tr['d2h'], 'gpu_launches': tr['launches'],
            'host_enqueue_ms_per_step': tr['host_ms'], 'loss': tr['loss']}


def reference_gpu_leg(dev, coords, feats, blk, flush, steps=5, warmup=2):
    try:
        from oracle import ref_gpu
        if not ref_gpu.available():
            return {'unavailable': 'oracle/_ref/backend_cuda.so not built'}
        p = {k: v.detach() for k, v in blk.state_dict().items()}
        evs = []
        with torch.no_grad():
            for k in range(warmup + steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ref_gpu.elk_block_forward(feats, coords, 1, p, S_BLK, R_BLK, BASEOP, GROUPS)
                e1.record()
                if k >= warmup:
                    evs.append((e0, e1))
        torch.cuda.synchronize()
        ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
        n = coords.shape[0]
        return {'value': n / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
                'kind': 'reference CUDA kernels (torchsparse-u backend built for sm_100a from /root/reference, '
                        'oracle/_ref/backend_cuda.so) driven by the restated python glue of oracle/ref_gpu.py; '
                        'same scan, L2 flushed between steps, CUDA events, median'}
    except Exception as e:   # a baseline leg must never take the bench line down
        return {'unavailable': repr(e)[:200]}


def reference_model_leg(what, voxels, steps=5, warmup=2):
    """The unmodified reference (its python package + model class on its own CUDA backend) on the same
    synthetic scan, in its own process (oracle/ref_model_gpu.py)."""
    try:
        p = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_model_gpu.py'), '--what', what, '--voxels',
                            str(voxels), '--steps', str(steps), '--warmup', str(warmup)], capture_output=True, text=True,
                           timeout=300)
        lines = [l for l in p.stdout.splitlines() if l.startswith('{')]
        if p.returncode != 0 or not lines:
            return {'unavailable': (p.stderr or 'no output')[-300:]}
        return json.loads(lines[-1])
    except Exception as e:   # noqa: BLE001 -- a baseline leg must never take the bench line down
        return {'unavailable': repr(e)[:200]}


def preagg_roofline(dev, coords, bounds, blk, nbuf=6, reps=4):
    import ctypes as C
    from link_b200 import SparseTensor, _capi
    from link_b200.elk import block_index, _kernel_gen
    from link_b200.nn.functional import _index
    n = coords.shape[0]
    st = SparseTensor(torch.zeros(n, 1, device=dev), coords, 1)
    _index.set_coord_bounds(st.kmaps, bounds[0], bounds[1])
    bi = block_index(st, S_BLK)
    m = bi.m
    w = blk.pos_weight[0].weight.detach().contiguous().float()
    gen = _kernel_gen(BASEOP, C_BLOCK, w, None, 1.0)
    bufs = [torch.randn(n, C_BLOCK, device=dev) for _ in range(nbuf)]
    sums = torch.zeros(n, 2 * C_BLOCK, device=dev)
    L, stream = _capi.lib(), _capi.stream()

    def launch(i, stream):
        _capi.check(L.lk_link_preagg_seg_fwd(_capi.ptr(bufs[i % nbuf]), _capi.ptr(coords),
                                             _capi.ptr(bi.order), _capi.ptr(bi.sorted_rank), n,
                                             C.byref(gen), _capi.ptr(sums), stream), 'preagg')

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return 1e3 * e0.elapsed_time(e1) / (nbuf * reps)

    for i in range(nbuf):
        launch(i, stream)
    torch.cuda.synchronize()
    # (a) launched one by one from python: each ctypes call costs ~10 us of host time, about the
    #     duration of the kernel, so this figure is partly the host's launch rate
    us_py = timed(lambda: [launch(i, stream) for i in range(nbuf * reps)])
    # (b) the same launches captured once into a CUDA graph and replayed: the device-side duration
    #     per launch (launch latency of back-to-back dependent kernels still included)
    us, how = us_py, 'python launches'
    try:
        gr, cap = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.graph(gr, stream=cap):
            cs = _capi.stream()
            for i in range(nbuf * reps):
                launch(i, cs)
        gr.replay()
        torch.cuda.synchronize()
        us, how = min(timed(gr.replay) for _ in range(3)), 'CUDA-graph replay'
    except Exception as e:                                       # keep the bench line; say what happened
        how = f'python launches (graph capture failed: {type(e).__name__})'
    nbytes = n * (4 * C_BLOCK + 16 + 4) + m * 4 * 2 * C_BLOCK     # F_in + coords + block idx + M sums
    return {'avg_us': us, 'python_launch_avg_us': us_py, 'how': how, 'bytes': float(nbytes),
            'launches': nbuf * reps, 'm': m,
            'inputs': f'{nbuf} rotating [N,{C_BLOCK}] fp32 feature buffers ({nbuf * n * C_BLOCK * 4 / 1e6:.0f} MB > L2)'}


def path_roofline(dev, coords, bounds, blk, nbuf=6, rounds=24):
    """The whole pre-aggregation PATH of SURVEY 8d -- everything between F_input and the pre-LayerNorm
    [N, C] output: zeroing of the block sums, segmented pre-aggregation (kernel generator + block sums),
    window mean over the r^3 neighbour blocks, per-voxel combine -- as ONE CUDA graph of `rounds` rounds
    over `nbuf` rotating input / output buffer sets (> L2), replayed; CUDA events around the replay.
    Algorithmic bytes: N (2*4C + 2*16 + 2*4) + M (2 (4kC + 4))."""
    import ctypes as C
    from link_b200 import SparseTensor, _capi
    from link_b200.elk import block_index, _kernel_gen
    from link_b200.nn.functional import _index
    n = coords.shape[0]
    c, k = C_BLOCK, 2
    st = SparseTensor(torch.zeros(n, 1, device=dev), coords, 1)
    _index.set_coord_bounds(st.kmaps, bounds[0], bounds[1])
    bi = block_index(st, S_BLK)
    m = bi.m
    nbr = bi.neighbors(R_BLK)
    w = blk.pos_weight[0].weight.detach().contiguous().float()
    gen = _kernel_gen(BASEOP, c, w, None, 1.0)
    fin = [torch.randn(n, c, device=dev) for _ in range(nbuf)]
    out = [torch.empty(n, c, device=dev) for _ in range(nbuf)]
    sums = torch.zeros(n, k * c, device=dev)
    mean = torch.zeros(n, k * c, device=dev)
    L, P = _capi.lib(), _capi.ptr

    def one_round(i, s):
        _capi.check(L.lk_zero_rows(P(sums), P(bi.num), n, k * c, s), 'zero')
        _capi.check(L.lk_link_preagg_seg_fwd(P(fin[i % nbuf]), P(coords), P(bi.order), P(bi.sorted_rank), n,
                                             C.byref(gen), P(sums), s), 'preagg')
        _capi.check(L.lk_link_window_mean_seg(P(sums), P(bi.seg), P(nbr), P(bi.num), n, nbr.shape[1], k * c,
                                              P(mean), s), 'wmean')
        _capi.check(L.lk_link_apply_fwd(P(mean), P(fin[i % nbuf]), P(coords), P(bi.idx_query), n, C.byref(gen), 0,
                                        None, None, None, None, None, P(out[i % nbuf]), s), 'apply')
    try:
        for i in range(2):
            one_round(i, _capi.stream())
        torch.cuda.synchronize()
        gr, cap = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.graph(gr, stream=cap):
            cs = _capi.stream()
            for i in range(rounds):
                one_round(i, cs)
        gr.replay()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, 1e3 * e0.elapsed_time(e1) / rounds)
    except Exception as e:   # noqa: BLE001
        return {'unavailable': repr(e)[:200]}
    nbytes = n * (2 * 4 * c + 2 * 16 + 2 * 4) + m * (2 * (4 * k * c + 4))
    return {'avg_us': best, 'bytes': float(nbytes), 'm': m, 'launches_per_round': 4, 'rounds': rounds}


def summarize(m, world, dev):
    """Local measurements -> whole-job numbers (max time over ranks, summed voxels).  e2e = the pipelined
    loop (UploadRing); the single-step latency loop is kept in m['e2e_latency_ms']."""
    from link_b200.sharding import reduce_throughput
    dev_ms, vox = reduce_throughput(m['dev_ms'], m['voxels'], None)     # CPU tensors -> gloo
    lat_ms, _ = reduce_throughput(m['e2e_ms'], m['voxels'], None)
    m['e2e_latency_ms'] = lat_ms
    e2e_ms, _ = reduce_throughput(m['e2e_ring']['ms'] if m.get('e2e_ring') else m['e2e_ms'], m['voxels'], None)
    return dev_ms, e2e_ms, vox


def main_ours(args):
    import torch.distributed as dist
    from link_b200 import _capi

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL is registered for CUDA tensors (a training loop's DDP all-reduce would use it), but the
        # forward hot path has NO data-path collective: the only communication is the timing
        # bookkeeping below, done on CPU tensors over gloo, so no NCCL communicator is ever created
        dist.init_process_group('cpu:gloo,cuda:nccl')
    _capi.lib()
    # pinned staging buffers on the GPU's own NUMA node: bind before anything is pinned (sharding.py)
    from link_b200.sharding import bind_host_to_gpu
    aff0 = os.sched_getaffinity(0)
    affinity = bind_host_to_gpu(local)
    print(f'[bench rank {rank}] host affinity: {affinity}', file=sys.stderr, flush=True)

    if world > 1:
        print(f'[bench rank {rank}] torch.distributed: cuda backend nccl (NCCL {".".join(map(str, torch.cuda.nccl.version()))}), '
              f'cpu backend gloo, world_size {world}', file=sys.stderr, flush=True)
    if args.workload == 'train_encoder':
        tr = run_train(args, args.steps, args.warmup, dev, rank, world, local)
        tl = train_line(tr, world, dev)
        if rank == 0:
            line = {'metric': 'ELKEncoder training voxels/sec (DDP)', 'value': tl['value'], 'unit': UNIT, 'n_gpus': world,
                    'steps': tl['steps'], 'warmup': tl['warmup'], 'ms_per_step': tl['ms_per_step'],
                    'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                    'dtype': 'bf16' if args.amp == 'bf16' else 'f32',
                    'data': 'synthetic', 'config': {'workload': tl['workload']},
                    'e2e': {'value': tl['value'], 'unit': UNIT, 'h2d_bytes_per_step': tl['h2d_bytes_per_step'],
                            'd2h_bytes_per_step': tl['d2h_bytes_per_step'],
                            'note': 'the timed region of this workload IS end to end (host batch in, loss out)'},
                    'gpu_launches': tl['gpu_launches'], 'train': tl}
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return
    m = run_workload(args, args.workload, args.steps, args.warmup, dev, rank, world, local, rank == 0)
    dev_ms, e2e_ms, vox_total = summarize(m, world, dev)
    extra = None
    if args.workload == 'block' and not args.no_encoder:
        # BASELINE config 2 on the same scans, as additional evidence (not the headline metric)
        me = run_workload(args, 'encoder', max(3, args.steps // 2), min(3, args.warmup), dev, rank,
                          world, local, False)
        ed, ee, ev = summarize(me, world, dev)
        extra = {'workload': workload_name_for('encoder', me['n0']), 'value': ev / (ed * 1e-3),
                 'unit': UNIT, 'ms_per_step': ed / me['steps'], 'steps': me['steps'],
                 'e2e': {'value': ev / (ee * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': me['h2d'],
                         'd2h_bytes_per_step': me['d2h']},
                 'gpu_launches': me['launches'], 'host_enqueue_ms_per_step': me['host_enqueue_ms']}
    train = None
    if args.workload == 'block' and not args.no_train:
        # BASELINE config 3 (DDP training step; the one place of the path with a collective), as
        # additional evidence next to the headline metric
        try:
            train = train_line(run_train(args, max(3, args.steps // 2), min(3, args.warmup), dev, rank, world, local),
                               world, dev)
        except Exception as e:   # noqa: BLE001 -- an extra leg must never take the bench line down
            if world > 1:
                raise            # ... but a rank that leaves a collective would hang the others: fail loudly
            train = {'unavailable': repr(e)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = vox_total / (dev_ms * 1e-3)
    e2e_value = vox_total / (e2e_ms * 1e-3)
    steps = m['steps']

    # ---- per-kernel breakdown + roofline of the HBM-bound kernel BASELINE.json names ----------
    peaks = {}
    pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    hbm_peak, peak_src = (peaks.get('hbm_gbs'), 'measured') if peaks.get('hbm_gbs') else (6650.0, 'fallback')
    kern = m['kernels']
    step_us = 1e3 * m['dev_ms'] / steps
    for v in kern.values():
        v['share_of_step'] = v['avg_us'] * v['launches_per_step'] / step_us
        v['gbs'] = v['bytes'] / (v['avg_us'] * 1e-6) / 1e9 if v['bytes'] else None
    roof = None
    if m.get('roof'):
        r = m['roof']
        gbs = r['bytes'] / (r['avg_us'] * 1e-6) / 1e9
        instep = kern.get('lk_link_preagg_fwd', {})
        roof = {'kernel': 'link_preagg_ring_kernel', 'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak,
                'peak_source': peak_src, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                'traffic': NCU_TRAFFIC.get('link_preagg_ring_kernel'), 'traffic_source': NCU_TRAFFIC_SOURCE,
                'algorithmic_bytes_per_launch': r['bytes'], 'avg_us': r['avg_us'],
                'python_launch_avg_us': r['python_launch_avg_us'],
                'in_step_avg_us': instep.get('avg_us'),
                'note': f"{r['launches']} back-to-back launches of the kernel alone over {r['inputs']}, "
                        f"timed by {r['how']} with CUDA events around the batch; python_launch_avg_us = the "
                        'same launches issued one by one from python (host-paced: ~10 us per ctypes call); '
                        'in_step_avg_us = the same kernel inside the step (events around the single '
                        'python-level call, includes launch latency)'}
    roof_path = None
    if m.get('roof') and m['roof'].get('path') and 'avg_us' in m['roof']['path']:
        rp = m['roof']['path']
        gbs = rp['bytes'] / (rp['avg_us'] * 1e-6) / 1e9
        roof_path = {'kernels': ['lk_zero_rows', 'link_preagg_ring_kernel', 'link_window_mean_kernel', 'link_apply_kernel (plain)'],
                     'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                     'algorithmic_bytes_per_round': rp['bytes'], 'avg_us': rp['avg_us'],
                     'note': 'SURVEY 8d pre-aggregation PATH (F_input -> pre-LayerNorm output): 4 launches per round, '
                             f"{rp['rounds']} rounds over rotating buffers (> L2) in one CUDA graph, replayed; bytes = "
                             'N (2*4C + 2*16 + 2*4) + M (2 (4kC + 4))'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
        'warmup': m['warmup'], 'ms_per_step': dev_ms / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_for(args.workload, m['n0'], world),
        'clocks': m['clocks'],
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': m['h2d'],
                'd2h_bytes_per_step': m['d2h'], 'ms_per_step': e2e_ms / steps,
                'how': 'K steps back to back through the public API: UploadRing.upload (pinned host coords + features '
                       '-> device, bounds computed on the host copy) -> ELKBlock -> per-channel checksum -> D2H; the '
                       'upload of scan i+1 overlaps the processing of scan i; ONE CUDA-event pair around the K steps, '
                       'max over ranks',
                'single_step_latency_ms': m['e2e_latency_ms'] / steps},
        'gpu_launches': m['launches'],
        'host_affinity': affinity,
        'host_enqueue_ms_per_step': m['host_enqueue_ms'],
        'l2_flush_memsets_per_step': m.get('n_flush'),
        'e2e_host_enqueue_ms_per_step': (m.get('e2e_ring') or {}).get('host_enqueue_ms_per_step'),
        'e2e_bf16_wire': m.get('e2e_bf16'),
        'roofline': roof,
        'roofline_path': roof_path,
        'kernels': kern,
    }
    if extra is not None:
        if world == 1 and not args.no_cpu_baseline:
            extra['reference_gpu'] = reference_model_leg('encoder', args.voxels)
            if extra['reference_gpu'].get('ms_per_step'):
                extra['vs_reference_gpu'] = extra['reference_gpu']['ms_per_step'] / extra['ms_per_step']
        line['encoder'] = extra
    if train is not None:
        line['train_encoder'] = train
    if m.get('ref_gpu'):
        line['reference_gpu'] = m['ref_gpu']
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, aff0)        # the CPU leg gets every core the process started with
        n_s = args.voxels if args.workload == 'block' else min(args.voxels, 30_000)
        v, dt, n, cores = run_cpu(args, n_s, 2, 1)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': f'2 steps after 1 warm-up of the same workload at N={n} voxels, '
                                          f'{dt:.2f} s/step'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        main_reference(a)
    else:
        main_ours(a)
