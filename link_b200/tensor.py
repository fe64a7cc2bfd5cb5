"""SparseTensor / PointTensor: the container types of the drop-in surface.

Field-for-field compatible with the reference (torchsparse/tensor.py:10-100): `feats`
[N, C] row-major, `coords` [N, 4] int32 (x, y, z, batch), `stride` 3-tuple, and the two
cache dictionaries `cmaps` / `kmaps` that every tensor derived from the same input SHARES
(conv.py:144-146).  link_b200 stores its own device-side index structures in `kmaps` under
string-tagged keys next to the reference-style keys (see nn/functional/conv.py, elk.py)."""
from typing import Any, Dict, Tuple, Union

import torch

from link_b200.utils import make_ntuple

__all__ = ['SparseTensor', 'PointTensor', 'UploadRing']

_COPY_STREAMS: Dict[Any, Any] = {}


def _copy_stream(device):
    key = (device.type, device.index)
    st = _COPY_STREAMS.get(key)
    if st is None:
        st = _COPY_STREAMS[key] = torch.cuda.Stream(device)
    return st


def _seed_bounds_from_host(st, coords_host: torch.Tensor) -> None:
    """Coordinate bounds of an uploaded scan, computed on the HOST copy while the upload is in flight
    (one pass over [N,4] int32), so the index build never reads them back from the device."""
    if (coords_host.device.type == 'cpu' and coords_host.dim() == 2 and coords_host.shape[0] > 0
            and coords_host.shape[1] == 4 and coords_host.dtype == torch.int32 and coords_host.is_contiguous()):
        import ctypes as C
        from link_b200 import _capi
        from link_b200.nn.functional import _index
        lo, hi = (C.c_int32 * 4)(), (C.c_int32 * 4)()
        _capi.check(_capi.lib().lk_host_coord_bounds(coords_host.data_ptr(), coords_host.shape[0], lo, hi),
                    'lk_host_coord_bounds')                 # ~0.1 ms for 120k voxels; ctypes drops the GIL
        _index.set_coord_bounds(st.kmaps, list(lo), list(hi))


class SparseTensor:

    def __init__(self, feats: torch.Tensor, coords: torch.Tensor,
                 stride: Union[int, Tuple[int, ...]] = 1) -> None:
        self._feats = feats
        self._feats_ready = None       # CUDA event of a pending async upload (from_host), or None
        self._coords_ready = None      # CUDA event of the coordinate upload when it ran on the copy stream (ahead=True)
        self.coords = coords
        self.stride = make_ntuple(stride, ndim=3)
        self.cmaps: Dict[Tuple[int, ...], torch.Tensor] = {}
        self.kmaps: Dict[Tuple[Any, ...], Any] = {}

    # `feats` is a plain attribute in the reference; here reading it also joins a pending
    # asynchronous upload, so every consumer sees complete data without knowing about it.
    @property
    def feats(self) -> torch.Tensor:
        ev = self._feats_ready
        if ev is not None:
            torch.cuda.current_stream(self._feats.device).wait_event(ev)
            self._feats_ready = None
        return self._feats

    @feats.setter
    def feats(self, feats: torch.Tensor) -> None:
        self._feats = feats
        self._feats_ready = None

    def take_feats_event(self):
        """(raw feats tensor, pending upload event or None) WITHOUT waiting: for consumers that
        enqueue index-only work first and wait for the features on the device, in stream order
        (the native block executor)."""
        ev, self._feats_ready = self._feats_ready, None
        return self._feats, ev

    @classmethod
    def from_host(cls, feats: torch.Tensor, coords: torch.Tensor,
                  stride: Union[int, Tuple[int, ...]] = 1, device=None, dtype=None, ahead: bool = False) -> 'SparseTensor':
        """Upload a scan from (pinned) host memory.  The coordinates go first on the current stream;
        the features -- 16x more bytes at C = 64 -- follow on a dedicated copy stream, so the index
        build of the first layer (hash grid, kernel map, conv plan, block sort: ~45 % of a LinK
        block) runs while they are still crossing PCIe.  Consumers join the upload through
        `feats` / `take_feats_event`.

        The step is PCIe-bound end to end, so the features may cross the wire narrower than they are
        computed in: host features in bf16 / fp16 are uploaded as they are (half the bytes) and widened
        to `dtype` (default: kept) on the device, on the copy stream, before the upload event.

        `ahead=True` (training loops, where the current stream still holds the previous step's backward):
        coordinates AND features are uploaded on the copy stream into buffers allocated there, without
        waiting for the current stream, so the upload -- and the index work that depends only on the
        coordinates (ELKEncoder.plan_levels builds the coordinate pyramid on side streams) -- overlaps
        the tail of the previous step.  The current stream waits for the coordinates by an event; the
        buffers are handed to it with `record_stream`."""
        device = torch.device(device if device is not None else ('cuda', torch.cuda.current_device()))
        main = torch.cuda.current_stream(device)
        copy_stream = _copy_stream(device)
        widen = dtype is not None and dtype != feats.dtype
        if ahead:
            with torch.cuda.stream(copy_stream):
                c_dev = torch.empty(coords.shape, dtype=coords.dtype, device=device)
                c_dev.copy_(coords, non_blocking=True)
                ev_c = torch.cuda.Event()
                ev_c.record(copy_stream)
                f_wire = torch.empty(feats.shape, dtype=feats.dtype, device=device)
                f_wire.copy_(feats, non_blocking=True)
                f_dev = f_wire.to(dtype) if widen else f_wire
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            c_dev.record_stream(main)
            f_dev.record_stream(main)
            main.wait_event(ev_c)          # every later consumer of the coordinates on this stream is ordered
            st = cls(f_dev, c_dev, stride)
            st._feats_ready = ev
            st._coords_ready = ev_c
            _seed_bounds_from_host(st, coords)
            return st
        c_dev = coords.to(device, non_blocking=True)
        f_wire = torch.empty(feats.shape, dtype=feats.dtype, device=device)
        f_dev = torch.empty(feats.shape, dtype=dtype, device=device) if widen else f_wire
        copy_stream.wait_stream(main)      # the buffers may recycle memory still in use by queued kernels
        with torch.cuda.stream(copy_stream):
            f_wire.copy_(feats, non_blocking=True)
            if widen:
                f_dev.copy_(f_wire)
                f_wire.record_stream(copy_stream)
            # the buffers were allocated on the main stream but are written here: if the tensor is dropped
            # before a consumer joins the upload, the allocator must not hand the block out while the copy runs
            f_dev.record_stream(copy_stream)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        st = cls(f_dev, c_dev, stride)
        st._feats_ready = ev
        _seed_bounds_from_host(st, coords)
        return st

    @property
    def F(self) -> torch.Tensor:
        return self.feats

    @F.setter
    def F(self, feats: torch.Tensor) -> None:
        self.feats = feats

    @property
    def C(self) -> torch.Tensor:
        return self.coords

    @C.setter
    def C(self, coords: torch.Tensor) -> None:
        self.coords = coords

    @property
    def s(self) -> Tuple[int, ...]:
        return self.stride

    @s.setter
    def s(self, stride: Union[int, Tuple[int, ...]]) -> None:
        self.stride = make_ntuple(stride, ndim=3)

    def cuda(self):
        self.feats = self.feats.cuda()
        self.coords = self.coords.cuda()
        return self

    def detach(self):
        self.feats = self.feats.detach()
        self.coords = self.coords.detach()
        return self

    def to(self, device, non_blocking: bool = True):
        self.feats = self.feats.to(device, non_blocking=non_blocking)
        self.coords = self.coords.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        output = SparseTensor(coords=self.coords, feats=self.feats + other.feats,
                              stride=self.stride)
        output.cmaps = self.cmaps
        output.kmaps = self.kmaps
        return output


class PointTensor:

    def __init__(self, feats, coords, idx_query=None, weights=None):
        self.F = feats
        self.C = coords
        self.idx_query = idx_query if idx_query is not None else {}
        self.weights = weights if weights is not None else {}
        self.additional_features = {'idx_query': {}, 'counts': {}}

    def cuda(self):
        self.F = self.F.cuda()
        self.C = self.C.cuda()
        return self

    def detach(self):
        self.F = self.F.detach()
        self.C = self.C.detach()
        return self

    def to(self, device, non_blocking=True):
        self.F = self.F.to(device, non_blocking=non_blocking)
        self.C = self.C.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        tensor = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        tensor.additional_features = self.additional_features
        return tensor


class UploadRing:
    """Caller-owned staging ring for a stream of scans arriving from pinned host memory: the feature
    upload of scan i+1 crosses PCIe while scan i is still being processed, without going through
    the caching allocator (a `record_stream`-managed buffer is never ready for reuse while the host
    runs ahead of the device, so every upload would pay a `cudaMalloc`).

    `depth` device buffers of `capacity` rows are reused round-robin.  Contract: everything that
    consumes the SparseTensor returned by `upload()` must be ENQUEUED on the current stream before
    the next `upload()` call, and the returned tensors alias ring memory that is overwritten
    `depth` uploads later (clone what must live longer).  Ordering is by events only: upload i
    waits (on the copy stream) for the compute-stream event recorded at the start of upload
    i - depth + 1, which is after the consumers of upload i - depth were enqueued."""

    def __init__(self, capacity: int, channels: int, device=None, depth: int = 2, dtype=torch.float32):
        assert depth >= 2 and capacity > 0 and channels > 0
        self.device = torch.device(device if device is not None else ('cuda', torch.cuda.current_device()))
        self.depth = depth
        self._feats = [torch.empty(capacity, channels, dtype=dtype, device=self.device) for _ in range(depth)]
        self._coords = [torch.empty(capacity, 4, dtype=torch.int32, device=self.device) for _ in range(depth)]
        self._issued = []                # compute-stream events, one per upload() call (last `depth` kept)
        self._count = 0

    def upload(self, feats: torch.Tensor, coords: torch.Tensor, stride: Union[int, Tuple[int, ...]] = 1) -> 'SparseTensor':
        n = feats.shape[0]
        assert n <= self._feats[0].shape[0] and coords.shape[0] == n, 'scan larger than the ring capacity'
        main = torch.cuda.current_stream(self.device)
        copy_stream = _copy_stream(self.device)
        slot = self._count % self.depth
        mark = torch.cuda.Event()
        mark.record(main)                # everything enqueued so far (consumers of earlier uploads) precedes it
        self._issued.append(mark)
        if len(self._issued) > self.depth:
            self._issued.pop(0)
        if self._count >= self.depth:
            # the slot was last used by upload count - depth; its consumers were enqueued before upload
            # count - depth + 1 started, i.e. before the oldest event still held
            copy_stream.wait_event(self._issued[0])
        f_dev, c_dev = self._feats[slot][:n], self._coords[slot][:n]
        c_dev.copy_(coords, non_blocking=True)                   # small: on the compute stream, needed first
        with torch.cuda.stream(copy_stream):
            f_dev.copy_(feats, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        self._count += 1
        st = SparseTensor(f_dev, c_dev, stride)
        st._feats_ready = ev
        _seed_bounds_from_host(st, coords)
        return st


class _PinnedBlock:
    """Owner of one cudaHostAlloc block (freed when the last tensor viewing it is gone)."""

    def __init__(self, nbytes: int, write_combined: bool):
        import ctypes as C
        from link_b200 import _capi
        p = C.c_void_p()
        _capi.check(_capi.lib().lk_host_alloc(nbytes, 1 if write_combined else 0, C.byref(p)), 'lk_host_alloc')
        self.ptr, self.nbytes = p.value, nbytes

    def __del__(self):
        try:
            from link_b200 import _capi
            _capi.lib().lk_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=torch.float32, write_combined: bool = False) -> torch.Tensor:
    """Uninitialised pinned host tensor (cudaHostAlloc through liblinkb200).  `write_combined=True` is for
    staging buffers the host only WRITES before they are uploaded (features on their way to
    `SparseTensor.from_host` / `UploadRing.upload`): the memory is uncached on the CPU side, so host READS of
    it are very slow -- keep coordinates (the host computes their bounds) in ordinary pinned memory."""
    import ctypes as C
    import numpy as np
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
    n = 1
    for v in shape:
        n *= v
    itemsize = torch.empty(0, dtype=dtype).element_size()
    block = _PinnedBlock(max(n * itemsize, 1), write_combined)
    buf = (C.c_uint8 * (n * itemsize)).from_address(block.ptr)
    buf._lk_owner = block                                  # the ctypes array keeps the allocation alive
    t = torch.frombuffer(buf, dtype=torch.uint8).view(dtype).view(shape) if n else torch.empty(shape, dtype=dtype)
    return t
