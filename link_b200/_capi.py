"""ctypes binding of liblinkb200.so (C ABI declared in include/linkb200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Tensors are passed as raw device pointers together with torch's
current CUDA stream; all memory (outputs, workspaces) is allocated by the caller as torch
tensors, so the library itself never allocates or synchronises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('LINKB200_LIB') or os.path.join(_HERE, 'liblinkb200.so')

_lib = None

c_i32p = C.c_void_p
vp = C.c_void_p
i64 = C.c_int64
i32 = C.c_int


class KeySpec(C.Structure):
    _fields_ = [('div', C.c_int32 * 3), ('mul', C.c_int32 * 3), ('order', C.c_int32 * 4),
                ('lo', C.c_int32 * 4), ('bits', C.c_int32 * 4)]


class KernelGen(C.Structure):
    _fields_ = [('op', C.c_int32), ('c', C.c_int32), ('wrows', C.c_int32),
                ('coord_scale', C.c_float), ('d_pos_weight', C.c_void_p),
                ('d_alpha', C.c_void_p), ('accurate_trig', C.c_int32), ('reserved', C.c_int32)]


class ConvEpilogue(C.Structure):
    _fields_ = [('d_scale', C.c_void_p), ('d_shift', C.c_void_p), ('d_residual', C.c_void_p),
                ('relu', C.c_int32), ('precision', C.c_int32)]


class ElkBlockArgs(C.Structure):
    _fields_ = [('n', C.c_int64), ('d_coords', C.c_void_p), ('d_feats', C.c_void_p),
                ('d_out', C.c_void_p), ('d_premix_w', C.c_void_p), ('d_premix_g', C.c_void_p),
                ('d_premix_b', C.c_void_p), ('premix_eps', C.c_float), ('kvol', C.c_int32),
                ('d_conv_w', C.c_void_p), ('d_conv_wt', C.c_void_p), ('d_conv_offsets', C.c_void_p),
                ('d_kmap', C.c_void_p), ('build_kmap', C.c_int32), ('build_plan', C.c_int32),
                ('d_plan_perm', C.c_void_p), ('d_plan_mask', C.c_void_p),
                ('keyspec', KeySpec), ('key_bits', C.c_int32), ('r3', C.c_int32),
                ('d_block_offsets', C.c_void_p), ('gen', KernelGen),
                ('d_g1', C.c_void_p), ('d_b1', C.c_void_p), ('d_g2', C.c_void_p), ('d_b2', C.c_void_p),
                ('use_tensor_cores', C.c_int32), ('conv_precision', C.c_int32),
                ('d_ws', C.c_void_p), ('ws_bytes', C.c_int64), ('single_stream', C.c_int32),
                ('reserved1', C.c_int32), ('feats_ready', C.c_void_p)]


ENC_MAX_LEVELS = 4


class ConvLayer(C.Structure):
    """lk_conv_layer_t"""
    _fields_ = [('d_wimg', C.c_void_p), ('d_scale', C.c_void_p), ('d_shift', C.c_void_p),
                ('c_in', C.c_int32), ('c_out', C.c_int32), ('relu', C.c_int32), ('reserved', C.c_int32)]


class EncLevel(C.Structure):
    """lk_enc_level_t"""
    _fields_ = [('down', ConvLayer), ('stage', ConvLayer * 4), ('tail', ConvLayer), ('elk_tail', ConvLayer),
                ('elk', ElkBlockArgs), ('down_spec', KeySpec), ('down_bits', C.c_int32), ('reserved', C.c_int32),
                ('d_off2', C.c_void_p), ('d_off3', C.c_void_p), ('d_out', C.c_void_p), ('d_coords', C.c_void_p)]


class ElkEncoderArgs(C.Structure):
    """lk_elk_encoder_args_t"""
    _fields_ = [('n0', C.c_int64), ('d_coords0', C.c_void_p), ('d_feats0', C.c_void_p), ('feats_ready', C.c_void_p),
                ('levels', C.c_int32), ('c_max', C.c_int32), ('conv_precision', C.c_int32),
                ('single_stream', C.c_int32), ('overlap_branches', C.c_int32), ('reserved', C.c_int32),
                ('stem', ConvLayer * 2), ('d_off3_0', C.c_void_p), ('d_out0', C.c_void_p),
                ('level', EncLevel * ENC_MAX_LEVELS), ('d_ws', C.c_void_p), ('ws_bytes', C.c_int64),
                ('n_out', C.c_int64 * (ENC_MAX_LEVELS + 1))]


# name -> (restype, argtypes); must list every symbol declared in include/linkb200.h
PROTOTYPES = {
    'lk_last_error': (C.c_char_p, []),
    'lk_version': (i32, []),
    'lk_launch_count': (i64, []),
    'lk_device_pci_bus_id': (i32, [i32, C.c_char_p, i32]),
    'lk_host_alloc': (i32, [i64, i32, C.POINTER(C.c_void_p)]),
    'lk_host_free': (i32, [vp]),
    'lk_hash': (i32, [vp, i64, vp, vp]),
    'lk_kernel_hash': (i32, [vp, i64, vp, i32, vp, vp]),
    'lk_table_capacity': (i64, [i64]),
    'lk_table_build': (i32, [vp, i64, vp, i64, vp]),
    'lk_table_build_coords': (i32, [vp, i64, vp, i64, vp]),
    'lk_table_query': (i32, [vp, i64, vp, i64, vp, vp]),
    'lk_hash_div': (i32, [vp, i64, i32, vp, vp]),
    'lk_table_query_div': (i32, [vp, i64, i32, vp, i64, vp, vp]),
    'lk_count': (i32, [vp, i64, vp, i64, vp]),
    'lk_voxelize_fwd': (i32, [vp, vp, vp, i64, i64, i32, vp, vp]),
    'lk_voxelize_bwd': (i32, [vp, vp, vp, i64, i64, i32, vp, vp]),
    'lk_devoxelize_fwd': (i32, [vp, vp, vp, i64, i32, i32, vp, vp]),
    'lk_devoxelize_bwd': (i32, [vp, vp, vp, i64, i32, i32, i64, vp, vp]),
    'lk_gather_concat': (i32, [vp, vp, i32, i32, i64, vp, i32, vp, vp]),
    'lk_pack_keys': (i32, [vp, i64, C.POINTER(KeySpec), vp, vp]),
    'lk_unpack_keys': (i32, [vp, vp, i64, C.POINTER(KeySpec), vp, vp]),
    'lk_sort_unique_ws_bytes': (i64, [i64]),
    'lk_sort_unique': (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    'lk_sort_unique_ex': (i32, [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    'lk_block_neighbors': (i32, [vp, vp, i64, C.POINTER(KeySpec), vp, i32, vp, vp]),
    'lk_block_neighbors_zero': (i32, [vp, vp, i64, C.POINTER(KeySpec), vp, i32, vp, vp, i32, vp]),
    'lk_sort_unique_coords': (i32, [vp, C.POINTER(KeySpec), i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    'lk_zero_rows': (i32, [vp, vp, i64, i32, vp]),
    'lk_link_preagg_fwd': (i32, [vp, vp, vp, i64, C.POINTER(KernelGen), vp, vp]),
    'lk_link_preagg_seg_fwd': (i32, [vp, vp, vp, vp, i64, C.POINTER(KernelGen), vp, vp]),
    'lk_link_window_mean': (i32, [vp, vp, vp, vp, i64, i32, i32, vp, vp]),
    'lk_link_window_mean_seg': (i32, [vp, vp, vp, vp, i64, i32, i32, vp, vp]),
    'lk_link_apply_fwd': (i32, [vp, vp, vp, vp, i64, C.POINTER(KernelGen), i32, vp, vp, vp, vp,
                                vp, vp, vp]),
    'lk_link_bwd_supported': (i32, [i32]),
    'lk_link_window_mean_tot': (i32, [vp, vp, vp, vp, i64, i32, i32, vp, vp, vp]),
    'lk_link_bwd_norm': (i32, [vp, vp, vp, vp, vp, i64, vp, C.POINTER(KernelGen), vp, vp, vp, vp, vp, vp, vp, vp,
                               vp, vp, vp]),
    'lk_link_bwd_apply': (i32, [vp, vp, vp, vp, vp, vp, i64, i32, vp, C.POINTER(KernelGen), vp, vp, vp, vp, vp]),
    'lk_abi_sizeof': (i32, [i32]),
    'lk_elk_block_ws_bytes': (i64, [i64, i32, i32, i32, i32, i32]),
    'lk_elk_block_fwd': (i32, [C.POINTER(ElkBlockArgs), vp]),
    'lk_elk_encoder_ws_bytes': (i64, [i64, i32, i32, i32, i32]),
    'lk_elk_encoder_fwd': (i32, [C.POINTER(ElkEncoderArgs), vp]),
    'lk_bn_supported': (i32, [i32]),
    'lk_bn_ws_bytes': (i64, [i32]),
    'lk_bn_train_fwd': (i32, [vp, vp, i64, i32, vp, vp, C.c_float, C.c_float, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    'lk_bn_train_bwd': (i32, [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]),
    'lk_linear_ln_fwd': (i32, [vp, vp, vp, vp, C.c_float, i64, i32, vp, vp]),
    'lk_linear_ln_tc_fwd': (i32, [vp, vp, vp, vp, C.c_float, i64, i32, vp, vp]),
    'lk_kmap_query': (i32, [vp, i64, vp, i32, vp, i64, vp, vp]),
    'lk_kmap_query_subm': (i32, [vp, i64, vp, i32, vp, i64, vp, vp]),
    'lk_kmap_query_subm_ev': (i32, [vp, i64, vp, i32, vp, i64, vp, vp, vp]),
    'lk_kmap_invert': (i32, [vp, i64, i32, i64, vp, vp]),
    'lk_conv_fwd': (i32, [vp, vp, vp, i64, i32, i32, i32, vp, vp, vp]),
    'lk_conv_fwd_ex': (i32, [vp, vp, vp, i64, i32, i32, i32, C.POINTER(ConvEpilogue), vp, vp]),
    'lk_conv_tc_fwd_ex': (i32, [vp, vp, vp, i64, i32, i32, i32, C.POINTER(ConvEpilogue), vp, vp]),
    'lk_kmap_build_ws_bytes': (i64, [i64, i64]),
    'lk_kmap_build': (i32, [vp, i64, vp, i64, vp, i32, i32, vp, i64, i32, vp, vp, vp, vp, i64, vp]),
    'lk_downsample_ws_bytes': (i64, [i64]),
    'lk_downsample': (i32, [vp, i64, C.POINTER(KeySpec), i32, vp, vp, vp, i64, vp]),
    'lk_points_to_voxel_ws_bytes': (i64, [i64]),
    'lk_points_to_voxel': (i32, [vp, i64, i32, vp, vp, i32, i32, vp, vp, vp, vp, vp, i64, vp]),
    'lk_conv_plan_ws_bytes': (i64, [i64]),
    'lk_conv_plan': (i32, [vp, i64, i32, vp, vp, vp, vp, i64, vp]),
    'lk_conv_tc_pack_weights': (i32, [vp, i32, i32, i32, vp, vp]),
    'lk_conv_tc_pack_weights_ex': (i32, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    'lk_conv_tc_fwd_plan': (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, C.POINTER(ConvEpilogue), vp, vp]),
    'lk_conv_tc_supported': (i32, [i32, i32]),
    'lk_conv_tc_bf16_supported': (i32, [i32, i32]),
    'lk_conv_tc_fwd_bf16': (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, C.POINTER(ConvEpilogue), vp, vp]),
    'lk_conv_tc_fwd': (i32, [vp, vp, vp, i64, i32, i32, i32, vp, vp, vp]),
    'lk_conv_bwd_weight': (i32, [vp, vp, vp, i64, i32, i32, i32, vp, vp]),
    'lk_host_coord_bounds': (i32, [vp, i64, vp, vp]),
    'lk_strided_candidates': (i32, [vp, i64, vp, vp, vp, vp, i32, vp, vp, vp]),
    'lk_kmap_from_pairs': (i32, [vp, vp, i32, i64, i32, i32, vp, vp]),
    'lk_bev_scatter': (i32, [vp, vp, i64, i32, i32, i32, i32, i32, vp, vp]),
    'lk_bev_gather': (i32, [vp, vp, i64, i32, i32, i32, i32, i32, vp, vp]),
    'lk_conv_wgrad_tc_supported': (i32, [i32, i32]),
    'lk_conv_wgrad_prepass': (i32, [vp, vp, i64, i32, vp, vp, vp]),
    'lk_conv_wgrad_tc': (i32, [vp, vp, vp, vp, vp, i64, i32, i32, i32, vp, i32, vp]),
    'lk_boxes_iou_bev': (i32, [vp, i64, vp, i64, vp, vp]),
    'lk_nms_bev_ws_bytes': (i64, [i64]),
    'lk_nms_bev': (i32, [vp, i64, C.c_float, vp, i64, vp, vp]),
    'lk_nms_circle': (i32, [vp, i64, C.c_float, vp, i64, vp, vp]),
    'lk_boxes_iou_bev_hostcheck': (i32, [vp, i64, vp, i64, vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing: build it with `python -m link_b200._build` '
                '(link_b200 has no CPU or PyTorch fallback)')
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
        for which, struct in enumerate((KeySpec, KernelGen, ElkBlockArgs, ConvLayer, EncLevel, ElkEncoderArgs)):
            if _lib.lk_abi_sizeof(which) != C.sizeof(struct):
                raise RuntimeError(f'ABI mismatch: {struct.__name__} is {C.sizeof(struct)} bytes here, '
                                   f'{_lib.lk_abi_sizeof(which)} in liblinkb200.so -- rebuild the library')
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().lk_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream() -> int:
    """Handle of torch's current CUDA stream on the current device (raw cudaStream_t as int)."""
    if _raw_stream is not None:      # ~0.3 us; torch.cuda.current_stream() costs ~8 us per call
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('link_b200 kernels need CUDA tensors (there is no CPU fallback); got '
                           f'a {t.device} tensor')
    if not t.is_contiguous():
        raise RuntimeError('link_b200 kernels need contiguous tensors')
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f'expected dtype {dtype}, got {t.dtype}')
    return t.data_ptr()


def check_device(t) -> None:
    """liblinkb200 launches on the CURRENT device's current stream: a tensor that lives on another GPU
    would be read through a foreign stream (or fault).  Called once per high-level op."""
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise RuntimeError(f'tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: '
                           'wrap the call in `with torch.cuda.device(tensor.device):`')


def launch_count() -> int:
    return int(lib().lk_launch_count())


# Optional per-kernel timing (bench.py): when TIMERS is a dict, `timed(name, bytes)` brackets a
# library call with CUDA events on the launching stream and appends (start, end, algorithmic
# bytes) to TIMERS[name].  Disabled (None) it costs one attribute test.
TIMERS = None


class timed:
    __slots__ = ('name', 'nbytes', 'e0')

    def __init__(self, name: str, nbytes: float = 0.0):
        self.name, self.nbytes, self.e0 = name, nbytes, None

    def __enter__(self):
        if TIMERS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and TIMERS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            TIMERS.setdefault(self.name, []).append((self.e0, e1, float(self.nbytes)))
        return False
