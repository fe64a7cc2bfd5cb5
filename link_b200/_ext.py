"""Optional C++ autograd nodes for the per-layer training ops (csrc_ext/autograd_ops.cpp): the same
liblinkb200 kernels as the python autograd Functions, with the per-layer glue (argument marshalling,
the backward nodes) in C++ so that a host-paced training step spends less time in the interpreter.

Built in-tree by `build()` (called from __graft_entry__.build(); torch.utils.cpp_extension, host
compiler only -- no kernels are compiled here) and imported from the built .so; when it is not there
(or LINKB200_CPP_AUTOGRAD=0) the python Functions run -- same kernels, same results."""
import ctypes as C
import glob
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc_ext', 'autograd_ops.cpp')
BUILD_DIR = os.path.join(HERE, 'csrc_ext', '_build')
NAME = 'linkb200_autograd'
ENABLED = os.environ.get('LINKB200_CPP_AUTOGRAD', '1') != '0'
_ENTRY_POINTS = ('lk_conv_tc_pack_weights_ex', 'lk_conv_tc_fwd_plan', 'lk_conv_wgrad_tc', 'lk_bn_train_fwd',
                 'lk_bn_train_bwd', 'lk_last_error')
_mod = None
_tried = False


def _so_path():
    hits = glob.glob(os.path.join(BUILD_DIR, NAME + '*.so'))
    return hits[0] if hits else None


def build(verbose: bool = False) -> str:
    """Compile csrc_ext/autograd_ops.cpp into csrc_ext/_build/ (skipped when the .so is newer than the
    source and the C ABI header)."""
    so = _so_path()
    deps = [SRC, os.path.join(HERE, '..', 'include', 'linkb200.h')]
    if so and os.path.getmtime(so) >= max(os.path.getmtime(d) for d in deps):
        return so
    from torch.utils import cpp_extension
    os.makedirs(BUILD_DIR, exist_ok=True)
    cpp_extension.load(name=NAME, sources=[SRC], build_directory=BUILD_DIR, extra_cflags=['-O2', '-std=c++17'],
                       with_cuda=True, is_python_module=True, verbose=verbose)
    return _so_path()


def module():
    """The extension module with liblinkb200's entry points bound, or None."""
    global _mod, _tried
    if _mod is not None or _tried or not ENABLED:
        return _mod
    _tried = True
    so = _so_path()
    if so is None:
        return None
    try:
        import torch  # noqa: F401  (libtorch must be loaded before the extension)
        spec = importlib.util.spec_from_file_location(NAME, so)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        from link_b200 import _capi
        lib = _capi.lib()
        for name in _ENTRY_POINTS:
            m.bind(name, C.cast(getattr(lib, name), C.c_void_p).value)
        assert m.ready()
        _mod = m
    except Exception as e:       # a stale or foreign build: the python autograd Functions take over
        import warnings
        warnings.warn(f'link_b200: C++ autograd extension not usable ({type(e).__name__}: {e}); '
                      'using the python autograd Functions')
        _mod = None
    return _mod
