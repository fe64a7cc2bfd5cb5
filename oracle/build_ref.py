"""Build recipe for oracle/_ref: the reference's OWN CPU backend, compiled in place.

TEST INFRASTRUCTURE ONLY.  Compiles the reference sources where they lie under
/root/reference/segmentation/torchsparse-u/torchsparse/backend (pybind_cpu.cpp and
every *_cpu.cpp) with g++ directly -- the reference's setup.py is not run -- and
writes one shared object into oracle/_ref/.  Nothing is copied into the repo.

The only accommodation is oracle/shim/google/dense_hash_map (google-sparsehash is
absent from the image).  libtorch headers come from the installed torch wheel,
which is also present on the GPU box, so the built .so travels with gpurun.
"""
import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/segmentation/torchsparse-u/torchsparse/backend'
OUT = os.path.join(HERE, '_ref', 'backend.so')


def build(force: bool = False) -> str:
    if not os.path.isdir(REF):
        # GPU box: the reference tree is not shipped; use the prebuilt file.
        return OUT if os.path.exists(OUT) else ''
    srcs = [os.path.join(REF, 'pybind_cpu.cpp')] + sorted(
        glob.glob(os.path.join(REF, '*', '*_cpu.cpp')))
    if (not force and os.path.exists(OUT) and
            all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs)):
        return OUT
    import torch
    from torch.utils.cpp_extension import include_paths
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    cmd = ['g++', '-O3', '-g0', '-fopenmp', '-shared', '-fPIC', '-std=c++17',
           '-DTORCH_EXTENSION_NAME=backend', '-DTORCH_API_INCLUDE_EXTENSION_H',
           f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}',
           '-I' + os.path.join(HERE, 'shim'),
           '-I' + sysconfig.get_paths()['include']]
    cmd += ['-I' + p for p in include_paths()]
    cmd += srcs + ['-L' + tlib, '-ltorch', '-ltorch_cpu', '-lc10',
                   '-ltorch_python', '-Wl,-rpath,' + tlib, '-o', OUT]
    print(' '.join(cmd))
    subprocess.check_call(cmd)
    return OUT


OUT_CUDA = os.path.join(HERE, '_ref', 'backend_cuda.so')


def build_cuda(force: bool = False, jobs: int = 8) -> str:
    """The reference's own CUDA backend (pybind_cuda.cpp + every *_cuda.cu + the *_cpu.cpp it also
    binds), compiled for sm_100a into oracle/_ref/backend_cuda.so.  nvcc cross-compiles here without a
    GPU; the GPU box only uses the prebuilt file.  torch 2.11 no longer converts
    `Tensor::type()` to a ScalarType inside AT_DISPATCH_*, so the sources are compiled from a
    scratch copy under /tmp in which the 9 `x.type()` dispatch arguments read `x.scalar_type()`
    (convolution_cuda.cu, devoxelize_cuda.cu, voxelize_cuda.cu) -- no other change, nothing copied
    into the repo."""
    if not os.path.isdir(REF):
        return OUT_CUDA if os.path.exists(OUT_CUDA) else ''
    if not force and os.path.exists(OUT_CUDA):
        return OUT_CUDA
    import concurrent.futures as cf
    import re
    import shutil
    import tempfile
    import torch
    from torch.utils.cpp_extension import include_paths
    tmp = tempfile.mkdtemp(prefix='lk_ref_cuda_')
    src = os.path.join(tmp, 'backend')
    shutil.copytree(REF, src)
    for f in glob.glob(os.path.join(src, '*', '*_cuda.cu')):
        txt = open(f).read()
        txt2 = re.sub(r'(AT_DISPATCH_FLOATING_TYPES_AND_HALF\(\s*[A-Za-z_]+)\.type\(\)', r'\1.scalar_type()', txt)
        if txt2 != txt:
            open(f, 'w').write(txt2)
    os.makedirs(os.path.dirname(OUT_CUDA), exist_ok=True)
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    incs = ['-I' + os.path.join(HERE, 'shim'), '-I' + sysconfig.get_paths()['include'],
            '-I/usr/local/cuda/include'] + ['-I' + p for p in include_paths()]
    defs = ['-DTORCH_EXTENSION_NAME=backend', '-DTORCH_API_INCLUDE_EXTENSION_H',
            f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}']
    cus = sorted(glob.glob(os.path.join(src, '*', '*_cuda.cu')))
    cpps = [os.path.join(src, 'pybind_cuda.cpp')] + sorted(glob.glob(os.path.join(src, '*', '*_cpu.cpp')))

    def cc(f):
        obj = f + '.o'
        if f.endswith('.cu'):
            cmd = ['/usr/local/cuda/bin/nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3',
                   '-std=c++17', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
                   '-D__CUDA_NO_HALF_OPERATORS__', '-D__CUDA_NO_HALF_CONVERSIONS__',
                   '-D__CUDA_NO_HALF2_OPERATORS__'] + defs + incs + ['-c', f, '-o', obj]
        else:
            cmd = ['g++', '-O3', '-g0', '-fopenmp', '-fPIC', '-std=c++17'] + defs + incs + ['-c', f, '-o', obj]
        subprocess.check_call(cmd)
        return obj

    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = list(ex.map(cc, cus + cpps))
    cmd = ['g++', '-shared', '-fopenmp'] + objs + ['-L' + tlib, '-L/usr/local/cuda/lib64', '-ltorch', '-ltorch_cpu',
                                       '-ltorch_cuda', '-lc10', '-lc10_cuda', '-ltorch_python', '-lcudart',
                                       '-Wl,-rpath,' + tlib, '-o', OUT_CUDA]
    subprocess.check_call(cmd)
    shutil.rmtree(tmp, ignore_errors=True)
    return OUT_CUDA


def load_cuda():
    """Import oracle/_ref/backend_cuda.so as a module named `backend` (20 reference ops)."""
    import importlib.machinery
    import importlib.util
    import torch  # noqa: F401
    if not os.path.exists(OUT_CUDA):
        raise FileNotFoundError(OUT_CUDA)
    loader = importlib.machinery.ExtensionFileLoader('backend', OUT_CUDA)
    spec = importlib.util.spec_from_loader('backend', loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def load():
    """Import the built module as `torchsparse.backend` (name the reference's
    python layer expects) without touching /root/reference."""
    import importlib.machinery
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    if not os.path.exists(OUT):
        raise FileNotFoundError(OUT)
    loader = importlib.machinery.ExtensionFileLoader('torchsparse.backend', OUT)
    spec = importlib.util.spec_from_loader('torchsparse.backend', loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
    if '--cuda' in sys.argv:
        print(build_cuda(force='--force' in sys.argv))


IOU3D_SRC = '/root/reference/detection/det3d/ops/iou3d_nms/src/iou3d_cpu.cpp'
OUT_IOU3D = os.path.join(HERE, '_ref', 'iou3d_cpu_ref.so')


def build_iou3d(force: bool = False) -> str:
    """The reference's own CPU rotated-BEV-IoU (det3d/ops/iou3d_nms/src/iou3d_cpu.cpp, the CPU twin
    of the iou3d_nms CUDA kernels), compiled in place with g++ plus oracle/bind/iou3d_bind.cpp (our
    pybind stub: the reference's own binding file also binds the CUDA entry points).  The source
    marks its Point methods `__device__` although it is a .cpp: that macro is defined away."""
    if not os.path.exists(IOU3D_SRC):
        return OUT_IOU3D if os.path.exists(OUT_IOU3D) else ''
    if not force and os.path.exists(OUT_IOU3D):
        return OUT_IOU3D
    import torch
    from torch.utils.cpp_extension import include_paths
    os.makedirs(os.path.dirname(OUT_IOU3D), exist_ok=True)
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    cmd = ['g++', '-O2', '-g0', '-shared', '-fPIC', '-std=c++17', '-w', '-D__device__=',
           '-DTORCH_EXTENSION_NAME=iou3d_cpu_ref', '-DTORCH_API_INCLUDE_EXTENSION_H',
           f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}',
           '-I' + sysconfig.get_paths()['include'], '-I/usr/local/cuda/include']
    cmd += ['-I' + p for p in include_paths()]
    cmd += [IOU3D_SRC, os.path.join(HERE, 'bind', 'iou3d_bind.cpp'), '-L' + tlib, '-ltorch', '-ltorch_cpu', '-lc10',
            '-ltorch_python', '-Wl,-rpath,' + tlib, '-o', OUT_IOU3D]
    print(' '.join(cmd))
    subprocess.check_call(cmd)
    return OUT_IOU3D


def load_iou3d():
    import importlib.util
    path = build_iou3d()
    if not path:
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location('iou3d_cpu_ref', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------------------------------------------
# The reference's PYTHON layer, staged for the GPU box (where /root/reference does not exist): the
# torchsparse package WITHOUT its backend/ sources, and the two model files of the hot path.  They
# go under baseline/_ref/ (git-ignored like the reference install of the base contract, shipped by
# gpurun), untouched.  tests/test_gpu_reference_python_on_shim.py runs them on top of
# link_b200.backend: the reference's own nn/functional glue and its own ELKBlock / ELKEncoder, our kernels.
PY_STAGE = os.path.join(os.path.dirname(HERE), 'baseline', '_ref', 'py')
_SEG = '/root/reference/segmentation'


def stage_python(force: bool = False) -> str:
    import shutil
    if not os.path.isdir(_SEG):
        return PY_STAGE if os.path.isdir(os.path.join(PY_STAGE, 'torchsparse')) else ''
    if os.path.isdir(os.path.join(PY_STAGE, 'torchsparse')) and not force:
        return PY_STAGE
    shutil.rmtree(PY_STAGE, ignore_errors=True)
    os.makedirs(PY_STAGE)
    shutil.copytree(os.path.join(_SEG, 'torchsparse-u', 'torchsparse'), os.path.join(PY_STAGE, 'torchsparse'),
                    ignore=shutil.ignore_patterns('backend', '__pycache__', '*.so'))
    dst = os.path.join(PY_STAGE, 'core', 'models', 'semantic_kitti')
    os.makedirs(dst)
    for pkg in ('core', os.path.join('core', 'models'), os.path.join('core', 'models', 'semantic_kitti')):
        open(os.path.join(PY_STAGE, pkg, '__init__.py'), 'w').close()       # empty packages: only two modules are used
    shutil.copy(os.path.join(_SEG, 'core', 'models', 'utils.py'), os.path.join(PY_STAGE, 'core', 'models', 'utils.py'))
    shutil.copy(os.path.join(_SEG, 'core', 'models', 'semantic_kitti', 'linkencoder.py'),
                os.path.join(dst, 'linkencoder.py'))
    return PY_STAGE
