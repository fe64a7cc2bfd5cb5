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
