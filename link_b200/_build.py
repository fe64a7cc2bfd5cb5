"""In-tree build of liblinkb200.so: nvcc -gencode arch=compute_100a,code=sm_100a, one object
per .cu compiled in parallel, linked into link_b200/liblinkb200.so (git-ignored, shipped to
the GPU box by gpurun).  No torch headers are involved: the library is a plain C ABI."""
import concurrent.futures as cf
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liblinkb200.so')
OBJ = os.path.join(HERE, 'csrc', '_obj')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC,-msse4.1', '-Xptxas', '-O3']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _deps_mtime():
    hs = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(
        os.path.join(HERE, '..', 'include', '*.h'))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), _deps_mtime()):
        return obj
    cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sources()
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJ, '*.o')):
            os.remove(f)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if (force or not os.path.exists(LIB)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = [NVCC, '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + objs
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


def build_variant(name: str, defines, verbose: bool = False, source: str = 'link.cu') -> str:
    """Tuning aid: liblinkb200_<name>.so with one source recompiled under extra -D flags (selected
    at run time with LINKB200_LIB=<path>); every other object is shared with the main build."""
    build(verbose=verbose)
    obj = os.path.join(OBJ, f'{source[:-3]}_{name}.o')
    src = os.path.join(CSRC, source)
    cmd = [NVCC] + FLAGS + [f'-D{d}' for d in defines] + ['-c', src, '-o', obj]
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + '.o') for s in sources() if not s.endswith('/' + source)]
    lib = os.path.join(HERE, f'liblinkb200_{name}.so')
    subprocess.check_call([NVCC, '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', lib, obj] + objs)
    return lib


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose=True))
