"""Build libnumga_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'ngb200.cu')
HDR = os.path.join(os.path.dirname(HERE), 'include', 'numga_b200.h')
LIB = os.path.join(HERE, 'libnumga_b200.so')


def nvcc_command():
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    return [os.path.join(cuda, 'bin', 'nvcc'), '-O3', '-std=c++17', '-shared', '-Xcompiler', '-fPIC',
            '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', SRC, '-o', LIB,
            '-lnvrtc', '-ldl', '-Xlinker', f'-rpath={cuda}/lib64']


def build(force=False, verbose=True):
    newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    cmd = nvcc_command()
    if verbose:
        print(' '.join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


FAST_SRC = os.path.join(HERE, 'csrc', 'fastcall.cpp')
FAST_NAME = '_ngb200_fast'


def fast_path():
    return os.path.join(HERE, FAST_NAME + '.so')


def build_fast(force=False, verbose=False):
    """The eager fast path (csrc/fastcall.cpp): a small torch C++ extension (host code only; it calls the CUDA
    library through function pointers), built in-tree with torch.utils.cpp_extension so that it travels to the GPU
    box next to libnumga_b200.so."""
    out = fast_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(FAST_SRC):
        return out
    import shutil
    from torch.utils.cpp_extension import load
    bdir = os.path.join(HERE, '_fastbuild')
    os.makedirs(bdir, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
    load(name=FAST_NAME, sources=[FAST_SRC], build_directory=bdir, extra_cflags=['-O2'], with_cuda=True, verbose=verbose,
         is_python_module=False)
    shutil.copyfile(os.path.join(bdir, FAST_NAME + '.so'), out)
    return out


if __name__ == '__main__':
    build(force='--force' in sys.argv)
    build_fast(force='--force' in sys.argv)
