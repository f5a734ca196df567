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


if __name__ == '__main__':
    build(force='--force' in sys.argv)
