"""Build the CUDA extension in-tree: ``simudo_b200/csrc/libsimudo_b200.so``.

Plain ``nvcc -shared`` for sm_100a (no JIT cache, the .so travels with the
repo snapshot to the GPU box).  ``python -m simudo_b200.csrc.build``.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libsimudo_b200.so')
SOURCES = ['pdd_kernels.cu', 'band_solver.cu']
DEPS = SOURCES + ["native.inl", 'pdd_device.cuh', 'pdd_tables.h', 'optics.inl', 'lcn.inl', 'points.inl', '../../include/simudo_b200.h']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '--shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', SO] + [os.path.join(HERE, s) for s in SOURCES]
    r = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed building libsimudo_b200.so')
    with open(os.path.join(HERE, 'ptxas_info.txt'), 'w') as f:
        f.write(r.stdout + r.stderr)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
