import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


EMUL_DIR = os.path.join(ROOT, 'tests', 'emul')
EMUL_SO = os.path.join(EMUL_DIR, 'libsimudo_b200_emul.so')


def build_emulation_library():
    """g++ build of csrc/pdd_kernels.cu against tests/emul/cuda_emul.h."""
    src = os.path.join(ROOT, 'simudo_b200', 'csrc', 'pdd_kernels.cu')
    src2 = os.path.join(ROOT, 'simudo_b200', 'csrc', 'band_solver.cu')
    deps = [src, src2, os.path.join(ROOT, 'simudo_b200', 'csrc', 'pdd_device.cuh'),
            os.path.join(ROOT, 'simudo_b200', 'csrc', 'pdd_tables.h'),
            os.path.join(ROOT, 'simudo_b200', 'csrc', 'optics.inl'),
            os.path.join(ROOT, 'simudo_b200', 'csrc', 'lcn.inl'),
            os.path.join(ROOT, 'simudo_b200', 'csrc', 'native.inl'),
            os.path.join(ROOT, 'simudo_b200', 'csrc', 'points.inl'),
            os.path.join(ROOT, 'include', 'simudo_b200.h'),
            os.path.join(EMUL_DIR, 'cuda_emul.h')]
    if (not os.path.exists(EMUL_SO)
            or os.path.getmtime(EMUL_SO) < max(os.path.getmtime(d) for d in deps)):
        subprocess.check_call(
            ['g++', '-std=c++20', '-O2', '-fPIC', '-shared', '-pthread',
             '-DSB_EMUL_BUILD', '-I', EMUL_DIR, '-x', 'c++', src, src2, '-o', EMUL_SO])
    return EMUL_SO


@pytest.fixture(scope='session')
def emulated_lib():
    """Kernel source compiled for the CPU (test infrastructure only)."""
    from simudo_b200 import _lib
    so = build_emulation_library()
    _lib._use_library_for_tests(so)
    yield _lib
    _lib._reset_for_tests()
