import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _build_library():
    """The C-ABI library must exist before any test module imports the package (CPU tests compile kernels with NVRTC,
    no launches).  The build script is loaded by path: importing `numga_b200` needs the library it produces."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('_numga_b200_build', os.path.join(ROOT, 'numga_b200', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build(verbose=False)


_build_library()
