"""Loader of the eager fast path (csrc/fastcall.cpp, built in-tree by numga_b200/build.py `build_fast`).

`module` is None when the extension has not been built; eager calls then take the python launch path (same CUDA
kernels, more host time per call).  This is NOT a compute fallback: both paths launch the same generated kernels
through libnumga_b200.so."""
import ctypes
import importlib.util
import os

from numga_b200 import runtime as rt

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, '_ngb200_fast.so')
module = None


def _load():
    global module
    if os.environ.get('NGB200_NO_FASTPATH', '0') == '1' or not os.path.exists(_PATH):
        return
    import torch  # noqa: F401  (the extension links against libtorch, which must be loaded first)
    spec = importlib.util.spec_from_file_location('_ngb200_fast', _PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    addr = lambda f: ctypes.cast(f, ctypes.c_void_p).value
    mod.bind(addr(rt.lib.ngb200_plan_launch), addr(rt.lib.ngb200_last_error))
    module = mod


try:
    _load()
except Exception as e:          # an extension built against another torch: report once, keep the python path
    import warnings
    warnings.warn(f'numga_b200 fast path unavailable ({e}); using the python launch path')
    module = None
