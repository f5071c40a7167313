"""ctypes loader of oracle/unwrap_herraez.c — TEST INFRASTRUCTURE, NOT A PRODUCT PATH (see that file's header:
a restatement of the published Herraez et al. algorithm behind skimage.restoration.unwrap_phase, which the reference
calls at tensor_tools.py:531; PARITY UNPINNED because scikit-image is not available here)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        subprocess.run(['make', '-s', '-C', HERE], check=True)
        _LIB = ctypes.CDLL(os.path.join(HERE, 'build', 'libunwrap_oracle.so'))
        _LIB.unwrap2d_oracle.restype = ctypes.c_int
        _LIB.unwrap2d_oracle.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return _LIB


def unwrap_phase(ang, return_increments=False):
    """Drop-in for ``skimage.restoration.unwrap_phase(ang)`` on a 2-D float array (no mask, no wrap-around)."""
    a = np.ascontiguousarray(ang, dtype=np.float64)
    assert a.ndim == 2
    out = np.empty_like(a)
    inc = np.empty(a.shape, dtype=np.int32)
    rc = _lib().unwrap2d_oracle(a.ctypes.data, out.ctypes.data, inc.ctypes.data, a.shape[1], a.shape[0])
    if rc != 0:
        raise MemoryError('unwrap2d_oracle')
    return (out, inc) if return_increments else out
