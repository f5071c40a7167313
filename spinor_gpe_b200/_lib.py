"""Loader of the CUDA extension ``libsgpe.so`` (built in-tree by ``__graft_entry__.build()`` / ``make``).

There is deliberately NO fallback: if the library is missing, or no CUDA device is present, every entry
point of this package raises."""
import ctypes
import os

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
# (SGPE_LIB: a differently built copy of the same library - timeline stamps, build-time kernel experiments - for the dev
# tools; it is the CUDA library either way)
LIB_PATH = os.environ.get('SGPE_LIB') or os.path.join(_HERE, 'libsgpe.so')
_lib = None


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """The bound CDLL; raises ExtensionMissing (never falls back) if it cannot be loaded."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(
                f"{LIB_PATH} not found: build the sm_100a extension first "
                "(python -c 'import __graft_entry__ as g; g.build()' or make -C spinor_gpe_b200/csrc). "
                "spinor_gpe_b200 has no CPU or PyTorch fallback.")
        try:
            _lib = _capi.bind(ctypes.CDLL(LIB_PATH))
        except (OSError, AttributeError) as exc:
            raise ExtensionMissing(f"cannot load {LIB_PATH}: {exc}") from exc
    return _lib


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise ExtensionMissing("no CUDA device: spinor_gpe_b200 runs only on a GPU (sm_100a); "
                               "there is no CPU fallback.")
