"""spinor_gpe_b200 — B200-native split-step propagator for the quasi-2D pseudospin-1/2 GPE.

Drop-in for the TensorPropagator path of ultracoldYEG/spinor-gpe: ``PSpinor`` (problem set-up),
``TensorPropagator`` (time stepping in hand-written sm_100a kernels behind the C ABI of include/sgpe.h),
``PropResult`` and the ``tensor_tools`` helpers.  No CPU fallback: the CUDA extension must be built
(``__graft_entry__.build()``) and a GPU present.
"""
from . import tensor_tools                                  # noqa: F401
from .prop_result import PropResult                          # noqa: F401
from .tensor_propagator import TensorPropagator              # noqa: F401
from .pspinor import PSpinor                                 # noqa: F401
from .plan import Plan                                       # noqa: F401

__all__ = ['PSpinor', 'TensorPropagator', 'PropResult', 'Plan', 'tensor_tools']
