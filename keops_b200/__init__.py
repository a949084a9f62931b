"""keops_b200 -- a B200-native (sm_100a) implementation of the KeOps symbolic map-reduce hot path.

``out_i = Reduce_j F(x_i, y_j, p)`` for Sum / LogSumExp / Min-ArgMin / ArgKMin reductions and their
autodiff-generated gradient reductions, behind the ``pykeops.torch`` API (``keops_b200.torch``).
There is no CPU fallback: the CUDA runtime ``libkb200.so`` must be built (``__graft_entry__.build()``).
"""
__version__ = "0.1.0"

from .runtime import capi as _capi  # noqa: F401  (fails loudly if the native runtime is missing)
from .binder import LoadKeOpsB200, get_binder, register_into_pykeops  # noqa: F401
from .runtime.plan import clean_cache  # noqa: F401
