"""Drop-in counterpart of ``pykeops.torch``: ``Genred``, ``KernelSolve``, ``LazyTensor`` (``ComplexLazyTensor``), ``Vi/Vj/Pm``
and the ``generic_*`` helpers, backed by the sm_100a engine."""
from .genred import Genred, generic_sum, generic_logsumexp, generic_argkmin, generic_argmin  # noqa: F401
from .kernel_solve import KernelSolve  # noqa: F401
from .lazytensor import LazyTensor, ComplexLazyTensor, Vi, Vj, Pm  # noqa: F401

__all__ = ["Genred", "KernelSolve", "LazyTensor", "ComplexLazyTensor", "Vi", "Vj", "Pm", "generic_sum", "generic_logsumexp",
           "generic_argkmin", "generic_argmin"]
