"""B200-native (sm_100a) differentiable render + photometric-consistency path.

Drop-in for the hot path of hassony2/handobjectconsist: the same ``autograd.Function`` /
``nn.Module`` surface as ``meshreg/neurender``, ``meshreg/warping`` and ``meshreg/optim``, backed by
hand-written CUDA kernels behind the C ABI of ``include/hoc_b200.h``.  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["neurender", "warping", "optim", "mano"]
