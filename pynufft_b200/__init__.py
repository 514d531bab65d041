"""
pynufft_b200 -- B200-native (sm_100a) implementation of PyNUFFT's device NUFFT hot path.

    from pynufft_b200 import NUFFT, helper
    A = NUFFT(helper.device_list()[0])          # or NUFFT('cuda:0')
    A.plan(om, Nd, Kd, Jd, batch=None)
    y = A.forward(x); x2 = A.adjoint(y); x3 = A.solve(y, 'cg', maxiter=30)

The compute path is libb200nufft.so (hand-written CUDA + cuFFT) behind the C ABI in
include/b200nufft.h; there is no CPU fallback.
"""
from . import helper  # noqa: F401


def __getattr__(name):
    # torch / the CUDA library are only needed once the operator is used
    if name == 'NUFFT':
        from .nufft import NUFFT
        return NUFFT
    if name in ('solve', 'krylov', 'dist'):
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)
