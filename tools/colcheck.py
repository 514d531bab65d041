"""Stage timings at configuration 3 for the default (tiled gather + column-sweep scatter) and the tiled-only variants.
Parity of the same kernels lives in tests/test_gpu_parity.py (test_gridding_kernels_agree_3d)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy, torch
import pynufft_b200
from pynufft_b200 import _lib

def rel(a, b):
    a = numpy.asarray(a).ravel(); b = numpy.asarray(b).ravel()
    return numpy.linalg.norm(a - b) / numpy.linalg.norm(b)

dev = torch.device('cuda', 0)
rng = numpy.random.default_rng(3)
if True:
    # timing at configuration 3
    lib = _lib.load()
    import ctypes
    P = ctypes.c_void_p
    ND, KD, JD, M = (128,) * 3, (256,) * 3, (6,) * 3, 2_000_000
    om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))
    A = pynufft_b200.NUFFT(dev); t0 = time.perf_counter(); A.plan(om, ND, KD, JD); torch.cuda.synchronize()
    print('layout', A.layout(), 'plan s', time.perf_counter() - t0, 'bytes', lib.b200nufft_plan_bytes(A._plan))
    x = torch.from_numpy((rng.standard_normal(ND) + 1j * rng.standard_normal(ND)).astype(numpy.complex64)).to(dev)
    st = lambda: P(torch.cuda.current_stream().cuda_stream)
    grid = torch.empty((1,) + KD, dtype=torch.complex64, device=dev)
    yv = torch.randn((M,), dtype=torch.complex64, device=dev)
    xo = torch.empty(ND, dtype=torch.complex64, device=dev)
    def timed(fn, it=30, warm=3):
        for _ in range(warm): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it * 1e3
    res = {}
    for gv in (0, 2):
        A.set_variant(0, gv)
        print(' gridding variant', gv)
        print('  pad_fft us', timed(lambda: lib.b200nufft_pad_fft(A._plan, P(x.data_ptr()), P(grid.data_ptr()), 1, 1, 0, None, st())))
        print('  interp us', timed(lambda: lib.b200nufft_interp(A._plan, P(grid.data_ptr()), P(yv.data_ptr()), 1, st())))
        print('  gridding (true grid; incl memset+gather[+demod]) us', timed(lambda: lib.b200nufft_gridding(A._plan, P(yv.data_ptr()), P(grid.data_ptr()), 1, st())))
        print('  gridding_modulated (incl memset+gather) us', timed(lambda: lib.b200nufft_gridding_modulated(A._plan, P(yv.data_ptr()), P(grid.data_ptr()), 1, st())))
        print('  ifft_crop_modulated us', timed(lambda: lib.b200nufft_ifft_crop_modulated(A._plan, P(grid.data_ptr()), P(xo.data_ptr()), 1, 1, 0, None, st())))
        print('  pair us', timed(lambda: A._adjoint_device(A._forward_device(x))))
        res[gv] = A._adjoint_device(yv).cpu().numpy()
    print('  col vs tile adjoint:', rel(res[0], res[2]))
    A.release()
