"""Quick GPU check of the column-sweep kernels against the oracle (small sizes) + timing at configuration 3."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy, torch
import pynufft_b200
from pynufft_b200 import _lib
from oracle import nufft_oracle as orc

def rel(a, b):
    a = numpy.asarray(a).ravel(); b = numpy.asarray(b).ravel()
    return numpy.linalg.norm(a - b) / numpy.linalg.norm(b)

dev = torch.device('cuda', 0)
rng = numpy.random.default_rng(3)
CASES = [] if (len(sys.argv) > 1 and sys.argv[1] == 'time') else [((32, 32, 32), (64, 64, 64), 20000), ((15, 20, 24), (30, 50, 48), 5000), ((8, 8, 8), (16, 16, 16), 300),
                  ((33, 31, 32), (66, 62, 64), 30000)]
for Nd, Kd, M in CASES:
    Jd = (6, 6, 6)
    om = rng.uniform(-numpy.pi, numpy.pi, (M, 3))
    om[:7] = [[numpy.pi, numpy.pi, numpy.pi], [-numpy.pi, -numpy.pi, -numpy.pi], [0, 0, 0], [numpy.pi, 0, -numpy.pi],
              [3.1, -3.1, 3.1], [-3.1, 3.1, -3.1], [0.001, -0.001, 3.14]]
    A = pynufft_b200.NUFFT(dev); A.plan(om, Nd, Kd, Jd)
    O = orc.NUFFT(); O.plan(om, Nd, Kd, Jd)
    x = (rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64)
    y = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(numpy.complex64)
    k = (rng.standard_normal(Kd) + 1j * rng.standard_normal(Kd)).astype(numpy.complex64)
    kindx, _, k0, perm, tile, sub = A._plan_arrays()
    print(Nd, Kd, 'layout', A.layout(), tile, sub, 'perm ok', numpy.array_equal(perm, orc.sort_permutation(O.p.k0, Kd, tile, sub)))
    print('  k2y', rel(A.k2y(k), O.k2y(k)), 'y2k', rel(A.y2k(y), O.y2k(y)),
          'fwd', rel(A.forward(x), O.forward(x)), 'adj', rel(A.adjoint(y), O.adjoint(y)),
          'selfadj', rel(A.selfadjoint(x), O.adjoint(O.forward(x).astype(numpy.complex64))))
    A.set_variant(1, 1)
    print('  generic: k2y', rel(A.k2y(k), O.k2y(k)), 'y2k', rel(A.y2k(y), O.y2k(y)), 'fwd', rel(A.forward(x), O.forward(x)))
    A.set_variant(0, 1)
    print('  mixed(0,1): fwd', rel(A.forward(x), O.forward(x)), 'adj', rel(A.adjoint(y), O.adjoint(y)))
    A.set_variant(1, 0)
    print('  mixed(1,0): fwd', rel(A.forward(x), O.forward(x)), 'adj', rel(A.adjoint(y), O.adjoint(y)))
    A.release()

if len(sys.argv) > 1:
    # timing at configuration 3
    lib = _lib.load()
    import ctypes
    P = ctypes.c_void_p
    ND, KD, JD, M = (128,) * 3, (256,) * 3, (6,) * 3, 2_000_000
    om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))
    for pref in ((1,) if sys.argv[1] == 'time' else (1, 0)):
        lib.b200nufft_set_layout_preference(pref)
        A = pynufft_b200.NUFFT(dev); t0 = time.perf_counter(); A.plan(om, ND, KD, JD); torch.cuda.synchronize()
        print('layout', A.layout(), 'plan s', time.perf_counter() - t0, 'bytes', lib.b200nufft_plan_bytes(A._plan))
        x = torch.from_numpy((rng.standard_normal(ND) + 1j * rng.standard_normal(ND)).astype(numpy.complex64)).to(dev)
        st = lambda: P(torch.cuda.current_stream().cuda_stream)
        grid = torch.empty((1,) + KD, dtype=torch.complex64, device=dev)
        yv = torch.empty((M,), dtype=torch.complex64, device=dev)
        xo = torch.empty(ND, dtype=torch.complex64, device=dev)
        def timed(fn, it=30, warm=3):
            for _ in range(warm): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(it): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / it * 1e3
        print('  pad_fft_native us', timed(lambda: lib.b200nufft_pad_fft_native(A._plan, P(x.data_ptr()), P(grid.data_ptr()), 1, 1, 0, None, st())))
        print('  interp_native us', timed(lambda: lib.b200nufft_interp_native(A._plan, P(grid.data_ptr()), P(yv.data_ptr()), 1, st())))
        print('  gridding_native us (incl memset+gather)', timed(lambda: lib.b200nufft_gridding_native(A._plan, P(yv.data_ptr()), P(grid.data_ptr()), 1, st())))
        print('  ifft_crop_native us', timed(lambda: lib.b200nufft_ifft_crop_native(A._plan, P(grid.data_ptr()), P(xo.data_ptr()), 1, 1, 0, None, st())))
        print('  pair us', timed(lambda: A._adjoint_device(A._forward_device(x))))
        ys = []
        xf = torch.from_numpy((numpy.random.default_rng(9).standard_normal(ND) + 0j).astype(numpy.complex64)).to(dev)
        if pref == 1: y1 = A._forward_device(xf).cpu().numpy(); x1 = A._adjoint_device(torch.from_numpy(y1).to(dev)).cpu().numpy()
        else:
            y0 = A._forward_device(xf).cpu().numpy(); x0 = A._adjoint_device(torch.from_numpy(y1).to(dev)).cpu().numpy()
            print('  col vs tile: fwd', rel(y1, y0), 'adj', rel(x1, x0))
        A.release()
