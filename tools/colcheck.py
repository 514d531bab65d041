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
CASES = [] if (len(sys.argv) > 1 and sys.argv[1] == 'time') else [((32, 32, 32), (64, 64, 64), 20000), ((15, 20, 24), (30, 50, 48), 5000),
                  ((8, 8, 8), (16, 16, 16), 300), ((33, 31, 32), (66, 62, 64), 30000), ((5, 4, 5), (6, 9, 10), 200)]
for Nd, Kd, M in CASES:
    Jd = (6, 6, 6)
    om = rng.uniform(-numpy.pi, numpy.pi, (M, 3))
    om[:7] = [[numpy.pi, numpy.pi, numpy.pi], [-numpy.pi, -numpy.pi, -numpy.pi], [0, 0, 0], [numpy.pi, 0, -numpy.pi],
              [3.1, -3.1, 3.1], [-3.1, 3.1, -3.1], [0.001, -0.001, 3.14]]
    A = pynufft_b200.NUFFT(dev); A.plan(om, Nd, Kd, Jd)
    O = orc.NUFFT(); O.plan(om, Nd, Kd, Jd)
    x = (rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64)
    y = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(numpy.complex64)
    cperm, ctile, csub = A._col_perm()
    print(Nd, Kd, 'layout', A.layout(), ctile, csub, 'col perm ok', numpy.array_equal(cperm, orc.sort_permutation(O.p.k0, Kd, ctile, csub)))
    for gv in (0, 2, 1):
        if gv == 2 and min(Kd) < 16: continue
        A.set_variant(0 if gv != 1 else 1, gv)
        print('  gridding variant', gv, 'y2k', rel(A.y2k(y), O.y2k(y)), 'adj', rel(A.adjoint(y), O.adjoint(y)),
              'selfadj', rel(A.selfadjoint(x), O.adjoint(O.forward(x).astype(numpy.complex64))))
    A.release()

if len(sys.argv) > 1:
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
