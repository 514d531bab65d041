"""Quick kernel timing + self-consistency check on configuration 3 (tiled vs generic kernels)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import pynufft_b200

Nd, Kd, Jd, M = (128,) * 3, (256,) * 3, (6,) * 3, int(os.environ.get('KB_M', 2_000_000))
om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))
A = pynufft_b200.NUFFT('cuda:0')
A.plan(om, Nd, Kd, Jd)
lib = A._lib
P = ctypes.c_void_p
st = lambda: P(torch.cuda.current_stream().cuda_stream)
rng = numpy.random.default_rng(1)
x = A.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
k = A._xx2k_device(A._x2xx_device(x))
yv = torch.empty((M,), dtype=torch.complex64, device='cuda')
grid = torch.empty(Kd, dtype=torch.complex64, device='cuda')

def timed(fn, it=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

out = {}
A.set_variant(1, 1)
y_gen = A._k2y_device(k); g_gen = A._y2k_device(y_gen)
A.set_variant(0, 0)
y_til = A._k2y_device(k); g_til = A._y2k_device(y_gen)
out['interp_tiled_vs_generic'] = float(torch.linalg.norm(y_til - y_gen) / torch.linalg.norm(y_gen))
out['gridding_tiled_vs_generic'] = float(torch.linalg.norm(g_til - g_gen) / torch.linalg.norm(g_gen))
out['interp_us'] = timed(lambda: lib.b200nufft_interp(A._plan, P(k.data_ptr()), P(yv.data_ptr()), 1, st()))
out['gridding_us'] = timed(lambda: lib.b200nufft_gridding(A._plan, P(y_gen.data_ptr()), P(grid.data_ptr()), 1, st()))
out['memset_us'] = timed(lambda: grid.zero_())
km = torch.empty((1,) + Kd, dtype=torch.complex64, device='cuda')
xs = A._x2xx_device(x)
if A._kspace_modulated():
    lib.b200nufft_pad_fft_modulated(A._plan, P(x.data_ptr()), P(km.data_ptr()), 1, 1, 0, None, st())
    ym = torch.empty((M,), dtype=torch.complex64, device='cuda')
    lib.b200nufft_interp_modulated(A._plan, P(km.data_ptr()), P(ym.data_ptr()), 1, st())
    out['interp_mod_vs_generic'] = float(torch.linalg.norm(ym - y_gen) / torch.linalg.norm(y_gen))
    out['interp_mod_us'] = timed(lambda: lib.b200nufft_interp_modulated(A._plan, P(km.data_ptr()), P(yv.data_ptr()), 1, st()))
    out['gridding_mod_us'] = timed(lambda: lib.b200nufft_gridding_modulated(A._plan, P(y_gen.data_ptr()), P(grid.data_ptr()), 1, st()))
    out['pad_fft_mod_us'] = timed(lambda: lib.b200nufft_pad_fft_modulated(A._plan, P(x.data_ptr()), P(km.data_ptr()), 1, 1, 0, None, st()))
    A.set_variant(3, 0)                 # column-sweep gather on the modulated grid
    out['interp_col_us'] = timed(lambda: lib.b200nufft_interp_modulated(A._plan, P(km.data_ptr()), P(yv.data_ptr()), 1, st()))
    out['interp_col_vs_generic'] = float(torch.linalg.norm(yv - y_gen) / torch.linalg.norm(y_gen))
    out['pair_col_us'] = timed(lambda: A._adjoint_device(A._forward_device(x)))
    A.set_variant(2, 2)
    out['interp_tiled_us'] = timed(lambda: lib.b200nufft_interp(A._plan, P(k.data_ptr()), P(yv.data_ptr()), 1, st()))
    out['gridding_tiled_us'] = timed(lambda: lib.b200nufft_gridding(A._plan, P(y_gen.data_ptr()), P(grid.data_ptr()), 1, st()))
    A.set_variant(0, 0)
A.set_variant(1, 1)
out['interp_generic_us'] = timed(lambda: lib.b200nufft_interp(A._plan, P(k.data_ptr()), P(yv.data_ptr()), 1, st()), it=5, warm=1)
out['gridding_generic_us'] = timed(lambda: lib.b200nufft_gridding(A._plan, P(y_gen.data_ptr()), P(grid.data_ptr()), 1, st()), it=5, warm=1)
A.set_variant(0, 0)
out['scale_pad_us'] = timed(lambda: lib.b200nufft_scale_pad(A._plan, P(x.data_ptr()), P(grid.data_ptr()), 1, 1, 0, None, st()))
out['fft_us'] = timed(lambda: lib.b200nufft_fft(A._plan, P(grid.data_ptr()), 1, 0, st()))
xo = torch.empty(Nd, dtype=torch.complex64, device='cuda')
out['crop_us'] = timed(lambda: lib.b200nufft_crop_scale(A._plan, P(grid.data_ptr()), P(xo.data_ptr()), 1, 1, 0, None, st()))
out['pair_us'] = timed(lambda: A._adjoint_device(A._forward_device(x)))
algo = 8 * 256**3 + 12 * M * 18 + 8 * M
out['interp_GBps'] = algo / out['interp_us'] / 1e3
out['gridding_GBps'] = algo / (out['gridding_us'] - out['memset_us']) / 1e3
print(json.dumps(out))
k2 = torch.empty((1,) + Kd, dtype=torch.complex64, device='cuda')
print(json.dumps({'pad_fft_us': timed(lambda: lib.b200nufft_pad_fft(A._plan, P(x.data_ptr()), P(k2.data_ptr()), 1, 1, 0, None, st())),
                  'ifft_crop_us': timed(lambda: lib.b200nufft_ifft_crop(A._plan, P(k2.data_ptr()), P(xo.data_ptr()), 1, 1, 0, None, st()))}))
