"""Times the column-sweep gather / scatter of configuration 3 (modulated grid) and checks them against the tiled gather."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import variant_env
import numpy, torch
import pynufft_b200

Nd, Kd, Jd, M = (128,) * 3, (256,) * 3, (6,) * 3, 2_000_000
om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))
A = pynufft_b200.NUFFT('cuda:0'); A.plan(om, Nd, Kd, Jd)
lib = A._lib
P = ctypes.c_void_p
st = lambda: P(torch.cuda.current_stream().cuda_stream)
rng = numpy.random.default_rng(1)
x = A.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
km = torch.empty((1,) + Kd, dtype=torch.complex64, device='cuda')
lib.b200nufft_pad_fft_modulated(A._plan, P(x.data_ptr()), P(km.data_ptr()), 1, 1, 0, None, st())
y1 = torch.empty((M,), dtype=torch.complex64, device='cuda')
y2 = torch.empty((M,), dtype=torch.complex64, device='cuda')
g1 = torch.empty((1,) + Kd, dtype=torch.complex64, device='cuda')

def timed(fn, it=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

out = {'variant': os.environ.get('B200NUFFT_VARIANT'), 'ICTAS': os.environ.get('B200NUFFT_COL_ICTAS')}
lib.b200nufft_interp_modulated(A._plan, P(km.data_ptr()), P(y1.data_ptr()), 1, st())
out['tiled_us'] = timed(lambda: lib.b200nufft_interp_modulated(A._plan, P(km.data_ptr()), P(y1.data_ptr()), 1, st()))
A.set_variant(3, 0)
out['col_us'] = timed(lambda: lib.b200nufft_interp_modulated(A._plan, P(km.data_ptr()), P(y2.data_ptr()), 1, st()))
out['col_vs_tiled'] = float(torch.linalg.norm(y1 - y2) / torch.linalg.norm(y1))
out['gridding_mod_us'] = timed(lambda: lib.b200nufft_gridding_modulated(A._plan, P(y1.data_ptr()), P(g1.data_ptr()), 1, st()))
out['pair_us'] = timed(lambda: A._adjoint_device(A._forward_device(x)))
print(json.dumps(out))
