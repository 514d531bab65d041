"""Times the fused pruned FFT passes of configuration 3 (forward / inverse, plain and modulated) and a pair."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import variant_env
import numpy, torch
import pynufft_b200
Nd, Kd, Jd, M = (128,) * 3, (256,) * 3, (6,) * 3, 2_000_000
om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))
A = pynufft_b200.NUFFT('cuda:0'); A.plan(om, Nd, Kd, Jd)
lib = A._lib; P = ctypes.c_void_p
st = lambda: P(torch.cuda.current_stream().cuda_stream)
rng = numpy.random.default_rng(1)
x = A.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
k = torch.empty((1,) + Kd, dtype=torch.complex64, device='cuda'); xo = torch.empty(Nd, dtype=torch.complex64, device='cuda')
def timed(fn, it=30, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / it * 1e3, 1)
out = {'variant': os.environ.get('B200NUFFT_VARIANT')}
out['pad_fft_us'] = timed(lambda: lib.b200nufft_pad_fft(A._plan, P(x.data_ptr()), P(k.data_ptr()), 1, 1, 0, None, st()))
out['pad_fft_mod_us'] = timed(lambda: lib.b200nufft_pad_fft_modulated(A._plan, P(x.data_ptr()), P(k.data_ptr()), 1, 1, 0, None, st()))
out['ifft_crop_mod_us'] = timed(lambda: lib.b200nufft_ifft_crop_modulated(A._plan, P(k.data_ptr()), P(xo.data_ptr()), 1, 1, 0, None, st()))
out['pair_us'] = timed(lambda: A._adjoint_device(A._forward_device(x)))
ref = A._adjoint_device(A._forward_device(x))
out['norm'] = float(torch.linalg.norm(ref))
print(json.dumps(out))
