"""Timing of configurations 2 and 4 (2D 256^2, 32 coils, golden-angle radial)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import variant_env            # B200NUFFT_VARIANT: experiment builds (tools/build_variant.sh)
import numpy, torch, ctypes
import pynufft_b200
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_gpu_parity import golden_angle_radial, coil_maps

Nd, Kd, Jd, B = (256, 256), (512, 512), (6, 6), 32
om = golden_angle_radial()
A = pynufft_b200.NUFFT('cuda:0'); A.plan(om, Nd, Kd, Jd, batch=B); A.set_sense(coil_maps(Nd, B))
rng = numpy.random.default_rng(0)
s = A.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
def timed(fn, it=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
y = A.forward_one2many(s)
k = A._y2k_device(y)
out = {'M': om.shape[0], 'forward_one2many_us': timed(lambda: A.forward_one2many(s)),
       'adjoint_many2one_us': timed(lambda: A.adjoint_many2one(y)),
       'interp32_us': timed(lambda: A._k2y_device(k)), 'gridding32_us': timed(lambda: A._y2k_device(y))}
import time
A._solve_device(y, 'cg', maxiter=3)          # warm-up: scratch buffers, first launches
torch.cuda.synchronize(); t0 = time.perf_counter(); A._solve_device(y, 'cg', maxiter=100); torch.cuda.synchronize()
out['cg100_s'] = time.perf_counter() - t0
A1 = pynufft_b200.NUFFT('cuda:0'); A1.plan(om, Nd, Kd, Jd)
y1 = A1._forward_device(s)
A1._solve_device(y1, 'L1TVOLS', maxiter=3, rho=2)
torch.cuda.synchronize(); t0 = time.perf_counter(); A1._solve_device(y1, 'L1TVOLS', maxiter=100, rho=2); torch.cuda.synchronize()
out['l1tvols100_s'] = time.perf_counter() - t0
algo = 8 * B * 512 * 512 + 12 * om.shape[0] * 12 + 8 * B * om.shape[0]
out['interp_GBps'] = algo / out['interp32_us'] / 1e3
out['gridding_GBps'] = algo / out['gridding32_us'] / 1e3
print(json.dumps(out))
lib = A._lib; P = ctypes.c_void_p
st = lambda: P(torch.cuda.current_stream().cuda_stream)
x32 = A.s2x(s)
g32 = torch.empty(Kd + (B,), dtype=torch.complex64, device='cuda')
xo = torch.empty(Nd + (B,), dtype=torch.complex64, device='cuda')
print(json.dumps({
  'scale_pad_us': timed(lambda: lib.b200nufft_scale_pad(A._plan, P(x32.data_ptr()), P(g32.data_ptr()), B, 1, 0, None, st())),
  'fft_us': timed(lambda: lib.b200nufft_fft(A._plan, P(g32.data_ptr()), B, 0, st())),
  'crop_us': timed(lambda: lib.b200nufft_crop_scale(A._plan, P(g32.data_ptr()), P(xo.data_ptr()), B, 1, 0, None, st())),
  'memset_us': timed(lambda: g32.zero_())}))
import torch.fft
gcm = torch.empty((B,) + Kd, dtype=torch.complex64, device='cuda')
print(json.dumps({'torch_fft2_coil_major_us': timed(lambda: torch.fft.fft2(gcm)), 'torch_fft2_batch_inner_us': timed(lambda: torch.fft.fftn(g32, dim=(0, 1)))}))
from test_gpu_parity import propeller
om1 = propeller()
A1 = pynufft_b200.NUFFT('cuda:0'); A1.plan(om1, Nd, Kd, Jd)
x1 = s
y1 = A1._forward_device(x1)
k1 = A1._y2k_device(y1)
print(json.dumps({'config1_M': om1.shape[0], 'c1_pair_us': timed(lambda: A1._adjoint_device(A1._forward_device(x1))),
                  'c1_interp_us': timed(lambda: A1._k2y_device(k1)), 'c1_gridding_us': timed(lambda: A1._y2k_device(y1)),
                  'c1_forward_us': timed(lambda: A1._forward_device(x1))}))
