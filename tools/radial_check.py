"""3-D radial ("kooshball") trajectory at configuration-3 size: kernel times of the tiled and column-sweep variants
on strongly non-uniform sample density (k-space centre: ~1e5 samples per column; periphery: nearly empty columns)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import pynufft_b200

Nd, Kd, Jd = (128,) * 3, (256,) * 3, (6,) * 3
nspokes, nread = 7812, 256                       # 2.0 M samples
rng = numpy.random.default_rng(0)
v = rng.standard_normal((nspokes, 3)); v /= numpy.linalg.norm(v, axis=1, keepdims=True)
r = numpy.pi * (numpy.arange(nread) - nread / 2) / (nread / 2)
om = (v[:, None, :] * r[None, :, None]).reshape(-1, 3)
M = om.shape[0]
A = pynufft_b200.NUFFT('cuda:0'); A.plan(om, Nd, Kd, Jd)
lib = A._lib; P = ctypes.c_void_p
st = lambda: P(torch.cuda.current_stream().cuda_stream)
x = A.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
grid = torch.empty((1,) + Kd, dtype=torch.complex64, device='cuda')
yv = torch.empty((M,), dtype=torch.complex64, device='cuda')
def timed(fn, it=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
out = {'M': M}
res = {}
for name, (iv, gv) in (('auto', (0, 0)), ('tiled', (2, 2))):
    A.set_variant(iv, gv)
    lib.b200nufft_pad_fft(A._plan, P(x.data_ptr()), P(grid.data_ptr()), 1, 1, 0, None, st())
    out[name + '_interp_us'] = timed(lambda: lib.b200nufft_interp(A._plan, P(grid.data_ptr()), P(yv.data_ptr()), 1, st()))
    out[name + '_gridding_us'] = timed(lambda: lib.b200nufft_gridding_modulated(A._plan, P(yv.data_ptr()), P(grid.data_ptr()), 1, st()))
    out[name + '_pair_us'] = timed(lambda: A._adjoint_device(A._forward_device(x)))
    res[name] = A._adjoint_device(A._forward_device(x))
out['auto_vs_tiled_selfadjoint'] = float(torch.linalg.norm(res['auto'] - res['tiled']) / torch.linalg.norm(res['tiled']))
print(json.dumps(out))
