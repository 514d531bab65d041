"""3-D multi-coil stage times (configuration 5's per-GPU share: 4 coils): interp, gridding, one CG iteration."""
import json, sys
import numpy, torch
sys.path.insert(0, '.')
import pynufft_b200

def timed(fn, it=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (2_000_000, 3))
out = {}
for B in (1, 2, 4):
    A = pynufft_b200.NUFFT('cuda:0')
    A.plan(om, (128,) * 3, (256,) * 3, (6,) * 3, batch=B if B > 1 else None)
    y = torch.view_as_complex(torch.randn((2_000_000, B, 2), device='cuda:0')) if B > 1 else torch.view_as_complex(torch.randn((2_000_000, 2), device='cuda:0'))
    k = A._y2k_device(y)
    out['B%d' % B] = {'gridding_us': round(timed(lambda: A._y2k_device(y, modulated=A._kspace_modulated())), 1),
                      'interp_us': round(timed(lambda: A._k2y_device(k)), 1)}
    if B == 4:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        A._solve_device(y, 'cg', maxiter=2)
        torch.cuda.synchronize(); e0.record(); A._solve_device(y, 'cg', maxiter=12); e1.record(); torch.cuda.synchronize()
        t12 = e0.elapsed_time(e1)
        e0.record(); A._solve_device(y, 'cg', maxiter=2); e1.record(); torch.cuda.synchronize()
        out['cg_ms_per_iter_4coils'] = round((t12 - e0.elapsed_time(e1)) / 10, 3)
    A.release()
print(json.dumps(out))
