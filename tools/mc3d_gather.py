"""3-D multi-coil gather on phase-modulated grids: column-sweep ("auto") against the tiled gather (variant 2), B coils."""
import json, sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import variant_env
import numpy, torch
import pynufft_b200

def timed(fn, it=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (2_000_000, 3))
out = {}
for B in [int(a) for a in sys.argv[1:]] or [4, 32]:
    A = pynufft_b200.NUFFT('cuda:0')
    A.plan(om, (128,) * 3, (256,) * 3, (6,) * 3, batch=B)
    y = torch.view_as_complex(torch.randn((2_000_000, B, 2), device='cuda:0'))
    km = A._y2k_device(y, modulated=True)
    r = {}
    for name, v in (('col', 0), ('tiled', 2)):
        A.set_variant(v, 0)
        ref = A._k2y_device(km, modulated=True)
        r[name + '_interp_us'] = round(timed(lambda: A._k2y_device(km, modulated=True)), 1)
        if name == 'col': y0 = ref
        else: r['col_vs_tiled'] = float(torch.linalg.norm(ref - y0) / torch.linalg.norm(ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        A._solve_device(y, 'cg', maxiter=2)
        torch.cuda.synchronize(); e0.record(); A._solve_device(y, 'cg', maxiter=8); e1.record(); torch.cuda.synchronize()
        t8 = e0.elapsed_time(e1)
        e0.record(); A._solve_device(y, 'cg', maxiter=2); e1.record(); torch.cuda.synchronize()
        r[name + '_cg_ms_per_iter'] = round((t8 - e0.elapsed_time(e1)) / 6, 3)
    r['gridding_us'] = round(timed(lambda: A._y2k_device(y, modulated=True)), 1)
    out['B%d' % B] = r
    A.release()
    del A, y, km, ref, y0
    torch.cuda.empty_cache()
print(json.dumps(out))
