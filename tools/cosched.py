"""Experiment: the shared-memory-bound tiled gather and the FP32-issue-bound column-sweep gather launched CONCURRENTLY on two
streams (each doing the full configuration-3 job on the modulated grid).  If the two kinds of CTA share the SMs, the pair
finishes sooner than the two kernels back to back; occupancy knobs: B200NUFFT_IT_PAD (bytes), B200NUFFT_COL_ICTAS."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import pynufft_b200

Nd, Kd, Jd, M = (128,) * 3, (256,) * 3, (6,) * 3, 2_000_000
om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))
A1 = pynufft_b200.NUFFT('cuda:0'); A1.plan(om, Nd, Kd, Jd)
A2 = pynufft_b200.NUFFT('cuda:0'); A2.plan(om, Nd, Kd, Jd); A2.set_variant(3, 0)
lib = A1._lib
P = ctypes.c_void_p
rng = numpy.random.default_rng(1)
x = A1.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
km = torch.empty((1,) + Kd, dtype=torch.complex64, device='cuda')
st0 = P(torch.cuda.current_stream().cuda_stream)
lib.b200nufft_pad_fft_modulated(A1._plan, P(x.data_ptr()), P(km.data_ptr()), 1, 1, 0, None, st0)
y1 = torch.empty((M,), dtype=torch.complex64, device='cuda')
y2 = torch.empty((M,), dtype=torch.complex64, device='cuda')
torch.cuda.synchronize()
sA = torch.cuda.Stream(priority=0)
sB = torch.cuda.Stream(priority=-1)          # higher priority
pa, pb = P(sA.cuda_stream), P(sB.cuda_stream)

def tiled(): lib.b200nufft_interp_modulated(A1._plan, P(km.data_ptr()), P(y1.data_ptr()), 1, pa)
def col(): lib.b200nufft_interp_modulated(A2._plan, P(km.data_ptr()), P(y2.data_ptr()), 1, pb)

def timed(fns, it=20, warm=3):
    cur = torch.cuda.current_stream()
    def once():
        e = torch.cuda.Event(); e.record(cur)
        sA.wait_event(e); sB.wait_event(e)
        for f in fns: f()
        ea, eb = torch.cuda.Event(), torch.cuda.Event()
        ea.record(sA); eb.record(sB)
        cur.wait_event(ea); cur.wait_event(eb)
    for _ in range(warm): once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cur)
    for _ in range(it): once()
    e1.record(cur); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3

out = {'IT_PAD': os.environ.get('B200NUFFT_IT_PAD'), 'COL_ICTAS': os.environ.get('B200NUFFT_COL_ICTAS')}
out['tiled_us'] = timed([tiled])
out['col_us'] = timed([col])
out['both_tiled_first_us'] = timed([tiled, col])
out['both_col_first_us'] = timed([col, tiled])
out['err'] = float(torch.linalg.norm(y1 - y2) / torch.linalg.norm(y1))
print(json.dumps(out))
