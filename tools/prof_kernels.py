"""Small driver for ncu captures: plan configuration 3 and run a few forward/adjoint pairs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import pynufft_b200

Nd, Kd, Jd, M = (128,) * 3, (256,) * 3, (6,) * 3, 2_000_000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (M, 3))
A = pynufft_b200.NUFFT('cuda:0')
A.plan(om, Nd, Kd, Jd)
if len(sys.argv) > 2:
    A.set_variant(int(sys.argv[2]), 0)      # interp variant (3 = column-sweep gather)
rng = numpy.random.default_rng(1)
x = A.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
for _ in range(n):
    y = A._forward_device(x)
    xa = A._adjoint_device(y)
torch.cuda.synchronize()
print('done', float(xa.abs().max()))
