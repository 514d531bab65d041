"""Small driver for ncu captures of the 2-D multi-coil kernels (configuration 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy, torch
import pynufft_b200
from test_gpu_parity import golden_angle_radial, coil_maps
Nd, Kd, Jd, B = (256, 256), (512, 512), (6, 6), 32
A = pynufft_b200.NUFFT('cuda:0'); A.plan(golden_angle_radial(), Nd, Kd, Jd, batch=B); A.set_sense(coil_maps(Nd, B))
rng = numpy.random.default_rng(0)
s = A.to_device((rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    y = A.forward_one2many(s)
    x = A.adjoint_many2one(y)
torch.cuda.synchronize()
print('done', float(x.abs().max()))
