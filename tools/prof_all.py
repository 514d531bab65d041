"""Driver for the per-kernel ncu table (profiles/r2_ncu_summary.md): runs every kernel on the paths of configurations
1-5 a few times.  Run under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,\
smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,\
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none --csv --log-file gpurun_out/r2_allkernels.csv \
    python tools/prof_all.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy
import torch
import pynufft_b200
from test_gpu_parity import golden_angle_radial, coil_maps, propeller

dev = 'cuda:0'
rng = numpy.random.default_rng(0)
c64 = numpy.complex64


def rnd(shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(c64)


def mark(name):
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push(name)
    torch.cuda.nvtx.range_pop()
    print('==', name, flush=True)


# configuration 1: 2-D single coil (PROPELLER)
mark('config1')
A = pynufft_b200.NUFFT(dev)
A.plan(propeller(), (256, 256), (512, 512), (6, 6))
x = A.to_device(rnd((256, 256)))
for _ in range(2):
    y = A._forward_device(x)
    xa = A._adjoint_device(y)
A.release()

# configuration 2 / 4: 2-D, 32 coils, radial; CG + L1TVOLS iterations
mark('config2')
A = pynufft_b200.NUFFT(dev)
A.plan(golden_angle_radial(), (256, 256), (512, 512), (6, 6), batch=32)
A.set_sense(coil_maps((256, 256), 32))
s = A.to_device(rnd((256, 256)))
for _ in range(2):
    y = A.forward_one2many(s)
    sa = A.adjoint_many2one(y)
mark('config4')
A._solve_device(y, 'cg', maxiter=2)
A._solve_device(y, 'L1TVOLS', maxiter=2, rho=2)
A.release()

# configuration 3: 3-D single coil, 2 M samples
mark('config3')
om3 = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (2_000_000, 3))
A = pynufft_b200.NUFFT(dev)
A.plan(om3, (128,) * 3, (256,) * 3, (6,) * 3)
x = A.to_device(rnd((128,) * 3))
for _ in range(2):
    y = A._forward_device(x)
    xa = A._adjoint_device(y)
pynufft_b200.solve.solve(A, y, 'bicgstab', maxiter=1)     # Krylov family on device vectors: k_axpby, k_dotc
A.set_variant(3, 0)                      # column-sweep gather (not the default)
y = A._forward_device(x)
A.set_variant(0, 0)
A.release()

# configuration 5 (one GPU's share at 8 GPUs): 3-D, 4 coils, k-space CG
mark('config5')
A = pynufft_b200.NUFFT(dev)
A.plan(om3, (128,) * 3, (256,) * 3, (6,) * 3, batch=4)
y4 = torch.view_as_complex(torch.randn((2_000_000, 4, 2), device=dev))
A._solve_device(y4, 'cg', maxiter=2)
A.release()
torch.cuda.synchronize()
print('done')
