"""Diagnostic (run by hand on a GPU box): how far float32 CG drifts from the exact-arithmetic iterates on configuration 4."""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy, torch
from oracle import nufft_oracle as orc
from test_gpu_parity import golden_angle_radial, coil_maps, rel, make
Nd, Kd, Jd, B = (256,256),(512,512),(6,6),32
om = golden_angle_radial(); sens = coil_maps(Nd,B)
O = orc.NUFFT(); O.plan(om,Nd,Kd,Jd,batch=B); O.set_sense(sens)
A = make(torch.device('cuda',0), om, Nd, Kd, Jd, batch=B); A.set_sense(sens)
rng = numpy.random.default_rng(4); s=(rng.standard_normal(Nd)+1j*rng.standard_normal(Nd)).astype(numpy.complex64)
y = O.forward_one2many(s).astype(numpy.complex64)
x64 = orc.solve_cg(O,y,10,dtype=numpy.complex128); x32 = orc.solve_cg(O,y,10)
xg = A.solve(y,'cg',maxiter=10)
A.set_variant(1,1); xgen = A.solve(y,'cg',maxiter=10)
print('gpu vs f64', rel(xg,x64), 'generic vs f64', rel(xgen,x64), 'oracle32 vs f64', rel(x32,x64), 'gpu vs oracle32', rel(xg,x32))
