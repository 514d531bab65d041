import sys
sys.path.insert(0, '/root/repo')
import numpy, torch
import pynufft_b200
from oracle import nufft_oracle as orc
dev = torch.device('cuda', 0)
Nd, Kd, Jd = (8, 8, 8), (16, 16, 16), (6, 6, 6)
for om in ([[0.3, -1.2, 0.7]], [[0.3, -1.2, 0.7], [0.35, -1.2, 0.7]], [[0.3, -1.2, 0.7], [1.3, -1.2, 0.7]]):
    om = numpy.array(om)
    y = numpy.ones(len(om), numpy.complex64)
    A = pynufft_b200.NUFFT(dev); A.plan(om, Nd, Kd, Jd)
    O = orc.NUFFT(); O.plan(om, Nd, Kd, Jd)
    g = A.y2k(y); r = O.y2k(y)
    print('k0', O.p.k0, 'err', numpy.linalg.norm(g - r) / numpy.linalg.norm(r))
    nz_g = numpy.argwhere(numpy.abs(g) > 1e-9); nz_r = numpy.argwhere(numpy.abs(r) > 1e-9)
    print(' nonzero count', len(nz_g), len(nz_r), 'bbox g', nz_g.min(0), nz_g.max(0), 'bbox r', nz_r.min(0), nz_r.max(0))
    ratio = g[numpy.abs(r) > 1e-9] / r[numpy.abs(r) > 1e-9]
    for p in range(16):
        m = numpy.abs(r[p]) > 1e-9
        if m.any():
            rt = g[p][m] / r[p][m]
            print('  plane', p, 'ratio min/max abs', numpy.abs(rt).min(), numpy.abs(rt).max(), 'phase', numpy.angle(rt).min(), numpy.angle(rt).max())
    for b in range(16):
        m = numpy.abs(r[:, b]) > 1e-9
        if m.any():
            rt = g[:, b][m] / r[:, b][m]
            print('  row', b, 'ratio min/max abs', numpy.abs(rt).min(), numpy.abs(rt).max())
    for b in range(16):
        m = numpy.abs(r[:, :, b]) > 1e-9
        if m.any():
            rt = g[:, :, b][m] / r[:, :, b][m]
            print('  col', b, 'ratio min/max abs', numpy.abs(rt).min(), numpy.abs(rt).max())
