"""Small 3-D / 2-D problems through every kernel family, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
import pynufft_b200
dev = torch.device('cuda', 0)
rng = numpy.random.default_rng(0)
for Nd, Kd, Jd, M, batch in (((15, 20, 24), (30, 50, 48), (6, 6, 6), 3000, None), ((8, 8, 8), (16, 16, 16), (6, 6, 6), 300, 2),
                             ((5, 4, 5), (6, 9, 10), (6, 6, 6), 100, None), ((32, 32), (64, 64), (6, 6), 2000, 8),
                             ((32, 32), (64, 64), (6, 6), 2000, None), ((12, 12, 12), (24, 24, 24), (5, 4, 3), 500, None),
                             ((20, 24), (40, 48), (6, 6), 1500, 40), ((16, 100), (32, 256), (6, 6), 1200, 10),
                             ((16, 16, 16), (32, 32, 32), (6, 6, 6), 1500, 3)):
    om = rng.uniform(-numpy.pi, numpy.pi, (M, len(Nd)))
    om[:2] = [[numpy.pi] * len(Nd), [-numpy.pi] * len(Nd)]
    A = pynufft_b200.NUFFT(dev); A.plan(om, Nd, Kd, Jd, batch=batch)
    shp = tuple(Nd) + ((batch,) if batch else ())
    x = (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).astype(numpy.complex64)
    for iv, gv in ((0, 0), (2, 2), (1, 1), (3, 0)):
        try:
            A.set_variant(iv, gv)
            y = A.forward(x); xa = A.adjoint(y); k = A.y2k(y); y2 = A.k2y(k)
        except RuntimeError as e:
            if 'unsupported' not in str(e) and 'variant 3' not in str(e): raise
    A.set_variant(0, 0)
    if batch is not None:
        s_img = x[..., 0].copy()
        ys = A.forward_one2many(s_img); A.adjoint_many2one(ys)
        A.solve(ys, 'cg', maxiter=2)
    if batch is None:
        A.solve(A.forward(x), 'cg', maxiter=2)
        A.solve(A.forward(x), 'bicgstab', maxiter=2)        # k_axpby, k_dotc on odd lengths
        A.solve(A.forward(x), 'lsmr', maxiter=2)
        xp = torch.from_numpy(x).pin_memory().numpy()
        A.forward(xp, out=torch.empty(M, dtype=torch.complex64).pin_memory().numpy(), slot=1); A.wait('forward', 1)
    A.release()
torch.cuda.synchronize()
print('sanitize_small done')
