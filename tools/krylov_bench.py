"""Time per iteration of the Krylov family on the device at configuration 3 (3-D 128^3 / 256^3 / J=6 / 2 M samples):
vectors stay in HBM, scalars only cross PCIe.  Prints one JSON line per solver, and round 1's arrangement (scipy on
host arrays around the device operator) for comparison."""
import json, sys, time
import numpy, torch
sys.path.insert(0, '.')
import pynufft_b200


def main():
    small = '--small' in sys.argv
    rng = numpy.random.default_rng(0)
    Nd, Kd, Jd, M = ((32,) * 3, (64,) * 3, (6,) * 3, 50000) if small else ((128,) * 3, (256,) * 3, (6,) * 3, 2_000_000)
    om = rng.uniform(-numpy.pi, numpy.pi, (M, 3))
    A = pynufft_b200.NUFFT('cuda:0')
    A.plan(om, Nd, Kd, Jd)
    y = A.to_device((rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(numpy.complex64))
    runs = [('cg', dict(maxiter=10), 10), ('bicgstab', dict(maxiter=5), 5), ('bicg', dict(maxiter=5), 5),
            ('gmres', dict(maxiter=1, restart=10), 10), ('lgmres', dict(maxiter=1, inner_m=10), 10),
            ('lsqr', dict(iter_lim=10), 10), ('lsmr', dict(maxiter=10), 10)]
    for name, kw, its in runs:
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = pynufft_b200.solve.solve(A, y, name, **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(json.dumps({'solver': name, 'args': kw, 'ms_total': round(dt * 1e3, 2), 'ms_per_iteration': round(dt * 1e3 / its, 3),
                          'finite': bool(torch.isfinite(x.abs()).all())}), flush=True)
    # round 1's arrangement for comparison: scipy runs the recurrence on host arrays, every operator application
    # ships a Kd grid each way
    import scipy.sparse.linalg as sla
    c64 = numpy.complex64
    K = int(numpy.prod(Kd))
    dev = lambda a, shape: A.to_device(numpy.ascontiguousarray(numpy.asarray(a).reshape(shape), dtype=c64))
    G = lambda k: A.to_host(A._y2k_device(A._k2y_device(dev(k, Kd)))).ravel()
    op = sla.LinearOperator((K, K), matvec=G, rmatvec=G, dtype=numpy.complex128)
    b = A.to_host(A._y2k_device(y)).ravel()
    t0 = time.perf_counter()
    sla.bicgstab(op, b, maxiter=5)
    dt = time.perf_counter() - t0
    print(json.dumps({'solver': 'bicgstab, scipy on host arrays + device operator (round 1)', 'ms_total': round(dt * 1e3, 2),
                      'ms_per_iteration': round(dt * 1e3 / 5, 3)}), flush=True)


main()
