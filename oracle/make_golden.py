"""
Generate tests/golden/*.npz by running the UNMODIFIED reference (CPU path) from
/root/reference.  Test infrastructure only; runs in the build container, never on
the GPU box (the fixtures travel, the reference does not).

    python oracle/make_golden.py

What is recorded per case (all produced by reference code, none by this repo):
  * helper.plan(format='pELL'): kindx, udata, meshindex, tensor_sn, alpha   (src/_helper/helper.py:620-802)
  * NUFFT() CPU operator: forward(x), adjoint(y), selfadjoint(x) and the six stages
    (nufft/_nufft_class_methods_cpu.py:168-362)
  * selfadjoint2 (Toeplitz-style approximation, _nufft_class_methods_cpu.py:127-146), solve(..., 'dc')
    (linalg/solve_cpu.py:165-225) and the scipy Krylov family solve(..., 'lsmr' | 'lsqr' | 'bicgstab' | 'gmres')
    (linalg/solve_cpu.py:226-288)
  * the reference's *device* solvers (linalg/solve_device.py: cg, L1TVOLS) executed on a
    numpy mock of the reikna thread/program objects, so the real solver control flow is
    what pins oracle.solve_cg / oracle.solve_l1tvols.
"""
import os
import sys
import types
import warnings

import numpy

warnings.filterwarnings('ignore')
sys.path.insert(0, '/root')
import reference as pynufft  # noqa: E402

sys.modules['pynufft'] = pynufft
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden')
c64 = numpy.complex64


# ----------------------------------------------------------------------------------------- #
# numpy mock of the reikna objects the reference's device solver touches
# ----------------------------------------------------------------------------------------- #
class GArr(numpy.ndarray):
    def get(self):
        return numpy.array(self)

    def fill(self, v):
        numpy.ndarray.fill(self, v)
        return self


def garr(a):
    return numpy.array(a, dtype=c64 if numpy.iscomplexobj(a) else a.dtype, order='C').view(GArr)


class MockThr:
    def to_device(self, a):
        return garr(a)

    def copy_array(self, a, dest=None):
        if dest is not None:
            dest[...] = a
            return dest
        return garr(numpy.array(a))

    def empty_like(self, a):
        return garr(numpy.zeros(a.shape, dtype=a.dtype))

    def array(self, shape, dtype):
        return garr(numpy.zeros(shape, dtype=dtype))

    def synchronize(self):
        pass


class MockPrg:
    """Element-wise kernels of src/re_subroutine.py, evaluated in float32 like the device."""

    def __init__(self, nufft):
        self.n = nufft

    def cMultiplyScalar(self, a, x, **kw):
        x[...] = (c64(a) * x).astype(c64)

    def cDiff(self, order, indata, outdata, **kw):
        o = numpy.asarray(order)
        outdata.flat[:] = (indata.ravel()[o] - indata.ravel()).astype(c64)

    def cHypot(self, x, y, **kw):
        r = numpy.hypot(numpy.abs(x).astype(numpy.float32), numpy.abs(y).astype(numpy.float32))
        x[...] = r.astype(numpy.float32).astype(c64)

    def cAnisoShrink(self, thr, indata, outdata, **kw):
        t = numpy.float32(numpy.real(thr))
        re, im = indata.real, indata.imag
        tr = (re > t) * (re - t) + (re < -t) * (re + t)
        ti = (im > t) * (im - t) + (im < -t) * (im + t)
        outdata[...] = (tr + 1j * ti).astype(c64)

    def cMultiplyVec(self, a, b, dest, **kw):
        dest[...] = (a * b).astype(c64)

    def cMultiplyConjVec(self, a, b, dest, **kw):
        dest[...] = (numpy.conj(a) * b).astype(c64)

    def cAddVec(self, a, b, dest, **kw):
        dest[...] = (a + b).astype(c64)

    def cTensorMultiply(self, batch, Tdims, Td, Td_el, invTd_el, tensor_sn, x, div, **kw):
        sn = self.n._cpu.sn.real.astype(numpy.float32)
        if div == 1:
            x[...] = (x / sn).astype(c64)
        else:
            x[...] = (x * sn).astype(c64)


class MockDeviceNUFFT:
    """The attributes/methods linalg/solve_device.py reads, served by the reference CPU object."""

    def __init__(self, cpu):
        self._cpu = cpu
        self.st = cpu.st
        self.thr = MockThr()
        self.prg = MockPrg(self)
        self.dtype = c64
        self.Nd, self.Kd = cpu.Nd, cpu.Kd
        self.Ndprod = int(numpy.prod(cpu.Nd))
        self.Kdprod = int(numpy.prod(cpu.Kd))
        self.batch = 1
        self.tSN = dict(Tdims=len(cpu.Nd), Td=None, Td_elements=None, invTd_elements=None, tensor_sn=None)
        self.volume = {}

    def _k2y_device(self, k):
        return garr(self._cpu._k2y_cpu(numpy.asarray(k)).astype(c64))

    def _y2k_device(self, y):
        return garr(self._cpu._y2k_cpu(numpy.asarray(y)).astype(c64))

    def _xx2k_device(self, xx):
        return garr(self._cpu._xx2k_cpu(numpy.asarray(xx)).astype(c64))

    def _k2xx_device(self, k):
        return garr(self._cpu._k2xx_cpu(numpy.asarray(k)).astype(c64))

    def _adjoint_device(self, y):
        return garr(self._cpu._adjoint_cpu(numpy.asarray(y)).astype(c64))

    def _selfadjoint_device(self, x):
        return garr(self._cpu._adjoint_cpu(self._cpu._forward_cpu(numpy.asarray(x))).astype(c64))


def install_fake_reikna():
    """solve_device.solve('cg') does `from reikna.algorithms import Reduce, Predicate, predicate_sum`."""
    reikna = types.ModuleType('reikna')
    alg = types.ModuleType('reikna.algorithms')

    class _Param:
        def __init__(self):
            self.output = garr(numpy.zeros((), dtype=c64))

    class Reduce:
        def __init__(self, arr, pred):
            self.parameter = _Param()

        def compile(self, thr):
            return self

        def __call__(self, out, arr):
            out[...] = c64(numpy.sum(numpy.asarray(arr), dtype=c64))

    alg.Reduce = Reduce
    alg.Predicate = object
    alg.predicate_sum = lambda dt: None
    reikna.algorithms = alg
    sys.modules['reikna'] = reikna
    sys.modules['reikna.algorithms'] = alg


def run_case(name, om, Nd, Kd, Jd, seed, solvers=False):
    rng = numpy.random.default_rng(seed)
    x = (rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(c64)
    M = om.shape[0]
    y_in = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(c64)
    A = pynufft.NUFFT()
    A.plan(om, Nd, Kd, Jd)
    st = pynufft.helper.plan(om, Nd, Kd, Jd, format='pELL')
    p = st['pELL']
    out = dict(
        om=om, Nd=numpy.array(Nd), Kd=numpy.array(Kd), Jd=numpy.array(Jd), x=x, y_in=y_in,
        kindx=p.kindx, udata=p.udata, meshindex=p.meshindex,
        tensor_sn=st['tSN'].tensor_sn,
        alpha=numpy.stack([numpy.asarray(a).reshape(-1) for a in st['alpha']]) if len(set(len(a) for a in st['alpha'])) == 1
        else numpy.concatenate([numpy.asarray(a).reshape(-1) for a in st['alpha']]),
        sn=A.sn,
        forward=A.forward(x).astype(numpy.complex128),
        adjoint=A.adjoint(y_in).astype(c64),
        selfadjoint=A.selfadjoint(x).astype(c64),
        xx=A.x2xx(x).astype(c64),
        k=A.xx2k(A.x2xx(x)).astype(c64),
        y2k=A.y2k(y_in).astype(c64),
        selfadjoint2=A._selfadjoint2_cpu(x).astype(c64),
    )
    if solvers:
        from reference.linalg import solve_device
        install_fake_reikna()
        y = A.forward(x).astype(c64)
        dev = MockDeviceNUFFT(A)
        out['solve_y'] = y
        out['cg10'] = numpy.asarray(solve_device.solve(dev, garr(y), 'cg', maxiter=10)).astype(c64)
        out['l1tvols5'] = numpy.asarray(solve_device.solve(dev, garr(y), 'L1TVOLS', maxiter=5, rho=2)).astype(c64)
        from reference.linalg import solve_cpu
        out['dc2'] = numpy.asarray(solve_cpu.solve(A, y, 'dc', 2)).astype(c64)       # linalg/solve_cpu.py:165-225
        # scipy Krylov family of the CPU solve (linalg/solve_cpu.py:226-288), few iterations, fixed counts
        out['lsmr6'] = numpy.asarray(solve_cpu.solve(A, y, 'lsmr', maxiter=6)).astype(c64)
        out['lsqr6'] = numpy.asarray(solve_cpu.solve(A, y, 'lsqr', iter_lim=6)).astype(c64)
        out['bicgstab3'] = numpy.asarray(solve_cpu.solve(A, y, 'bicgstab', maxiter=3)).astype(c64)
        out['gmres1'] = numpy.asarray(solve_cpu.solve(A, y, 'gmres', maxiter=1, restart=4)).astype(c64)
    path = os.path.join(OUT, name + '.npz')
    numpy.savez_compressed(path, **out)
    print(name, 'M=%d' % M, '%.1f KB' % (os.path.getsize(path) / 1024))


def main():
    os.makedirs(OUT, exist_ok=True)
    om2d = numpy.load('/root/reference/src/data/om2D.npz')['arr_0']
    rng = numpy.random.default_rng(1234)
    # 2D, PROPELLER trajectory subsample of the reference fixture, K/N == 2 branch
    run_case('ref_2d_64', om2d[::60], (64, 64), (128, 128), (6, 6), 1, solvers=True)
    # 2D, K/N != 2 branch, mixed J, samples on the +-pi edges and exactly on grid points
    om = rng.uniform(-numpy.pi, numpy.pi, (500, 2))
    om[0] = (-numpy.pi, numpy.pi)
    om[1] = (numpy.pi, -numpy.pi)
    om[2] = (0.0, 0.0)
    om[3] = (2 * numpy.pi / 40 * 7, -2 * numpy.pi / 36 * 5)
    run_case('ref_2d_odd', om, (24, 20), (40, 36), (4, 5), 2)
    # 3D
    om = rng.uniform(-numpy.pi, numpy.pi, (1500, 3))
    run_case('ref_3d_16', om, (16, 16, 16), (32, 32, 32), (6, 6, 6), 3, solvers=True)
    # 1D
    om = rng.uniform(-numpy.pi, numpy.pi, (300, 1))
    run_case('ref_1d_64', om, (64,), (128,), (6,), 4)
    # 3D with small N (L = ceil(N/3) branch of the alpha fit) and J = 4
    om = rng.uniform(-numpy.pi, numpy.pi, (400, 3))
    run_case('ref_3d_small', om, (12, 10, 8), (24, 20, 16), (4, 4, 4), 5)


if __name__ == '__main__':
    main()
