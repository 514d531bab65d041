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

    # the method names linalg/solve_hsa.py uses (NUFFT_hsa API)
    def y2k(self, y):
        return self._y2k_device(y)

    def xx2k(self, xx):
        return self._xx2k_device(xx)

    def k2xx(self, k):
        return self._k2xx_device(k)

    def adjoint(self, y):
        return self._adjoint_device(y)

    def selfadjoint(self, x):
        return self._selfadjoint_device(x)


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
        # L1TVLAD only exists in the batched twin (linalg/solve_hsa.py:74-274); run single-coil on the same mock
        from reference.linalg import solve_hsa
        out['l1tvlad5'] = numpy.asarray(solve_hsa.L1TVLAD(dev, garr(y), 5, 2)).astype(c64)
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


def kindx_digest(kindx):
    """sha256 of the reference's uint32 index array (C order): pins the WHOLE array bit for bit in 64 characters."""
    import hashlib
    return hashlib.sha256(numpy.ascontiguousarray(kindx.astype(numpy.uint32)).tobytes()).hexdigest()


def run_config1_full():
    """BASELINE configs[0] at full size on the reference's own fixtures: om2D.npz (PROPELLER, M = 122 880) and the
    256 x 256 phantom, Kd 512^2, Jd 6^2 (tests/test_init.py of the reference).  The index array is pinned by its
    sha256; forward / adjoint are stored as complex64 (the tolerance is 1e-5)."""
    om = numpy.load('/root/reference/src/data/om2D.npz')['arr_0']
    x = numpy.load('/root/reference/src/data/phantom_256_256.npz')['arr_0'].astype(c64)
    Nd, Kd, Jd = (256, 256), (512, 512), (6, 6)
    A = pynufft.NUFFT()
    A.plan(om, Nd, Kd, Jd)
    st = pynufft.helper.plan(om, Nd, Kd, Jd, format='pELL')
    y = A.forward(x)
    y64 = y.astype(c64)
    out = dict(om=om, Nd=numpy.array(Nd), Kd=numpy.array(Kd), Jd=numpy.array(Jd), x=x.real.astype(numpy.float32),
               kindx_sha256=numpy.array(kindx_digest(st['pELL'].kindx)),
               forward=y64, adjoint=A.adjoint(y64).astype(c64), selfadjoint=A.selfadjoint(x).astype(c64))
    path = os.path.join(OUT, 'ref_c1_full.npz')
    numpy.savez_compressed(path, **out)
    print('ref_c1_full', 'M=%d' % om.shape[0], '%.1f KB' % (os.path.getsize(path) / 1024))


def run_config3_subset(nsel=20000, nvox=40000):
    """BASELINE configs[2] (3-D 128^3 / 256^3 / 6^3, om = default_rng(0).uniform(-pi, pi, (2M, 3))) through the
    reference on a random subset of the samples: forward rows are independent and the adjoint is linear in y, so the
    subset pins the full-size operator.  Stored: the subset, its forward values, its index digest, and the adjoint of a
    data vector supported on the subset at `nvox` random voxels (the full 128^3 image would be 16.8 MB)."""
    Nd, Kd, Jd = (128, 128, 128), (256, 256, 256), (6, 6, 6)
    om = numpy.random.default_rng(0).uniform(-numpy.pi, numpy.pi, (2_000_000, 3))
    rng = numpy.random.default_rng(1)
    x = (rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(c64)
    sel = numpy.sort(rng.choice(om.shape[0], nsel, replace=False))
    ysub = (rng.standard_normal(nsel) + 1j * rng.standard_normal(nsel)).astype(c64)
    vox = numpy.sort(rng.choice(int(numpy.prod(Nd)), nvox, replace=False))
    A = pynufft.NUFFT()
    A.plan(om[sel], Nd, Kd, Jd)
    st = pynufft.helper.plan(om[sel], Nd, Kd, Jd, format='pELL')
    adj = A.adjoint(ysub).astype(c64)
    out = dict(Nd=numpy.array(Nd), Kd=numpy.array(Kd), Jd=numpy.array(Jd), sel=sel.astype(numpy.int32), ysub=ysub,
               vox=vox.astype(numpy.int32), kindx_sha256=numpy.array(kindx_digest(st['pELL'].kindx)),
               forward=A.forward(x).astype(c64), adjoint_vox=adj.ravel()[vox],
               adjoint_norm=numpy.array(numpy.linalg.norm(adj)))
    path = os.path.join(OUT, 'ref_c3_subset.npz')
    numpy.savez_compressed(path, **out)
    print('ref_c3_subset', 'M=%d of 2000000' % nsel, '%.1f KB' % (os.path.getsize(path) / 1024))


def gaussian_coil_maps(Nd, B, seed=0):
    rng = numpy.random.default_rng(seed)
    grids = numpy.meshgrid(*[numpy.arange(n) for n in Nd], indexing='ij')
    maps = numpy.zeros(Nd + (B,), dtype=c64)
    for c in range(B):
        ang = 2 * numpy.pi * c / B
        ctr = [Nd[0] / 2 + 0.45 * Nd[0] * numpy.cos(ang), Nd[1] / 2 + 0.45 * Nd[1] * numpy.sin(ang)] + \
              [n / 2 for n in Nd[2:]]
        r2 = sum((g - c0) ** 2 for g, c0 in zip(grids, ctr))
        maps[..., c] = numpy.exp(-r2 / (2 * (0.6 * Nd[0]) ** 2)) * numpy.exp(1j * rng.uniform(0, 2 * numpy.pi))
    return maps


def run_multicoil(name, om, Nd, Kd, Jd, B, seed):
    """B > 1 through the reference: its NUFFT_cpu.forward_one2many / adjoint_many2one (linalg/nufft_cpu.py:177-203) are
    single-coil objects (plan() forces batch = 1, :107-108), so they are looped over the coils with the coil's map as
    `cpu_coil_profile`, and the images are averaged over the coils as x2s / cAggregate do
    (linalg/nufft_hsa.py:628-656, re_subroutine.py:441-497) -- BASELINE.md section 3."""
    from reference.linalg.nufft_cpu import NUFFT_cpu
    rng = numpy.random.default_rng(seed)
    sens = gaussian_coil_maps(Nd, B, seed)
    s = (rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(c64)
    M = om.shape[0]
    A = NUFFT_cpu()
    A.plan(om, Nd, Kd, Jd)
    y = numpy.zeros((M, B), dtype=c64)
    for c in range(B):
        A.volume['cpu_coil_profile'] = sens[..., c]
        y[:, c] = A.forward_one2many(s)
    y_in = (rng.standard_normal((M, B)) + 1j * rng.standard_normal((M, B))).astype(c64)
    imgs, imgs2 = [], []
    for c in range(B):
        A.volume['cpu_coil_profile'] = sens[..., c]
        imgs.append(A.adjoint_many2one(y_in[:, c]))
        imgs2.append(A.adjoint_many2one(y[:, c]))
    # batched forward / adjoint without coil maps (Nd+(B,) <-> (M, B)), coil by coil
    A.volume['cpu_coil_profile'] = numpy.ones(Nd)
    xb = (rng.standard_normal(Nd + (B,)) + 1j * rng.standard_normal(Nd + (B,))).astype(c64)
    fwd_b = numpy.stack([A.forward(numpy.ascontiguousarray(xb[..., c])) for c in range(B)], -1).astype(c64)
    adj_b = numpy.stack([A.adjoint(numpy.ascontiguousarray(y_in[:, c])) for c in range(B)], -1).astype(c64)
    out = dict(om=om, Nd=numpy.array(Nd), Kd=numpy.array(Kd), Jd=numpy.array(Jd), B=numpy.array(B), sens=sens, s=s,
               y_in=y_in, xb=xb, forward_one2many=y, adjoint_many2one=numpy.mean(imgs, axis=0).astype(c64),
               selfadjoint_one2many2one=numpy.mean(imgs2, axis=0).astype(c64), forward_batch=fwd_b, adjoint_batch=adj_b)
    path = os.path.join(OUT, name + '.npz')
    numpy.savez_compressed(path, **out)
    print(name, 'M=%d B=%d' % (M, B), '%.1f KB' % (os.path.getsize(path) / 1024))


def main():
    os.makedirs(OUT, exist_ok=True)
    if 'big' in sys.argv[1:] or 'all' in sys.argv[1:]:
        run_config1_full()
        run_config3_subset()
        rng = numpy.random.default_rng(77)
        th = numpy.arange(24) * numpy.pi * (numpy.sqrt(5.0) - 1.0) / 2.0        # golden-angle radial, 24 spokes x 128
        r = numpy.pi * (numpy.arange(128) - 64) / 64
        om = numpy.stack([numpy.outer(numpy.cos(th), r), numpy.outer(numpy.sin(th), r)], -1).reshape(-1, 2)
        run_multicoil('ref_2d_batch8', om, (64, 64), (128, 128), (6, 6), 8, 11)
        run_multicoil('ref_3d_batch3', rng.uniform(-numpy.pi, numpy.pi, (2500, 3)), (16, 16, 16), (32, 32, 32),
                      (6, 6, 6), 3, 12)
        if 'all' not in sys.argv[1:]:
            return
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
