"""
Device solvers on the NUFFT class: 'cg' and 'L1TVOLS' (+ 'dc'), the algorithms of
linalg/solve_device.py (cg :351-481, L1TVOLS :74-275; batched cg linalg/solve_hsa.py:551-682).

Same iterates as the reference, different execution: each CG iteration is
interp + gridding + 3 fused streaming kernels, and alpha/beta/rho live in a small device buffer
(float64 pairs) instead of being fetched with .get() three times per iteration
(solve_device.py:424,432,449).  When a process group is passed (coil-sharded operators,
pynufft_b200/dist.py) the two scalars are summed with one all-reduce each.
"""
import ctypes
import math

import numpy
import torch

from . import _lib

_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr())


def _stream(device=None):
    return _vp(torch.cuda.current_stream(device).cuda_stream)


class CudaVectorOps:
    """The fused CUDA vector kernels of csrc/solver.cu behind a tiny interface (so that the CG driver's
    control flow -- including where the all-reduces sit -- can be exercised on CPU by the tests)."""

    def __init__(self, lib):
        self.L = lib

    def scalars(self, device):
        return torch.zeros(6, dtype=torch.float64, device=device)

    def cg_init(self, b, Ax, r, p, rsold):
        _lib.check(self.L.b200nufft_cg_init(_ptr(b), _ptr(Ax), _ptr(r), _ptr(p), _ptr(rsold), r.numel(), _stream(r.device)))

    def dotc(self, a, b, out):
        out.zero_()
        _lib.check(self.L.b200nufft_dotc(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream(a.device)))

    def update_xr(self, x, r, p, Ap, rsold, pAp, rsnew):
        rsnew.zero_()
        _lib.check(self.L.b200nufft_cg_update_xr(_ptr(x), _ptr(r), _ptr(p), _ptr(Ap), _ptr(rsold), _ptr(pAp),
                                                 _ptr(rsnew), x.numel(), _stream(x.device)))

    def update_p(self, p, r, rsnew, rsold):
        _lib.check(self.L.b200nufft_cg_update_p(_ptr(p), _ptr(r), _ptr(rsnew), _ptr(rsold), p.numel(), _stream(p.device)))


def cg_kspace(G, b, maxiter, ops, allreduce=None):
    """
    The k-space CG of linalg/solve_device.py:351-461 on flat complex64 vectors:
    x0 = b, r = b - G x, exactly `maxiter` steps, one alpha/beta for everything in the vector
    (all coils, linalg/solve_hsa.py:555).  `G(v)` applies interp^H interp to a flat vector and returns a
    flat vector; `allreduce(t)` (optional) sums a 2-element float64 tensor over the coil shards.
    Returns x (flat).
    """
    x = b.clone()
    Ax = G(x)
    r = torch.empty_like(b)
    p = torch.empty_like(b)
    sc = ops.scalars(b.device)
    rsold, pAp, rsnew = sc[0:2], sc[2:4], sc[4:6]
    ops.cg_init(b, Ax, r, p, rsold)
    if allreduce is not None:
        allreduce(rsold)
    del Ax
    for _ in range(int(maxiter)):
        Ap = G(p)
        ops.dotc(p, Ap, pAp)
        if allreduce is not None:
            allreduce(pAp)
        ops.update_xr(x, r, p, Ap, rsold, pAp, rsnew)
        if allreduce is not None:
            allreduce(rsnew)
        ops.update_p(p, r, rsnew, rsold)
        rsold.copy_(rsnew)
        del Ap
    return x


def cg(nufft, gy, maxiter=30, group=None):
    """k-space CG on G = interp^H interp, x0 = b, exactly `maxiter` steps, then k2xx and /sn
    (linalg/solve_device.py:351-481).  `group`: process group of the coil shards (pynufft_b200.dist)."""
    L = nufft._lib
    allreduce = None
    if group is not None:
        import torch.distributed as dist
        allreduce = lambda t: dist.all_reduce(t, group=group)
    nb_in = int(gy.shape[1]) if gy.dim() == 2 else 1
    mod = nufft._kspace_modulated(nb_in)   # iterate on phase-modulated k-space vectors where the kernels allow it
    bview = nufft._y2k_device(gy, modulated=mod)
    batched = bview.dim() == nufft.ndims + 1
    nb = int(bview.shape[-1]) if batched else 1
    store = lambda view: nufft._grid_storage(view)[0]          # contiguous storage in the library layout (no copy)
    view = lambda flat: nufft._view_of(flat, nb, batched)

    def G(flat):
        return store(nufft._y2k_device(nufft._k2y_device(view(flat), modulated=mod), modulated=mod))

    xs = cg_kspace(G, store(bview), maxiter, CudaVectorOps(L), allreduce)
    # inverse FFT, crop, divide by sn  (solve_device.py:463-480)
    x2 = torch.empty(tuple(nufft.Nd) + ((nb,) if batched else ()), dtype=torch.complex64, device=nufft.device)
    crop = L.b200nufft_ifft_crop_modulated if mod else L.b200nufft_ifft_crop
    _lib.check(crop(nufft._plan, _ptr(xs), _ptr(x2), nb, 2, 0, None, nufft._stream()))
    return x2


def _sampling_density(nufft):
    """|y2k(1)|  (solve_device.py:25-36)"""
    ones = torch.ones((nufft.M,), dtype=torch.complex64, device=nufft.device)
    return nufft._y2k_device(ones).abs()


def _laplacian_kernel(nufft):
    """FFT of the (2d+1)-point Laplacian stencil on the Kd grid (src/_helper/helper.py:11-45)."""
    nd = nufft.ndims
    uker = numpy.zeros(nufft.Kd, dtype=numpy.complex64)
    uker[(0,) * nd] = -2.0 * nd
    for pp in range(nd):
        i1 = [0] * nd
        i1[pp] = 1
        uker[tuple(i1)] = 1
        i1[pp] = -1
        uker[tuple(i1)] = 1
    return numpy.fft.fftn(uker)


def L1TVLAD(nufft, gy, maxiter, rho):
    """L1-TV regularised least absolute deviation (linalg/solve_hsa.py:74-274): L1TVOLS with a second
    split variable on the data term: rhs from AHyk + df - bf, df = shrink(zf + bf, 1/mu), bf += zf - df."""
    return L1TVOLS(nufft, gy, maxiter, rho, lad=True)


def _aniso_shrink(t, thr):
    """cAnisoShrink (re_subroutine.py:973-993): soft threshold on the real and imaginary parts separately"""
    v = torch.view_as_real(t)
    return torch.view_as_complex(torch.where(v > thr, v - thr, torch.where(v < -thr, v + thr, torch.zeros_like(v))))


def L1TVOLS(nufft, gy, maxiter, rho, lad=False):
    """Split-Bregman total variation, device variant (solve_device.py:74-275).

    Multi-coil data (M, B) on a batch plan: the batched twin's closures (linalg/solve_hsa.py:275-476, AHA =
    nufft.selfadjoint, AH = nufft.adjoint at :282-287) between the ONE Nd image the solver iterates on (:291-318)
    and the B coils: AH = adjoint_many2one, AHA = selfadjoint_one2many2one (linalg/nufft_hsa.py:658-672, 747-769),
    i.e. TV-SENSE with the coil maps of set_sense().  uker and the Ku = rhs solve stay single-image."""
    multi = gy.dim() == 2
    if multi and int(gy.shape[1]) != nufft.batch:
        raise ValueError('y has %d coils, the plan has batch=%d' % (int(gy.shape[1]), nufft.batch))
    if not multi and nufft.batch != 1:
        raise ValueError('single-coil data on a batch=%d plan: pass y of shape (M, %d)' % (nufft.batch, nufft.batch))
    AH = nufft.adjoint_many2one if multi else nufft._adjoint_device
    AHA = nufft.selfadjoint_one2many2one if multi else nufft._selfadjoint_device
    L = nufft._lib
    st = nufft._stream
    mu = 1.0
    LMBD = rho * mu
    nd = nufft.ndims
    dev = nufft.device
    uker = (mu * _sampling_density(nufft).cpu().numpy() - LMBD * _laplacian_kernel(nufft)).astype(numpy.complex64)
    uker = torch.from_numpy(uker).to(dev)
    AHy = AH(gy)
    Nd = tuple(nufft.Nd)
    n = int(numpy.prod(Nd))
    xkp1 = torch.zeros(Nd, dtype=torch.complex64, device=dev)
    AHyk = torch.zeros(Nd, dtype=torch.complex64, device=dev)
    dd = torch.zeros((nd,) + Nd, dtype=torch.complex64, device=dev)
    bb = torch.zeros((nd,) + Nd, dtype=torch.complex64, device=dev)
    rhs = torch.empty(Nd, dtype=torch.complex64, device=dev)
    k = torch.empty(tuple(nufft.Kd), dtype=torch.complex64, device=dev)
    bf = torch.zeros(Nd, dtype=torch.complex64, device=dev) if lad else None
    df = torch.zeros(Nd, dtype=torch.complex64, device=dev) if lad else None
    for _ in range(int(maxiter)):
        src = ((AHyk + df) - bf).contiguous() if lad else AHyk           # solve_hsa.py:134-138
        _lib.check(L.b200nufft_tv_rhs(nufft._plan, _ptr(src), _ptr(dd), _ptr(bb), mu, LMBD, _ptr(rhs), st()))
        # xkp1 = k2xx(xx2k(rhs) / uker): zero-pad + FFT without sn scaling (:174-181)
        _lib.check(L.b200nufft_pad_fft(nufft._plan, _ptr(rhs), _ptr(k), 1, 0, 0, None, st()))
        _lib.check(L.b200nufft_cdiv(_ptr(k), _ptr(uker), k.numel(), st()))
        _lib.check(L.b200nufft_ifft_crop(nufft._plan, _ptr(k), _ptr(xkp1), 1, 0, 0, None, st()))
        zf = AHA(xkp1)
        _lib.check(L.b200nufft_tv_shrink(nufft._plan, _ptr(xkp1), _ptr(dd), _ptr(bb), LMBD, st()))
        if lad:                                                           # solve_hsa.py:229-260 (zf there = AHA x - AHy)
            zfm = zf - AHy
            df = _aniso_shrink(zfm + bf, 1.0 / mu)
            bf = bf + (zfm - df)
        _lib.check(L.b200nufft_tv_bregman(_ptr(AHyk), _ptr(zf), _ptr(AHy), n, st()))
    return xkp1


def density_compensation(nufft, gy, maxiter=1):
    """'dc': Pipe's iteration W <- W / (A A^H W), then adjoint(W*y)  (solve_device.py:277-310, solve_cpu.py:165-225)."""
    W = torch.ones((nufft.M,), dtype=torch.complex64, device=nufft.device)
    for _ in range(int(maxiter)):
        E = nufft._forward_device(nufft._adjoint_device(W))
        W = W / E
    return nufft._adjoint_device(W * gy)


KRYLOV = ('lsmr', 'lsqr', 'bicgstab', 'bicg', 'gmres', 'lgmres')


class CudaKrylovOps:
    """The vector interface of pynufft_b200/krylov.py on flat complex64 CUDA tensors: `b200nufft_axpby` for every
    update, `b200nufft_dotc` (double accumulation, result read back as two doubles) for inner products and norms."""

    def __init__(self, lib, device):
        self.L = lib
        self.acc = torch.zeros(2, dtype=torch.float64, device=device)

    def new(self, like):
        return torch.empty_like(like)

    def dot(self, a, b):
        self.acc.zero_()
        _lib.check(self.L.b200nufft_dotc(_ptr(a), _ptr(b), a.numel(), _ptr(self.acc), _stream(a.device)))
        re, im = self.acc.tolist()
        return complex(re, im)

    def nrm(self, a):
        return math.sqrt(max(self.dot(a, a).real, 0.0))

    def axpby(self, out, a, x, b=0.0, y=None):
        a, b = complex(a), complex(b)
        _lib.check(self.L.b200nufft_axpby(_ptr(out), a.real, a.imag, _ptr(x), b.real, b.imag,
                                          _ptr(y) if y is not None else None, out.numel(), _stream(out.device)))
        return out


def krylov(nufft, gy, solver, *args, **kwargs):
    """The scipy Krylov family of the reference's CPU solve (linalg/solve_cpu.py:226-288): 'lsmr' /
    'lsqr' on A = k2y (rmatvec y2k), the others on G = y2k . k2y with right-hand side y2k(y); then k2xx and DIVIDE by
    sn.  The recurrences are scipy's, restated on device vectors (pynufft_b200/krylov.py): the iterates never leave
    the GPU, only the scalars of the recurrences do.  Extra positional / keyword arguments are the scipy routine's
    (x0: flat device tensor or array; M: callable on flat device tensors).  Single coil."""
    from . import krylov as kr
    if nufft.batch not in (None, 1) or gy.dim() != 1:
        raise ValueError('the scipy Krylov solvers are single-coil (as the reference CPU object)')
    Kd, K, M = tuple(nufft.Kd), int(nufft.Kdprod), int(nufft.M)
    ops = CudaKrylovOps(nufft._lib, nufft.device)
    flat = lambda view: nufft._grid_storage(view)[0].reshape(-1)
    k2y = lambda k: nufft._k2y_device(k.view(Kd))
    y2k = lambda v: flat(nufft._y2k_device(v))
    if 'x0' in kwargs and kwargs['x0'] is not None and not torch.is_tensor(kwargs['x0']):
        kwargs['x0'] = nufft.to_device(numpy.ascontiguousarray(numpy.asarray(kwargs['x0']).reshape(Kd),
                                                               dtype=numpy.complex64)).reshape(-1)
    if solver in ('lsmr', 'lsqr'):
        like = torch.empty(K, dtype=torch.complex64, device=nufft.device)
        vec = {'lsmr': kr.lsmr, 'lsqr': kr.lsqr}[solver](ops, k2y, y2k, gy.contiguous(), like, *args, **kwargs)[0]
    else:
        G = lambda k: y2k(k2y(k))
        b = y2k(gy)
        if solver == 'bicg':
            vec = kr.bicg(ops, G, G, b, *args, **kwargs)[0]
        else:
            vec = {'bicgstab': kr.bicgstab, 'gmres': kr.gmres, 'lgmres': kr.lgmres}[solver](ops, G, b, *args, **kwargs)[0]
    x2 = torch.empty(tuple(nufft.Nd), dtype=torch.complex64, device=nufft.device)
    _lib.check(nufft._lib.b200nufft_ifft_crop(nufft._plan, _ptr(vec), _ptr(x2), 1, 2, 0, None, nufft._stream()))
    return x2


def solve(nufft, gy, solver=None, maxiter=30, *args, **kwargs):
    """solve(nufft, y, solver, maxiter, **kw) -- linalg/solve_device.py:312; plus the CPU solve's scipy family."""
    if solver in KRYLOV:
        if 'maxiter' not in kwargs and solver != 'lsqr' and maxiter != 30:
            kwargs['maxiter'] = maxiter
        return krylov(nufft, gy, solver, *args, **kwargs)
    if solver == 'cg':
        return cg(nufft, gy, maxiter=maxiter, **kwargs)
    if solver == 'L1TVOLS':
        return L1TVOLS(nufft, gy, maxiter=maxiter, *args, **kwargs)
    if solver == 'L1TVLAD':
        return L1TVLAD(nufft, gy, maxiter=maxiter, *args, **kwargs)
    if solver == 'dc':
        return density_compensation(nufft, gy, maxiter=maxiter)
    raise ValueError("solver must be 'cg', 'L1TVOLS', 'L1TVLAD', 'dc' or one of %s (got %r)" % (', '.join(KRYLOV), solver))
