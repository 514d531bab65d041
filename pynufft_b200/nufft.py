"""
class NUFFT -- the reference operator API (nufft/__init__.py:496-739 `class NUFFT`, batch API from
linalg/nufft_hsa.py `class NUFFT_hsa`) driven through libb200nufft.so.

Method names, argument meaning, return values and error behaviour follow the reference:
  plan(om, Nd, Kd, Jd, ft_axes=None, batch=None, radix=None) -> 0
  forward / adjoint / selfadjoint / solve and the stage methods x2xx, xx2k, k2y, y2k, k2xx, xx2x
      take and return numpy arrays (the reference's `_host` wrappers,
      nufft/_nufft_class_methods_cpu.py:366-445)
  _forward_device / _adjoint_device / _selfadjoint_device / _x2xx_device / _xx2k_device /
  _k2y_device / _y2k_device / _k2xx_device / _xx2x_device / _solve_device
      take and return device arrays (here: torch CUDA complex64 tensors instead of reikna arrays)
  to_device, to_host, release
  set_sense, reset_sense, s2x, x2s, forward_one2many, adjoint_many2one, selfadjoint_one2many2one

There is no CPU processor: `NUFFT()` without a CUDA device raises.

Concurrency contract (as the reference, whose kernels are queued on one reikna Thread): a plan owns its scratch grids,
staging buffers and cuFFT handles, so ONE plan is single-stream and single-thread -- issue its calls from one thread on
one CUDA stream at a time (the pipelined host API manages its own copy streams).  Use one NUFFT object per stream /
thread for concurrent work.  Every library call runs on the plan's device and restores the caller's current device.
"""
import ctypes

import numpy
import torch

from . import _lib, planmath

_vp = ctypes.c_void_p


def _ptr(t):
    return _vp(t.data_ptr())


def _stream(device=None):
    """cudaStream_t of torch's current stream ON `device` (the plan's device, which need not be the current one)"""
    return _vp(torch.cuda.current_stream(device).cuda_stream)


def _as_device(device_indx):
    """Accept torch.device, 'cuda:N', int, or a reference-style helper.device_list() tuple."""
    if device_indx is None:
        raise RuntimeError('pynufft_b200.NUFFT has no CPU processor: pass a CUDA device, e.g. NUFFT("cuda:0")')
    if isinstance(device_indx, torch.device):
        dev = device_indx
    elif isinstance(device_indx, int):
        dev = torch.device('cuda', device_indx)
    elif isinstance(device_indx, str):
        dev = torch.device(device_indx)
    elif isinstance(device_indx, (tuple, list)) and len(device_indx) >= 3:
        dev = torch.device('cuda', int(device_indx[2]))      # ('cuda', api_n, dev_num, ...) helper.py:1195
    else:
        raise TypeError('device must be a torch.device, "cuda:N" or an int')
    if dev.type != 'cuda':
        raise RuntimeError('pynufft_b200 runs on CUDA devices only (got %s)' % dev)
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    return dev


class NUFFT:
    def __init__(self, device_indx=None, legacy=None):
        if legacy:
            raise NotImplementedError('the legacy CSR device format is out of scope (SURVEY.md 8f)')
        self.device = _as_device(device_indx)
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError('no CUDA device available; pynufft_b200 has no CPU fallback')
        self.dtype = numpy.complex64
        self.processor = 'hsa'
        self.verbosity = 0
        self._plan = None
        self._inflight = {}
        self.Nd = self.Kd = self.Jd = ()
        self.ndims = 0
        self.ft_axes = ()
        self.batch = None

    def _stream(self):
        return _stream(self.device)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    # ------------------------------------------------------------------ plan / release
    def plan(self, om, Nd, Kd, Jd, ft_axes=None, batch=None, radix=None):
        # argument errors exactly as helper.plan (src/_helper/helper.py:644-654)
        if type(Nd) != tuple:
            raise TypeError('Nd must be tuple, e.g. (256, 256)')
        if type(Kd) != tuple:
            raise TypeError('Kd must be tuple, e.g. (512, 512)')
        if type(Jd) != tuple:
            raise TypeError('Jd must be tuple, e.g. (6, 6)')
        if (len(Nd) != len(Kd)) or (len(Nd) != len(Jd)):
            raise KeyError('Nd, Kd, Jd must be in the same length, e.g. Nd=(256,256),Kd=(512,512),Jd=(6,6)')
        nd = len(Nd)
        if ft_axes is not None and tuple(ft_axes) != tuple(range(nd)):
            raise NotImplementedError('partial ft_axes is unsupported (broken upstream, SURVEY.md 8c)')
        # radix (helper.py:497-556, doc/source/manu/variable_radix.rst) only regroups the reference's pELL storage
        # (products of `radix` dimensions pre-multiplied to save in-kernel multiplies); results are identical.  The
        # plan here always keeps per-dimension REAL factors, so any valid radix is accepted and changes nothing.
        if radix is not None and not (isinstance(radix, (int, numpy.integer)) and 1 <= int(radix) <= max(nd, 1)):
            raise ValueError('radix must be an integer in 1..%d' % max(nd, 1))
        if nd > _lib.MAX_DIM:
            raise NotImplementedError('ndim <= 3 supported')
        self.release()
        om = numpy.ascontiguousarray(numpy.asarray(om, dtype=numpy.float64))   # float32 om is up-cast (documented)
        if om.ndim == 1:
            om = om.reshape(-1, 1)
        if om.ndim != 2 or om.shape[1] != nd:
            raise ValueError('om must have shape (M, %d)' % nd)
        M = om.shape[0]

        alpha = numpy.zeros((nd, _lib.MAX_L), dtype=numpy.float64)
        alpha_len = numpy.zeros(nd, dtype=numpy.int32)
        Tm = numpy.zeros((nd, _lib.MAX_J * _lib.MAX_J), dtype=numpy.float64)
        sn_parts, self.st = [], {}
        self.st['alpha'], self.st['beta'] = [], []
        for d in range(nd):
            a, b = planmath.kb_fit_alpha(Nd[d], Jd[d], Kd[d])
            alpha[d, :len(a)] = a
            alpha_len[d] = len(a)
            Tm[d, :Jd[d] * Jd[d]] = planmath.interp_T(Nd[d], Jd[d], Kd[d], a, b).reshape(-1)
            sn_parts.append(numpy.real(planmath.scaling_vector(Nd[d], Kd[d], a, b)))
            self.st['alpha'].append(a)
            self.st['beta'].append(b)
        tensor_sn = numpy.ascontiguousarray(numpy.concatenate(sn_parts).astype(numpy.float32))

        Nd_a = numpy.asarray(Nd, dtype=numpy.int32)
        Kd_a = numpy.asarray(Kd, dtype=numpy.int32)
        Jd_a = numpy.asarray(Jd, dtype=numpy.int32)
        handle = _vp()
        with torch.cuda.device(self.device):
            rc = self._lib.b200nufft_plan_create(
                ctypes.byref(handle), self.device.index, nd, Nd_a.ctypes.data, Kd_a.ctypes.data, Jd_a.ctypes.data,
                M, om.ctypes.data, 1 if batch is None else int(batch), alpha.ctypes.data, alpha_len.ctypes.data,
                Tm.ctypes.data, tensor_sn.ctypes.data, self._stream())
        _lib.check(rc)
        self._plan = handle

        self.ndims = nd
        self.ft_axes = tuple(range(nd))
        self.Nd, self.Kd, self.Jd = Nd, Kd, Jd
        self.st.update(dict(Nd=Nd, Kd=Kd, Jd=Jd, M=numpy.int32(M), om=om, tol=0, tensor_sn=tensor_sn))
        self.parallel_flag = 0 if batch is None else 1
        self.batch = 1 if batch is None else int(batch)
        self.M = int(M)
        self.Ndprod = int(numpy.prod(Nd))
        self.Kdprod = int(numpy.prod(Kd))
        self.Jdprod = int(numpy.prod(Jd))
        if self.parallel_flag == 0:
            self.multi_Nd, self.multi_Kd, self.multi_M = Nd, Kd, (M,)
        else:
            self.multi_Nd, self.multi_Kd, self.multi_M = Nd + (self.batch,), Kd + (self.batch,), (M, self.batch)
        self._sense = None
        self._W = None
        return 0

    def release(self):
        if getattr(self, '_plan', None) is not None:
            self._lib.b200nufft_plan_destroy(self._plan)      # synchronises the pipelined host copies
            self._plan = None
        self._inflight = {}
        self._sense = None

    # ------------------------------------------------------------------ helpers
    def _require_plan(self):
        if self._plan is None:
            raise RuntimeError('plan() has not been called')

    def _nb_of(self, t, base_ndim, what):
        """Coil count of a device array whose single-coil rank is base_ndim (batch axis is last)."""
        if t.dim() == base_ndim:
            return 1
        if t.dim() == base_ndim + 1:
            return int(t.shape[-1])
        raise ValueError('%s has wrong rank %d' % (what, t.dim()))

    def _check_dev(self, t, shape_prefix, what):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex64):
            raise TypeError('%s must be a CUDA complex64 torch tensor (use to_device())' % what)
        if tuple(t.shape[:len(shape_prefix)]) != tuple(shape_prefix):
            raise ValueError('%s has shape %s, expected %s(+batch)' % (what, tuple(t.shape), tuple(shape_prefix)))
        return t.contiguous()

    def _bi(self, nb):
        """True when the library takes the grids of an nb-coil call batch-innermost (Kd + (nb,) contiguous, the
        reference's own layout; 2-D multi-coil kernels, csrc/sweep2d.cu) instead of coil-major ((nb,) + Kd)."""
        return int(self._lib.b200nufft_grid_layout(self._plan, int(nb))) == 1

    def _new_grid(self, nb, batched):
        """(Kd(+ (B,)) view, contiguous storage in the library's layout) of a fresh grid."""
        if batched and self._bi(nb):
            store = torch.empty(tuple(self.Kd) + (nb,), dtype=torch.complex64, device=self.device)
            return store, store
        store = torch.empty((nb,) + tuple(self.Kd), dtype=torch.complex64, device=self.device)
        if not batched:
            return store[0], store
        return store.permute(*range(1, self.ndims + 1), 0), store

    def _grid_storage(self, k, what='k'):
        """Return (contiguous storage tensor in the library's layout, nb, batched) for a user grid tensor."""
        if not (isinstance(k, torch.Tensor) and k.is_cuda and k.dtype == torch.complex64):
            raise TypeError('%s must be a CUDA complex64 torch tensor' % what)
        if tuple(k.shape[:self.ndims]) != tuple(self.Kd):
            raise ValueError('%s has shape %s, expected %s(+batch)' % (what, tuple(k.shape), tuple(self.Kd)))
        nb = self._nb_of(k, self.ndims, what)
        if k.dim() == self.ndims:
            return k.contiguous(), 1, False
        if self._bi(nb):
            return k.contiguous(), nb, True      # batch-innermost: the user's Kd + (B,) array as it is
        cm = k.permute(self.ndims, *range(self.ndims))
        return cm.contiguous(), nb, True         # no copy when k is already a coil-major view

    def _view_of(self, store, nb, batched):
        """Kd(+B) view of a storage tensor produced by _new_grid / _grid_storage."""
        if not batched or self._bi(nb):
            return store
        return store.permute(*range(1, self.ndims + 1), 0)

    # ------------------------------------------------------------------ host <-> device
    def to_device(self, x, shape=None):
        return torch.from_numpy(numpy.ascontiguousarray(numpy.asarray(x).astype(self.dtype))).to(self.device)

    def to_host(self, data):
        return data.detach().cpu().numpy()

    # ------------------------------------------------------------------ device stages
    def _x2xx_device(self, x):
        self._require_plan()
        x = self._check_dev(x, self.Nd, 'x')
        out = torch.empty_like(x)
        _lib.check(self._lib.b200nufft_x2xx(self._plan, _ptr(x), _ptr(out), self._nb_of(x, self.ndims, 'x'), 0,
                                            self._stream()))
        return out

    def _xx2x_device(self, xx):
        return self._x2xx_device(xx)

    def _xx2k_device(self, xx):
        self._require_plan()
        xx = self._check_dev(xx, self.Nd, 'xx')
        nb = self._nb_of(xx, self.ndims, 'xx')
        view, store = self._new_grid(nb, xx.dim() == self.ndims + 1)
        _lib.check(self._lib.b200nufft_pad_fft(self._plan, _ptr(xx), _ptr(store), nb, 0, 0, None, self._stream()))
        return view

    def _k2y_device(self, k, modulated=False):
        """modulated: k is a phase-modulated k-space vector (only when _kspace_modulated())."""
        self._require_plan()
        store, nb, batched = self._grid_storage(k)
        y = torch.empty((self.M, nb) if batched else (self.M,), dtype=torch.complex64, device=self.device)
        fn = self._lib.b200nufft_interp_modulated if modulated else self._lib.b200nufft_interp
        _lib.check(fn(self._plan, _ptr(store), _ptr(y), nb, self._stream()))
        return y

    def _y2k_device(self, y, modulated=False):
        """modulated: leave the grid phase-modulated (only when _kspace_modulated())."""
        self._require_plan()
        y = self._check_dev(y, (self.M,), 'y')
        nb = self._nb_of(y, 1, 'y')
        view, store = self._new_grid(nb, y.dim() == 2)
        fn = self._lib.b200nufft_gridding_modulated if modulated else self._lib.b200nufft_gridding
        _lib.check(fn(self._plan, _ptr(y), _ptr(store), nb, self._stream()))
        return view

    def _kspace_modulated(self, nb=1):
        """True when gridding, interp and the inverse FFT passes can all work on the phase-modulated grid of the sweep
        kernels (csrc/col3d.cu; csrc/sweep2d.cu for 2-D calls with an even number >= 8 of coils): k-space solvers then
        iterate on modulated vectors (G' = D G D^H with D diagonal and unitary has the same CG scalars) and skip the
        modulation work per G."""
        self._require_plan()
        return int(self._lib.b200nufft_kspace_modulated_nb(self._plan, int(nb))) == 1

    def _k2xx_device(self, k):
        """Inverse FFT (in place on k when k is a coil-major view, like the reference's in-place FFT) + crop."""
        self._require_plan()
        store, nb, batched = self._grid_storage(k)
        xx = torch.empty(tuple(self.Nd) + ((nb,) if batched else ()), dtype=torch.complex64, device=self.device)
        _lib.check(self._lib.b200nufft_fft(self._plan, _ptr(store), nb, 1, self._stream()))
        _lib.check(self._lib.b200nufft_crop_scale(self._plan, _ptr(store), _ptr(xx), nb, 0, 0, None, self._stream()))
        return xx

    def _forward_device(self, gx):
        self._require_plan()
        gx = self._check_dev(gx, self.Nd, 'x')
        nb = self._nb_of(gx, self.ndims, 'x')
        y = torch.empty((self.M, nb) if gx.dim() == self.ndims + 1 else (self.M,), dtype=torch.complex64,
                        device=self.device)
        _lib.check(self._lib.b200nufft_forward(self._plan, _ptr(gx), _ptr(y), nb, self._stream()))
        return y

    def _adjoint_device(self, gy):
        self._require_plan()
        gy = self._check_dev(gy, (self.M,), 'y')
        nb = self._nb_of(gy, 1, 'y')
        x = torch.empty(tuple(self.Nd) + ((nb,) if gy.dim() == 2 else ()), dtype=torch.complex64, device=self.device)
        _lib.check(self._lib.b200nufft_adjoint(self._plan, _ptr(gy), _ptr(x), nb, self._stream()))
        return x

    def _selfadjoint_device(self, gx):
        return self._adjoint_device(self._forward_device(gx))

    def _selfadjoint2_device(self, gx):
        """Toeplitz-style approximation of A^H A without interpolation: k2xx(W * xx2k(x)),
        W = |xx2k(adjoint(1_M))| computed on first use (reference: _precompute_sp_cpu / _selfadjoint2_cpu,
        nufft/_nufft_class_methods_cpu.py:127-146, 216-222).  Single coil."""
        self._require_plan()
        gx = self._check_dev(gx, self.Nd, 'x')
        if gx.dim() != self.ndims:
            raise ValueError('selfadjoint2 is single-coil')
        if getattr(self, '_W', None) is None:
            ones = torch.ones((self.M,), dtype=torch.complex64, device=self.device)
            self._W = self._xx2k_device(self._adjoint_device(ones)).abs().to(torch.complex64).contiguous()
        k = self._xx2k_device(gx).contiguous()
        _lib.check(self._lib.b200nufft_cmul(_ptr(k), _ptr(self._W), k.numel(), self._stream()))
        x2 = torch.empty(tuple(self.Nd), dtype=torch.complex64, device=self.device)
        _lib.check(self._lib.b200nufft_ifft_crop(self._plan, _ptr(k), _ptr(x2), 1, 0, 0, None, self._stream()))
        return x2

    def _solve_device(self, gy, solver=None, *args, **kwargs):
        from .solve import solve
        return solve(self, gy, solver, *args, **kwargs)

    # ------------------------------------------------------------------ multi-coil (NUFFT_hsa API)
    def set_sense(self, coil_profile):
        self._require_plan()
        shape = tuple(coil_profile.shape)
        if shape != tuple(self.Nd) + (self.batch,):
            print('The shape of coil_profile is ', shape)
            print('But it should be', tuple(self.Nd) + (self.batch,))
            raise ValueError
        if isinstance(coil_profile, torch.Tensor):
            self._sense = coil_profile.to(self.device, torch.complex64).contiguous()
        else:
            self._sense = self.to_device(coil_profile)

    def reset_sense(self):
        self._sense = None        # == all ones (linalg/nufft_hsa.py:333-335)

    def s2x(self, s):
        self._require_plan()
        s = self._check_dev(s, self.Nd, 's')
        x = s.unsqueeze(-1).expand(*s.shape, self.batch)
        if self._sense is not None:
            return x * self._sense
        return x.contiguous()

    def x2s(self, x):
        self._require_plan()
        x = self._check_dev(x, tuple(self.Nd) + (self.batch,), 'x')
        if self._sense is not None:
            x = x * self._sense.conj()
        return x.mean(dim=-1)

    def forward_one2many(self, s):
        self._require_plan()
        host = not isinstance(s, torch.Tensor)
        gs = self.to_device(s) if host else s
        gs = self._check_dev(gs, self.Nd, 's')
        if gs.dim() != self.ndims:
            raise ValueError('s must have shape Nd')
        y = torch.empty((self.M, self.batch), dtype=torch.complex64, device=self.device)
        sens = _ptr(self._sense) if self._sense is not None else None
        _lib.check(self._lib.b200nufft_forward_one2many(self._plan, _ptr(gs), sens, _ptr(y), self.batch, self._stream()))
        return self.to_host(y) if host else y

    def adjoint_many2one(self, y):
        self._require_plan()
        host = not isinstance(y, torch.Tensor)
        gy = self.to_device(y) if host else y
        gy = self._check_dev(gy, (self.M, self.batch), 'y')
        s = torch.empty(tuple(self.Nd), dtype=torch.complex64, device=self.device)
        sens = _ptr(self._sense) if self._sense is not None else None
        _lib.check(self._lib.b200nufft_adjoint_many2one(self._plan, _ptr(gy), sens, _ptr(s), self.batch, self._stream()))
        return self.to_host(s) if host else s

    def selfadjoint_one2many2one(self, s):
        host = not isinstance(s, torch.Tensor)
        gs = self.to_device(s) if host else s
        out = self.adjoint_many2one(self.forward_one2many(gs))
        return self.to_host(out) if host else out

    # ------------------------------------------------------------------ host API (numpy in / out)
    def _host_c64(self, a, shape_prefix, what):
        a = numpy.ascontiguousarray(numpy.asarray(a, dtype=self.dtype))     # no copy if already complex64 C-order
        if tuple(a.shape[:len(shape_prefix)]) != tuple(shape_prefix) or a.ndim > len(shape_prefix) + 1:
            raise ValueError('%s has shape %s, expected %s(+batch)' % (what, a.shape, tuple(shape_prefix)))
        return a

    def _host_out(self, out, shape, what):
        if out is None:
            return numpy.empty(shape, dtype=self.dtype)
        if not (isinstance(out, numpy.ndarray) and out.dtype == self.dtype and out.flags.c_contiguous
                and tuple(out.shape) == tuple(shape)):
            raise ValueError('%s must be a C-contiguous complex64 array of shape %s' % (what, tuple(shape)))
        return out

    def forward(self, x, out=None, slot=None):
        """Host forward NUFFT (reference: _forward_host).  `out` may be a preallocated (pinned) array.

        slot=None: blocking, like the reference.  slot=0..3: pipelined -- the call only enqueues H2D copy, operator
        and D2H copy (they overlap with the neighbouring calls' copies and kernels) and returns `out` immediately;
        call wait('forward', slot) before reading it or reusing x / out.  Use pinned arrays and cycle through three
        slots to keep copy-in, kernels and copy-out busy at the same time."""
        self._require_plan()
        x = self._host_c64(x, self.Nd, 'x')
        nb = x.shape[-1] if x.ndim == self.ndims + 1 else 1
        y = self._host_out(out, (self.M, nb) if x.ndim == self.ndims + 1 else (self.M,), 'out')
        with torch.cuda.device(self.device):
            if slot is None:
                _lib.check(self._lib.b200nufft_forward_host(self._plan, x.ctypes.data, y.ctypes.data, nb, self._stream()))
            else:
                self._inflight[(0, int(slot))] = (x, y)      # keep the host arrays alive until wait()
                _lib.check(self._lib.b200nufft_forward_host_async(self._plan, x.ctypes.data, y.ctypes.data, nb,
                                                                  int(slot), self._stream()))
        return y

    def adjoint(self, y, out=None, slot=None):
        """Host adjoint NUFFT (reference: _adjoint_host).  slot: as in forward (wait('adjoint', slot))."""
        self._require_plan()
        y = self._host_c64(y, (self.M,), 'y')
        nb = y.shape[-1] if y.ndim == 2 else 1
        x = self._host_out(out, tuple(self.Nd) + ((nb,) if y.ndim == 2 else ()), 'out')
        with torch.cuda.device(self.device):
            if slot is None:
                _lib.check(self._lib.b200nufft_adjoint_host(self._plan, y.ctypes.data, x.ctypes.data, nb, self._stream()))
            else:
                self._inflight[(1, int(slot))] = (y, x)
                _lib.check(self._lib.b200nufft_adjoint_host_async(self._plan, y.ctypes.data, x.ctypes.data, nb,
                                                                  int(slot), self._stream()))
        return x

    def wait(self, op, slot):
        """Block until the last pipelined forward / adjoint on `slot` has delivered its output to the host."""
        self._require_plan()
        o = {'forward': 0, 'adjoint': 1}[op]
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200nufft_host_wait(self._plan, o, int(slot)))
        self._inflight.pop((o, int(slot)), None)

    def selfadjoint(self, x):
        return self.to_host(self._selfadjoint_device(self.to_device(x)))

    def selfadjoint2(self, x):
        return self.to_host(self._selfadjoint2_device(self.to_device(x)))

    def solve(self, y, *args, **kwargs):
        return self.to_host(self._solve_device(self.to_device(y), *args, **kwargs))

    def x2xx(self, x):
        return self.to_host(self._x2xx_device(self.to_device(x)))

    def xx2x(self, xx):
        return self.to_host(self._xx2x_device(self.to_device(xx)))

    def xx2k(self, xx):
        return self.to_host(self._xx2k_device(self.to_device(xx)))

    def k2xx(self, k):
        return self.to_host(self._k2xx_device(self.to_device(k)))

    def k2y(self, k):
        return self.to_host(self._k2y_device(self.to_device(k)))

    def y2k(self, y):
        return self.to_host(self._y2k_device(self.to_device(y)))

    # ------------------------------------------------------------------ parity / introspection
    def _plan_arrays(self):
        """(kindx uint32 (M,sumJ), udata c64 (M,sumJ), k0 int32 (M,d), perm int32 (M,), tile, sub) as numpy."""
        self._require_plan()
        sumJ = int(numpy.sum(self.Jd))
        dev = self.device
        kindx = torch.empty((self.M, sumJ), dtype=torch.int32, device=dev)
        udata = torch.empty((self.M, sumJ), dtype=torch.complex64, device=dev)
        k0 = torch.empty((self.M, self.ndims), dtype=torch.int32, device=dev)
        perm = torch.empty((self.M,), dtype=torch.int32, device=dev)
        tile = numpy.zeros(2 * self.ndims, dtype=numpy.int32)
        _lib.check(self._lib.b200nufft_plan_get_kindx(self._plan, _ptr(kindx), self._stream()))
        _lib.check(self._lib.b200nufft_plan_get_udata(self._plan, _ptr(udata), self._stream()))
        _lib.check(self._lib.b200nufft_plan_get_k0(self._plan, _ptr(k0), self._stream()))
        _lib.check(self._lib.b200nufft_plan_get_perm(self._plan, _ptr(perm), self._stream()))
        _lib.check(self._lib.b200nufft_plan_get_tile(self._plan, tile.ctypes.data))
        return (kindx.cpu().numpy().view(numpy.uint32), udata.cpu().numpy(), k0.cpu().numpy(), perm.cpu().numpy(),
                tuple(int(v) for v in tile[:self.ndims]), tuple(int(v) for v in tile[self.ndims:]))

    def _col_perm(self):
        """(perm int32 (M,), tile, sub): sweep-order permutation of the column-sweep gridding records and its sort key."""
        self._require_plan()
        perm = torch.empty((self.M,), dtype=torch.int32, device=self.device)
        tile = numpy.zeros(6, dtype=numpy.int32)
        _lib.check(self._lib.b200nufft_plan_get_col_perm(self._plan, _ptr(perm), tile.ctypes.data, self._stream()))
        nd = self.ndims
        return perm.cpu().numpy(), tuple(int(v) for v in tile[:nd]), tuple(int(v) for v in tile[nd:2 * nd])

    def layout(self):
        """1 if the plan also holds sweep-ordered records (3-D J = 6: csrc/col3d.cu; 2-D J = 6: csrc/sweep2d.cu), else 0."""
        self._require_plan()
        return int(self._lib.b200nufft_plan_get_layout(self._plan))

    def set_variant(self, interp=0, gridding=0):
        """0 auto, 1 generic kernels, 2 tiled kernels (error if the geometry is unsupported); interp=3 selects the
        column-sweep gather (csrc/col3d.cu) everywhere; auto uses it on phase-modulated grids (forward on the fused
        FFT passes, k-space solvers) and the tiled gather on true grids."""
        self._require_plan()
        _lib.check(self._lib.b200nufft_set_variant(self._plan, int(interp), int(gridding)))
