"""
CPU oracle for the PyNUFFT device NUFFT hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy/scipy restatement of the reference algorithm (the
Fessler-Sutton min-max NUFFT as implemented by pynufft).  Nothing under
``pynufft_b200/`` may import it; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs do, and there only as the
checker / reported CPU baseline.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from /root/reference (CPU path, ``NUFFT()`` and
``helper.plan(format='pELL')``) and stores its outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against them
(indices bit-exact, values to ~1e-6).  The reference ships no golden vectors of
its own.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
import numpy
import scipy.linalg
import scipy.sparse
import scipy.special

dtype = numpy.complex64


# --------------------------------------------------------------------------- #
# plan math, host part (per dimension constants)
# --------------------------------------------------------------------------- #
_KB_BEST = {2: 2.5, 3: 2.27, 4: 2.31, 5: 2.34, 6: 2.32, 7: 2.32, 8: 2.35, 9: 2.34,
            10: 2.34, 11: 2.35, 12: 2.34, 13: 2.35, 14: 2.35, 15: 2.35, 16: 2.33}


def kaiser_bessel_params(J, K_N):
    """src/_helper/helper.py:961-997 -> (kb_a, kb_m)."""
    if K_N != 2:
        return 2.34 * J, 0
    if J in _KB_BEST:
        return J * _KB_BEST[J], 0
    keys = numpy.array(sorted(_KB_BEST.keys()))
    return J * _KB_BEST[int(keys[numpy.argmin(abs(keys - J))])], 0


def kaiser_bessel_ft(u, J, alpha, kb_m, d):
    """src/_helper/helper.py:1000-1012."""
    u = u * (1.0 + 0.0j)
    z = numpy.sqrt((2 * numpy.pi * (J / 2) * u) ** 2.0 - alpha ** 2.0)
    nu = d / 2 + kb_m
    y = ((2 * numpy.pi) ** (d / 2)) * ((J / 2) ** d) * (alpha ** kb_m) / \
        scipy.special.iv(kb_m, alpha) * scipy.special.jv(nu, z) / (z ** nu)
    return numpy.real(y)


def alpha_kb_fit(N, J, K):
    """src/_helper/helper.py:910-958.  Returns (alpha float32 (L+1,), beta)."""
    beta = 1
    Nmid = (N - 1.0) / 2.0
    L = 13 if N > 40 else int(numpy.ceil(N / 3))
    nlist = numpy.arange(0, N) * 1.0 - Nmid
    kb_a, kb_m = kaiser_bessel_params(J, K / N)
    if J > 1:
        sn_kaiser = 1 / kaiser_bessel_ft(nlist / K, J, kb_a, kb_m, 1.0)
    else:
        sn_kaiser = numpy.ones((1, N), dtype=dtype)
    gam = 2 * numpy.pi / K
    X = numpy.cos(numpy.dot((beta * gam * nlist).reshape((N, 1)),
                            numpy.arange(0, L + 1).reshape((1, L + 1))))
    sn_kaiser = sn_kaiser.reshape((N, 1), order='F').conj()
    X = numpy.array(X, dtype=dtype)                       # complex64 least squares
    sn_kaiser = numpy.array(sn_kaiser, dtype=dtype)
    coef = numpy.linalg.lstsq(numpy.nan_to_num(X), numpy.nan_to_num(sn_kaiser), rcond=-1)[0]
    alphas = coef
    if J > 1:
        alphas[1:] = alphas[1:] / 2.0
    else:
        alphas[0] = 1.0
        alphas[1:] = 0.0
    return numpy.real(alphas).reshape(-1), beta            # float32


def scale1(N, K, alpha, beta):
    """src/_helper/helper.py:1015-1036 (real part is what is used, :274)."""
    Nmid = (N - 1) / 2.0
    L = len(alpha) - 1
    if L > 0:
        sn = numpy.zeros((N,))
        n = numpy.arange(0, N)
        i_gam_n_n0 = 1j * (2 * numpy.pi / K) * (n - Nmid) * beta
        for l1 in range(-L, L + 1):
            alf = alpha[abs(l1)]                       # numpy.float32 scalar
            sn = sn + alf * numpy.exp(i_gam_n_n0 * l1)
    else:
        sn = alpha[0] * numpy.ones((N,))
    return sn


def T_matrix(N, J, K, alpha, beta):
    """src/_helper/helper.py:1060-1083  (J x J, real, pinv of CSSC)."""
    L = numpy.size(alpha) - 1
    cssc = numpy.zeros((J, J))
    j1, j2 = numpy.mgrid[1:J + 1, 1:J + 1]
    overlapping_mat = j2 - j1
    for l1 in range(-L, L + 1):
        for l2 in range(-L, L + 1):
            alf1 = alpha[abs(l1)]                      # float32 * float32 product is float32
            alf2 = alpha[abs(l2)]
            tmp = overlapping_mat + beta * (l1 - l2)
            tmp = numpy.sinc(1.0 * tmp / (1.0 * K / N))
            cssc = cssc + alf1 * alf2 * tmp
    return scipy.linalg.pinv(cssc)


# --------------------------------------------------------------------------- #
# plan math, per sample
# --------------------------------------------------------------------------- #
def offset_k0(om, J, K):
    """src/_helper/helper.py:900-907: k0 = floor(om/gam - J/2), float64."""
    gam = 2.0 * numpy.pi / (K * 1.0)
    return numpy.floor(1.0 * om / gam - 1.0 * J / 2.0)


def indices_1d(om, J, K):
    """src/_helper/helper.py:164-180: (M, J) wrapped grid indices (before stride)."""
    k0 = offset_k0(om, J, K)
    k_indx = numpy.mod(numpy.add.outer(numpy.arange(1, J + 1) * 1.0, k0), K)
    return k_indx.T                                    # (M, J) float64 holding integers


def coeff_1d(om, N, J, K, alpha, beta, T=None):
    """
    min_max + OMEGA_u, src/_helper/helper.py:606-618, 1086-1117, 148-162.
    Returns u (M, J) complex128 (the reference casts to complex64 when packing).
    """
    if T is None:
        T = T_matrix(N, J, K, alpha, beta)
    gam = 2.0 * numpy.pi / (K * 1.0)
    k0 = offset_k0(om, J, K)
    dk = 1.0 * om / gam - k0
    arg = numpy.add.outer(-numpy.arange(1, J + 1) * 1.0, dk)          # (J, M)
    L = numpy.size(alpha) - 1
    rr = numpy.zeros((J, om.shape[0]), dtype=numpy.float32)
    ratio = (1.0 * K / N)
    for l1 in range(-L, L + 1):
        alf = alpha[abs(l1)] * 1.0
        rr = rr + alf * numpy.sinc((arg + 1.0 * l1 * beta) / ratio)
    c = T.dot(rr)
    phase = numpy.exp(1.0j * gam * (N * 1.0 - 1.0) / 2.0 * arg)
    u = numpy.exp(-1.0j * om * N / 2.0) * (phase * c)
    return u.T.conj()                                   # (M, J)


class Plan:
    """
    Restatement of helper.plan(format='pELL') (src/_helper/helper.py:620-802) with the
    pELL / Tensor_sn containers (:212-282, :346-393).  radix == 1, all axes transformed.
    """

    def __init__(self, om, Nd, Kd, Jd):
        if type(Nd) != tuple:
            raise TypeError('Nd must be tuple, e.g. (256, 256)')
        if type(Kd) != tuple:
            raise TypeError('Kd must be tuple, e.g. (512, 512)')
        if type(Jd) != tuple:
            raise TypeError('Jd must be tuple, e.g. (6, 6)')
        if not (len(Nd) == len(Kd) == len(Jd)):
            raise KeyError('Nd, Kd, Jd must be in the same length')
        om = numpy.asarray(om)
        self.Nd, self.Kd, self.Jd = Nd, Kd, Jd
        self.ndims = dd = len(Nd)
        self.M = M = om.shape[0]
        self.om = om
        self.alpha, self.beta, self.snd, self.T = [], [], [], []
        for d in range(dd):
            a, b = alpha_kb_fit(Nd[d], Jd[d], Kd[d])
            self.alpha.append(a)
            self.beta.append(b)
            self.snd.append(scale1(Nd[d], Kd[d], a, b))
            self.T.append(T_matrix(Nd[d], Jd[d], Kd[d], a, b))
        # Tensor_sn (radix 1): concatenated real 1-D scaling vectors, float32
        self.tensor_sn = numpy.concatenate([s.real for s in self.snd]).astype(numpy.float32)
        sumJ = int(numpy.sum(Jd))
        self.kindx = numpy.zeros((M, sumJ), dtype=numpy.uint32)
        self.udata = numpy.zeros((M, sumJ), dtype=numpy.complex128)
        self.k0 = numpy.zeros((M, dd), dtype=numpy.int64)
        off = 0
        for d in range(dd):
            J = Jd[d]
            ki = indices_1d(om[:, d], J, Kd[d])
            if d < dd - 1:
                ki = ki * numpy.prod(Kd[d + 1:dd])                    # helper.py:197-199
            self.kindx[:, off:off + J] = ki
            self.udata[:, off:off + J] = coeff_1d(om[:, d], Nd[d], J, Kd[d], self.alpha[d],
                                                   self.beta[d], self.T[d])
            self.k0[:, d] = offset_k0(om[:, d], J, Kd[d]).astype(numpy.int64)
            off += J
        self.udata = self.udata.astype(numpy.complex64)
        # meshindex: row-major decode of the flat neighbour id (helper.py:376-383)
        prodJ = int(numpy.prod(Jd))
        self.meshindex = numpy.stack(numpy.unravel_index(numpy.arange(prodJ), Jd), axis=1).astype(numpy.uint32)
        self.prodJd, self.sumJd = prodJ, sumJ

    def sn_full(self):
        """float64 outer product of the 1-D scaling vectors (helper.py:558-584), real."""
        sn = numpy.reshape(1.0, (1,) * self.ndims)
        for d in range(self.ndims):
            shp = [1] * self.ndims
            shp[d] = self.Nd[d]
            sn = sn * numpy.reshape(self.snd[d].real, shp)
        return sn

    def columns_weights(self, rows=None):
        """
        Flat grid column and complex weight of every neighbour, as the device kernels form
        them (src/re_subroutine.py:784-806): col = sum_d kindx, w = prod_d udata.
        Returns (col (m, prodJ) int64, w (m, prodJ) complex64-rounded factors multiplied in c128).
        """
        if rows is None:
            rows = slice(None)
        kin = self.kindx[rows]
        ud = self.udata[rows]
        off = 0
        col = None
        w = None
        for d in range(self.ndims):
            J = self.Jd[d]
            kd = kin[:, off:off + J].astype(numpy.int64)
            u = ud[:, off:off + J].astype(numpy.complex128)
            if col is None:
                col, w = kd, u
            else:
                col = (col[:, :, None] + kd[:, None, :]).reshape(kd.shape[0], -1)
                w = (w[:, :, None] * u[:, None, :]).reshape(kd.shape[0], -1)
            off += J
        return col, w

    def csr(self, rows=None):
        """CSR interpolator as helper.full_kron/create_csr build it (helper.py:284-311, 417-428)."""
        col, w = self.columns_weights(rows)
        m, pj = col.shape
        rowptr = numpy.arange(0, (m + 1) * pj, pj)
        return scipy.sparse.csr_matrix((w.astype(numpy.complex64).ravel(), col.ravel(), rowptr),
                                       shape=(m, int(numpy.prod(self.Kd))))


# --------------------------------------------------------------------------- #
# bin / sort permutation contract (new in the B200 build)
# --------------------------------------------------------------------------- #
def column_sweep_tiles(Kd, Jd):
    """(tile, sub) of the second sort the plan keeps for the column-sweep gridding kernel (csrc/col3d.cu
    col3d_supported: 3-D, J = 6), or None.  The key (column(q1, q2) * K0 + first plane) is the generic bin key
    with tile = (K0, 4, 5), sub-tile = (1, 4, 5)."""
    if (len(Kd) == 3 and tuple(Jd) == (6, 6, 6) and Kd[0] >= 6 and Kd[1] >= 9 and Kd[2] >= 10):
        return (int(Kd[0]), 4, 5), (1, 4, 5)
    # 2-D multi-coil row sweep (csrc/sweep2d.cu sweep2d_supported): key = strip(3 columns) * K0 + first row
    if (len(Kd) == 2 and tuple(Jd) == (6, 6) and Kd[0] >= 16 and Kd[1] >= 8):
        return (int(Kd[0]), 3), (1, 3)
    return None


def default_tiles(Kd):
    """Tile / sub-tile edges the plan uses (csrc/plan.cu choose_tiles)."""
    nd = len(Kd)
    want = 16 if nd == 3 else (32 if nd == 2 else 256)
    tile = tuple(min(want, k) for k in Kd)
    sub = tuple(8 if (nd == 3 and t % 8 == 0) else t for t in tile)
    if nd == 2:
        sub = (8 if tile[0] % 8 == 0 else tile[0], 16 if tile[1] % 16 == 0 else tile[1])
    return tile, sub


def bin_keys(k0, Kd, tile, sub):
    """
    key = tile_linear * subtiles_per_tile + subtile_linear (both C-order) of the cell that holds
    the first neighbour (k0+1) mod K.  k0: (M, d) integers from offset_k0 (bit-exact contract).
    """
    k0 = numpy.asarray(k0, dtype=numpy.int64)
    key = numpy.zeros(k0.shape[0], dtype=numpy.int64)
    skey = numpy.zeros(k0.shape[0], dtype=numpy.int64)
    nsubprod = 1
    for d in range(len(Kd)):
        ks = numpy.mod(k0[:, d] + 1, Kd[d])
        q = ks // tile[d]
        ntile = -(-Kd[d] // tile[d])
        nsub = -(-tile[d] // sub[d])
        key = key * ntile + q
        skey = skey * nsub + (ks - q * tile[d]) // sub[d]
        nsubprod *= nsub
    return key * nsubprod + skey


def sort_permutation(k0, Kd, tile, sub):
    """Stable sort by bin key: the permutation the GPU plan must reproduce bit-exactly."""
    return numpy.argsort(bin_keys(k0, Kd, tile, sub), kind='stable')


# --------------------------------------------------------------------------- #
# operator stages (device semantics = CPU semantics; nufft/_nufft_class_methods_cpu.py)
# --------------------------------------------------------------------------- #
class NUFFT:
    """
    CPU restatement of the NUFFT operator: stage methods follow
    nufft/_nufft_class_methods_cpu.py:168-362; the interpolation is the CSR SpMV the
    reference CPU path uses (scipy), the FFT is numpy.fft (as the reference).
    Multi-coil helpers follow linalg/nufft_hsa.py:358-388, 628-656 (== nufft_cpu.py:177-203).
    """

    def __init__(self):
        self.dtype = numpy.complex64

    def plan(self, om, Nd, Kd, Jd, batch=None):
        self.p = Plan(om, Nd, Kd, Jd)
        self.Nd, self.Kd, self.Jd = Nd, Kd, Jd
        self.ndims = len(Nd)
        self.M = self.p.M
        self.batch = 1 if batch is None else int(batch)
        self.sn = numpy.asarray(self.p.sn_full().astype(self.dtype), order='C')
        self.sp = self.p.csr()
        self.spH = self.sp.getH().tocsr()
        self.Kdprod = int(numpy.prod(Kd))
        self.Ndprod = int(numpy.prod(Nd))
        self.ft_axes = tuple(range(self.ndims))
        self._corner = tuple(slice(0, n) for n in Nd)
        self.sense = None
        return 0

    # --- single-coil stages; a trailing batch axis is broadcast through every stage ---
    def _b(self, a, base_ndim):
        return a.ndim == base_ndim + 1

    def x2xx(self, x):
        sn = self.sn[..., None] if self._b(x, self.ndims) else self.sn
        return x * sn

    def xx2k(self, xx):
        shp = self.Kd + ((xx.shape[-1],) if self._b(xx, self.ndims) else ())
        out = numpy.zeros(shp, dtype=self.dtype, order='C')
        out[self._corner] = xx                          # corner-aligned zero padding (re_subroutine.py:120-131)
        return numpy.fft.fftn(out, axes=self.ft_axes)

    def k2y(self, k):
        if self._b(k, self.ndims):
            return self.sp.dot(k.reshape(self.Kdprod, -1))
        return self.sp.dot(k.reshape(self.Kdprod))

    def y2k(self, y):
        kv = self.spH.dot(y)
        return kv.reshape(self.Kd + ((y.shape[1],) if y.ndim == 2 else ()))

    def k2xx(self, k):
        k = numpy.fft.ifftn(k, axes=self.ft_axes)
        return numpy.ascontiguousarray(k[self._corner]).astype(self.dtype)

    def xx2x(self, xx):
        return self.x2xx(xx)

    def forward(self, x):
        return self.k2y(self.xx2k(self.x2xx(x)))

    def adjoint(self, y):
        return self.xx2x(self.k2xx(self.y2k(y)))

    def selfadjoint(self, x):
        return self.adjoint(self.forward(x))

    def selfadjoint2(self, x):
        """Toeplitz-style approximation of A^H A (nufft/_nufft_class_methods_cpu.py:127-146, 216-222):
        W = |xx2k(adjoint(1_M))| once, then k2xx(W * xx2k(x)) -- no sn scaling, no interpolation."""
        if getattr(self, 'W', None) is None:
            W = self.xx2k(self.adjoint(numpy.ones((self.M,), dtype=numpy.complex64)))
            self.W = (W * W.conj()) ** 0.5
        return self.k2xx(self.W * self.xx2k(x))

    # --- multi-coil (linalg/nufft_hsa.py:333-388, 628-672, 723-769) ---
    def set_sense(self, coil_profile):
        if coil_profile.shape != self.Nd + (self.batch,):
            raise ValueError
        self.sense = coil_profile.astype(self.dtype)

    def reset_sense(self):
        self.sense = None

    def s2x(self, s):
        x = numpy.repeat(s[..., None], self.batch, axis=-1).astype(self.dtype)   # cPopulate
        if self.sense is not None:
            x = x * self.sense                                                    # cMultiplyVecInplace
        return x

    def x2s(self, x):
        if self.sense is not None:
            x = x * self.sense.conj()                                             # cMultiplyConjVecInplace
        return numpy.mean(x, axis=-1)                                             # cAggregate: mean over coils

    def forward_one2many(self, s):
        return self.forward(self.s2x(s))

    def adjoint_many2one(self, y):
        return self.x2s(self.adjoint(y))

    def selfadjoint_one2many2one(self, s):
        return self.adjoint_many2one(self.forward_one2many(s))


# --------------------------------------------------------------------------- #
# device solvers (linalg/solve_device.py), restated on numpy in complex64
# --------------------------------------------------------------------------- #
def solve_cg(nufft, y, maxiter=30, dtype=numpy.complex64):
    """
    linalg/solve_device.py:351-481 (batched twin linalg/solve_hsa.py:551-682):
    CG on G = interp^H interp over the oversampled grid, x0 = b, exactly `maxiter`
    steps, one alpha/beta for all coils; then k2xx and DIVIDE by sn.
    dtype=complex64 mirrors the device arithmetic; complex128 gives the exact-arithmetic iterates (used to
    judge how much float32 rounding the iteration amplifies on ill-conditioned trajectories).
    """
    c = dtype
    sp, spH = nufft.sp.astype(c), nufft.spH.astype(c)
    shp = lambda v: v.reshape(nufft.Kdprod, -1) if v.ndim == nufft.ndims + 1 else v.reshape(nufft.Kdprod)
    G = lambda v: spH.dot(sp.dot(shp(v))).reshape(v.shape).astype(c)
    b = nufft.y2k(y).astype(c)
    x = b.copy()
    r = (b - G(x)).astype(c)
    p = r.copy()
    rsold = c(numpy.sum(numpy.conj(r) * r))
    for _ in range(maxiter):
        Ap = G(p)
        alpha = c(rsold / c(numpy.sum(numpy.conj(p) * Ap)))
        x = (x + alpha * p).astype(c)
        r = (r - alpha * Ap).astype(c)
        rsnew = c(numpy.sum(numpy.conj(r) * r))
        beta = c(rsnew / rsold)
        p = (r + beta * p).astype(c)
        rsold = rsnew
    x2 = numpy.fft.ifftn(x, axes=nufft.ft_axes)[nufft._corner]
    sn = nufft.sn.real
    if x2.ndim == nufft.ndims + 1:
        sn = sn[..., None]
    return (x2 / sn).astype(c)


def solve_krylov(nufft, y, solver, *args, **kwargs):
    """The scipy Krylov family of the CPU solve, linalg/solve_cpu.py:226-288: 'lsmr' / 'lsqr' on the rectangular
    interpolator A = k2y (rmatvec y2k), everything else ('bicgstab', 'bicg', 'gmres', 'lgmres', scipy 'cg') on
    G = y2k . k2y with right-hand side y2k(y); the k-space solution goes through k2xx and is DIVIDED by sn."""
    import scipy.sparse.linalg as sla
    Kd, K, M = tuple(nufft.Kd), nufft.Kdprod, nufft.M
    if solver in ('lsmr', 'lsqr'):
        A = sla.LinearOperator((M, K), matvec=lambda k: nufft.k2y(k.reshape(Kd)).ravel(),
                               rmatvec=lambda v: nufft.y2k(v.reshape(M)).ravel(), dtype=numpy.complex128)
        vec = {'lsmr': sla.lsmr, 'lsqr': sla.lsqr}[solver](A, numpy.asarray(y).ravel(), *args, **kwargs)[0]
    else:
        G = lambda k: nufft.y2k(nufft.k2y(k.reshape(Kd))).ravel()
        A = sla.LinearOperator((K, K), matvec=G, rmatvec=G, dtype=numpy.complex128)
        methods = {'cg': sla.cg, 'bicgstab': sla.bicgstab, 'bicg': sla.bicg, 'gmres': sla.gmres, 'lgmres': sla.lgmres}
        vec = methods[solver](A, nufft.y2k(y).ravel(), *args, **kwargs)[0]
    return nufft.k2xx(vec.reshape(Kd)) / nufft.sn


def solve_dc(nufft, y, maxiter=1):
    """'dc' (Pipe density compensation), linalg/solve_cpu.py:165-225: W = 1; W <- W / (A A^H W) `maxiter`
    times; x = A^H (W y)."""
    W = numpy.ones(nufft.M, dtype=numpy.complex64)
    for _ in range(maxiter):
        W = W / nufft.forward(nufft.adjoint(W))
    return nufft.adjoint(W * y)


def laplacian_kernel(Kd):
    """src/_helper/helper.py:11-45."""
    nd = len(Kd)
    uker = numpy.zeros(Kd, dtype=numpy.complex64)
    uker[(0,) * nd] = -2.0 * nd
    for pp in range(nd):
        i1 = [0] * nd
        i1[pp] = 1
        uker[tuple(i1)] = 1
        i1[pp] = -1
        uker[tuple(i1)] = 1
    return numpy.fft.fftn(uker)


def solve_l1tvlad(nufft, y, maxiter, rho):
    """linalg/solve_hsa.py:74-274: L1TVOLS plus a second split variable for the data term (least absolute deviation)."""
    return solve_l1tvols(nufft, y, maxiter, rho, lad=True)


def solve_l1tvols(nufft, y, maxiter, rho, lad=False):
    """linalg/solve_device.py:74-275 (split-Bregman TV, device variant).
    lad=True: the L1TVLAD variant of the batched twin (linalg/solve_hsa.py:74-274): rhs uses AHyk + df - bf (:134-138),
    df = shrink(zf + bf, 1/mu) and bf += zf - df after the TV shrinkage (:229-260).

    Multi-coil data y (M, B) on a batch operator: the same iteration with the closures of the batched twin
    (linalg/solve_hsa.py:275-476, AHA = nufft.selfadjoint, AH = nufft.adjoint at :282-287) acting between the ONE image
    the solver carries (every work array has shape st['Nd'], :291-318) and the B coils, i.e.
    AH = adjoint_many2one and AHA = selfadjoint_one2many2one (linalg/nufft_hsa.py:658-672, 747-769): TV-SENSE.
    The sampling density |y2k(1_M)| and the xx2k / k2xx of the Ku = rhs solve stay single-image (:288, :366-375)."""
    c64 = numpy.complex64
    multi = y.ndim == 2
    AH = nufft.adjoint_many2one if multi else nufft.adjoint
    AHA = nufft.selfadjoint_one2many2one if multi else nufft.selfadjoint
    mu = 1.0
    LMBD = rho * mu
    nd = nufft.ndims
    w = numpy.abs(nufft.y2k(numpy.ones((nufft.M,), dtype=c64)).astype(c64))   # :25-36
    uker = (mu * w - LMBD * laplacian_kernel(nufft.Kd)).astype(c64)
    AHy = AH(y).astype(c64)
    z = numpy.zeros(nufft.Nd, dtype=c64)
    xkp1 = z.copy()
    AHyk = z.copy()
    zz = [z.copy() for _ in range(nd)]
    bb = [z.copy() for _ in range(nd)]
    dd = [z.copy() for _ in range(nd)]
    D = lambda v, ax: (numpy.roll(v, +1, ax) - v).astype(c64)      # d_indx: roll +1  (helper.py:66)
    Dt = lambda v, ax: (numpy.roll(v, -1, ax) - v).astype(c64)     # dt_indx: roll -1 (helper.py:67)

    def pad_fft(v):                                                 # _xx2k_device: no sn scaling
        out = numpy.zeros(nufft.Kd, dtype=c64)
        out[nufft._corner] = v
        return numpy.fft.fftn(out).astype(c64)

    thr = numpy.float32(1.0 / LMBD)
    thr_f = numpy.float32(1.0 / mu)
    bf, df = z.copy(), z.copy()
    for _ in range(maxiter):
        rhs = (c64(mu) * ((AHyk + df).astype(c64) - bf).astype(c64) if lad else c64(mu) * AHyk).astype(c64)
        for pp in range(nd):
            rhs = (rhs + c64(LMBD) * Dt((dd[pp] - bb[pp]).astype(c64), pp)).astype(c64)
        k = (pad_fft(rhs) / uker).astype(c64)
        xkp1 = nufft.k2xx(k).astype(c64)
        for pp in range(nd):
            zz[pp] = D(xkp1, pp)
        zf = (AHA(xkp1).astype(c64) - AHy).astype(c64)
        s_tmp = [(zz[pp] + bb[pp]).astype(c64) for pp in range(nd)]
        s = s_tmp[0].copy()
        for pp in range(1, nd):                                     # cHypot, re_subroutine.py:656-678
            s = numpy.hypot(numpy.abs(s), numpy.abs(s_tmp[pp])).astype(numpy.float32).astype(c64)
        s = (s + c64(1e-6)).astype(c64)
        sr, si = s.real, s.imag                                     # cAnisoShrink, :973-993
        tr = (sr > thr) * (sr - thr) + (sr < -thr) * (sr + thr)
        ti = (si > thr) * (si - thr) + (si < -thr) * (si + thr)
        t = ((tr + 1j * ti).astype(c64) / s).astype(c64)
        for pp in range(nd):
            dd[pp] = (s_tmp[pp] * t).astype(c64)
            bb[pp] = (bb[pp] + (zz[pp] - dd[pp])).astype(c64)
        if lad:
            tf = (zf + bf).astype(c64)
            fr, fi = tf.real, tf.imag
            df = (((fr > thr_f) * (fr - thr_f) + (fr < -thr_f) * (fr + thr_f)) +
                  1j * ((fi > thr_f) * (fi - thr_f) + (fi < -thr_f) * (fi + thr_f))).astype(c64)
            bf = (bf + (zf - df)).astype(c64)
        AHyk = (AHyk - zf).astype(c64)
    return xkp1


# --------------------------------------------------------------------------- #
# ground truth for tiny sizes: exact non-uniform DFT (linalg/nudft_cpu.py:27-41)
# --------------------------------------------------------------------------- #
def nudft_forward(x, om):
    Nd = x.shape
    grids = numpy.meshgrid(*[numpy.arange(n) - n / 2 for n in Nd], indexing='ij')
    ph = numpy.zeros((om.shape[0],) + Nd)
    for d in range(len(Nd)):
        ph = ph + om[:, d].reshape((-1,) + (1,) * len(Nd)) * grids[d]
    return numpy.sum(numpy.exp(-1j * ph) * x, axis=tuple(range(1, len(Nd) + 1)))
