"""
The Krylov family of the reference's CPU solve (linalg/solve_cpu.py:226-288: scipy's lsqr, lsmr, bicgstab, bicg, gmres,
lgmres with k2y / y2k / k2y2k as the operator) with every VECTOR on the device.

The reference hands scipy a LinearOperator and lets it run the recurrences on host arrays; with a device operator that
ships one Kd grid (134 MB at 256^3) over PCIe each way per operator application.  Here the recurrences themselves run on
flat complex64 device vectors: the vector updates are one CUDA kernel (`b200nufft_axpby`), the inner products the
double-accumulating `b200nufft_dotc`; only the scalars of the recurrences (a few doubles per iteration) and the small
Hessenberg / QR problems of gmres / lgmres live on the host, as they do in scipy.

Each routine restates the published algorithm in the form scipy 1.x implements (same stopping rules, same order of
updates, same defaults), so that the iterates follow the reference's: Paige & Saunders (LSQR), Fong & Saunders (LSMR),
van der Vorst (BiCGSTAB), Fletcher (BiCG), Saad & Schultz (GMRES with Givens rotations), Baker, Jessup & Manteuffel
(LGMRES).  The vector type of the reference's run is complex64 (`make_system` takes it from the operator and from
b = y2k(y)), so eps below is float32's.

The routines only touch vectors through an `ops` object (`dot`, `nrm`, `axpby`, `new`), which lets the CPU tests drive
the same control flow with torch CPU tensors against scipy itself (tests/test_krylov_host.py).
"""
import math

import numpy
import torch

EPS32 = float(numpy.finfo(numpy.float32).eps)
EPS64 = float(numpy.finfo(numpy.float64).eps)


class TorchVectorOps:
    """Reference implementation of the vector interface on torch tensors (any device); used by the CPU tests."""

    def new(self, like):
        return torch.empty_like(like)

    def dot(self, a, b):
        """conj(a) . b as a Python complex (double accumulation)"""
        return complex(torch.vdot(a.to(torch.complex128), b.to(torch.complex128)).item())

    def nrm(self, a):
        return math.sqrt(max(self.dot(a, a).real, 0.0))

    def axpby(self, out, a, x, b=0.0, y=None):
        """out = a x + b y   (y None or b == 0: out = a x); out may alias x or y"""
        if y is None or b == 0:
            torch.mul(x, complex(a), out=out)
        else:
            t = x * complex(a)
            torch.mul(y, complex(b), out=out)
            out.add_(t)
        return out


class _V:
    """small helpers on top of ops"""

    def __init__(self, ops):
        self.o = ops
        self.dot, self.nrm = ops.dot, ops.nrm

    def copy(self, x):
        return self.o.axpby(self.o.new(x), 1.0, x)

    def scal(self, x, a):
        return self.o.axpby(x, a, x)

    def axpy(self, y, a, x):
        """y += a x"""
        return self.o.axpby(y, a, x, 1.0, y)

    def xpay(self, y, x, a):
        """y = x + a y"""
        return self.o.axpby(y, 1.0, x, a, y)

    def comb(self, a, x, b, y):
        return self.o.axpby(self.o.new(x), a, x, b, y)


def _tolerances(name, bnrm2, atol, rtol):
    if atol is None or atol == 'legacy' or atol < 0:
        raise ValueError("'%s' called with invalid `atol`=%r; if set, `atol` must be a real, non-negative number."
                         % (name, atol))
    return max(float(atol), float(rtol) * float(bnrm2))


def _sym_ortho(a, b):
    """stable Givens rotation of LSQR / LSMR (Choi's SymOrtho): c a + s b = r, real arguments"""
    sign = lambda t: (t > 0) - (t < 0)
    if b == 0:
        return sign(a), 0, abs(a)
    if a == 0:
        return 0, sign(b), abs(b)
    if abs(b) > abs(a):
        tau = a / b
        s = sign(b) / math.sqrt(1 + tau * tau)
        return s * tau, s, b / s
    tau = b / a
    c = sign(a) / math.sqrt(1 + tau * tau)
    return c, c * tau, a / c


def _start(V, matvec, b, x0):
    """x and r = b - A x of the square solvers (make_system + `r = b - matvec(x) if x.any() else b.copy()`)"""
    if x0 is None:
        x = V.o.axpby(V.o.new(b), 0.0, b)
        return x, V.copy(b)
    x = V.copy(x0)
    if V.nrm(x) == 0:
        return x, V.copy(b)
    return x, V.comb(1.0, b, -1.0, matvec(x))


def bicgstab(ops, matvec, b, x0=None, rtol=1e-5, atol=0.0, maxiter=None, M=None, callback=None):
    V = _V(ops)
    psolve = M if M is not None else (lambda v: v)
    bnrm2 = V.nrm(b)
    atol = _tolerances('bicgstab', bnrm2, atol, rtol)
    if bnrm2 == 0:
        return b, 0
    if maxiter is None:
        maxiter = b.numel() * 10
    rhotol = omegatol = EPS32 ** 2
    x, r = _start(V, matvec, b, x0)
    rtilde = V.copy(r)
    rho_prev = omega = alpha = p = v = None
    for iteration in range(maxiter):
        if V.nrm(r) < atol:
            return x, 0
        rho = V.dot(rtilde, r)
        if abs(rho) < rhotol:
            return x, -10
        if iteration > 0:
            if abs(omega) < omegatol:
                return x, -11
            beta = (rho / rho_prev) * (alpha / omega)
            V.axpy(p, -omega, v)                     # p -= omega v ; p *= beta ; p += r
            V.xpay(p, r, beta)
        else:
            p = V.copy(r)
        phat = psolve(p)
        v = matvec(phat)
        rv = V.dot(rtilde, v)
        if rv == 0:
            return x, -11
        alpha = rho / rv
        V.axpy(r, -alpha, v)                         # s = r
        if V.nrm(r) < atol:
            V.axpy(x, alpha, phat)
            return x, 0
        shat = psolve(r)
        t = matvec(shat)
        omega = V.dot(t, r) / V.dot(t, t)
        V.axpy(x, alpha, phat)
        V.axpy(x, omega, shat)
        V.axpy(r, -omega, t)
        rho_prev = rho
        if callback:
            callback(x)
    return x, maxiter


def bicg(ops, matvec, rmatvec, b, x0=None, rtol=1e-5, atol=0.0, maxiter=None, M=None, MH=None, callback=None):
    V = _V(ops)
    psolve = M if M is not None else (lambda v: v)
    rpsolve = MH if MH is not None else (lambda v: v)
    bnrm2 = V.nrm(b)
    atol = _tolerances('bicg', bnrm2, atol, rtol)
    if bnrm2 == 0:
        return b, 0
    if maxiter is None:
        maxiter = b.numel() * 10
    rhotol = EPS32 ** 2
    x, r = _start(V, matvec, b, x0)
    rtilde = V.copy(r)
    rho_prev = p = ptilde = None
    for iteration in range(maxiter):
        if V.nrm(r) < atol:
            return x, 0
        z = psolve(r)
        ztilde = rpsolve(rtilde)
        rho_cur = V.dot(rtilde, z)
        if abs(rho_cur) < rhotol:
            return x, -10
        if iteration > 0:
            beta = rho_cur / rho_prev
            V.xpay(p, z, beta)
            V.xpay(ptilde, ztilde, beta.conjugate())
        else:
            p = V.copy(z)
            ptilde = V.copy(ztilde)
        q = matvec(p)
        qtilde = rmatvec(ptilde)
        rv = V.dot(ptilde, q)
        if rv == 0:
            return x, -11
        alpha = rho_cur / rv
        V.axpy(x, alpha, p)
        V.axpy(r, -alpha, q)
        V.axpy(rtilde, -alpha.conjugate(), qtilde)
        rho_prev = rho_cur
        if callback:
            callback(x)
    return x, maxiter


def _lartg(f, g):
    """LAPACK ?lartg on host scalars: [c s; -conj(s) c] [f; g] = [r; 0], c real"""
    from scipy.linalg import get_lapack_funcs
    fn = get_lapack_funcs('lartg', dtype=numpy.complex64)
    c, s, r = fn(numpy.complex64(f), numpy.complex64(g))
    return complex(c), complex(s), complex(r)


def gmres(ops, matvec, b, x0=None, rtol=1e-5, atol=0.0, restart=None, maxiter=None, M=None, callback=None,
          callback_type=None):
    """restarted GMRES, modified Gram-Schmidt + Givens rotations; `maxiter` counts restart cycles, except with a
    callback and callback_type 'legacy' (or none given), where scipy counts inner iterations."""
    V = _V(ops)
    if callback_type is None:
        callback_type = 'legacy'
    if callback_type not in ('x', 'pr_norm', 'legacy'):
        raise ValueError('Unknown callback_type: %r' % (callback_type,))
    if callback is None:
        callback_type = None
    legacy = callback_type == 'legacy'
    psolve = M if M is not None else (lambda v: v)
    n = b.numel()
    bnrm2 = V.nrm(b)
    atol = _tolerances('gmres', bnrm2, atol, rtol)
    if bnrm2 == 0:
        return b, 0
    eps = EPS32
    if maxiter is None:
        maxiter = n * 10
    restart = min(20 if restart is None else restart, n)
    Mb_nrm2 = V.nrm(psolve(b))
    ptol_max_factor = 1.0
    ptol = Mb_nrm2 * min(ptol_max_factor, atol / bnrm2)
    presid = 0.0
    c64 = numpy.complex64
    h = numpy.zeros((restart, restart + 1), dtype=c64)
    givens = numpy.zeros((restart, 2), dtype=c64)
    inner_iter = 0
    x = rnorm = None
    for iteration in range(maxiter):
        if iteration == 0:
            x, r = _start(V, matvec, b, x0)
            if V.nrm(r) < atol:
                return x, 0
        v = [None] * (restart + 1)
        v[0] = V.copy(psolve(r)) if M is not None else r
        tmp = V.nrm(v[0])
        V.scal(v[0], 1 / tmp)
        S = numpy.zeros(restart + 1, dtype=c64)
        S[0] = tmp
        breakdown = False
        col = 0
        for col in range(restart):
            w = psolve(matvec(v[col]))
            h0 = V.nrm(w)
            for k in range(col + 1):
                tmp = V.dot(v[k], w)
                h[col, k] = tmp
                V.axpy(w, -complex(h[col, k]), v[k])
            h1 = V.nrm(w)
            h[col, col + 1] = h1
            v[col + 1] = w
            if h1 <= eps * h0:
                h[col, col + 1] = 0
                breakdown = True
            else:
                V.scal(w, 1 / h1)
            for k in range(col):
                c, s = givens[k, 0], givens[k, 1]
                n0, n1 = h[col, [k, k + 1]]
                h[col, [k, k + 1]] = [c * n0 + s * n1, -s.conj() * n0 + c * n1]
            c, s, mag = _lartg(h[col, col], h[col, col + 1])
            givens[col, :] = [c, s]
            h[col, [col, col + 1]] = mag, 0
            tmp = -numpy.conjugate(c64(s)) * S[col]
            S[[col, col + 1]] = [c64(c) * S[col], tmp]
            presid = float(numpy.abs(tmp))
            inner_iter += 1
            if callback is not None and callback_type in ('legacy', 'pr_norm'):
                callback(presid / bnrm2)
            if legacy and inner_iter == maxiter:
                break
            if presid <= ptol or breakdown:
                break
        if h[col, col] == 0:
            S[col] = 0
        y = numpy.zeros(col + 1, dtype=c64)
        y[:] = S[:col + 1]
        for k in range(col, 0, -1):
            if y[k] != 0:
                y[k] /= h[k, k]
                tmp = y[k]
                y[:k] -= tmp * h[k, :k]
        if y[0] != 0:
            y[0] /= h[0, 0]
        for k in range(col + 1):
            V.axpy(x, complex(y[k]), v[k])
        r = V.comb(1.0, b, -1.0, matvec(x))
        rnorm = V.nrm(r)
        if legacy and inner_iter == maxiter:
            return x, (0 if rnorm <= atol else maxiter)
        if callback is not None and callback_type == 'x':
            callback(x)
        if rnorm <= atol:
            break
        elif breakdown:
            break
        elif presid <= ptol:
            ptol_max_factor = max(eps, 0.25 * ptol_max_factor)
        else:
            ptol_max_factor = min(1.0, 1.5 * ptol_max_factor)
        ptol = presid * min(ptol_max_factor, atol / rnorm)
    return x, (0 if rnorm <= atol else maxiter)


def _fgmres(V, matvec, v0, m, atol, lpsolve, outer_v, prepend_outer_v):
    """flexible-GMRES Arnoldi process with LGMRES augmentation vectors; H kept as a QR factorisation on the host"""
    from scipy.linalg import qr_insert, lstsq
    c64 = numpy.complex64
    vs, zs = [v0], []
    res = float('nan')
    m = m + len(outer_v)
    Q = numpy.ones((1, 1), dtype=c64)
    R = numpy.zeros((1, 0), dtype=c64)
    eps = EPS32
    breakdown = False
    j = 0
    for j in range(m):
        if prepend_outer_v and j < len(outer_v):
            z, w = outer_v[j]
        elif prepend_outer_v and j == len(outer_v):
            z, w = v0, None
        elif not prepend_outer_v and j >= m - len(outer_v):
            z, w = outer_v[j - (m - len(outer_v))]
        else:
            z, w = vs[-1], None
        if w is None:
            w = lpsolve(matvec(z))
        else:
            w = V.copy(w)
        w_norm = V.nrm(w)
        hcur = numpy.zeros(j + 2, dtype=Q.dtype)
        for i, v in enumerate(vs):
            alpha = V.dot(v, w)
            hcur[i] = alpha
            V.axpy(w, -complex(hcur[i]), v)
        hcur[len(vs)] = V.nrm(w)
        with numpy.errstate(over='ignore', divide='ignore'):
            alpha = 1 / hcur[-1]
        if numpy.isfinite(alpha):
            V.scal(w, complex(alpha))
        if not (hcur[-1].real > eps * w_norm):
            breakdown = True
        vs.append(w)
        zs.append(z)
        Q2 = numpy.zeros((j + 2, j + 2), dtype=Q.dtype, order='F')
        Q2[:j + 1, :j + 1] = Q
        Q2[j + 1, j + 1] = 1
        R2 = numpy.zeros((j + 2, j), dtype=R.dtype, order='F')
        R2[:j + 1, :] = R
        Q, R = qr_insert(Q2, R2, hcur, j, which='col', overwrite_qru=True, check_finite=False)
        res = abs(Q[0, -1])
        if res < atol or breakdown:
            break
    if not numpy.isfinite(R[j, j]):
        raise numpy.linalg.LinAlgError()
    y = lstsq(R[:j + 1, :j + 1], Q[0, :j + 1].conj())[0]
    return Q, R, vs, zs, y, res


def lgmres(ops, matvec, b, x0=None, rtol=1e-5, atol=0.0, maxiter=1000, M=None, callback=None, inner_m=30, outer_k=3,
           outer_v=None, store_outer_Av=True, prepend_outer_v=False):
    V = _V(ops)
    psolve = M if M is not None else (lambda v: v)
    if outer_v is None:
        outer_v = []
    b_norm = V.nrm(b)
    if not math.isfinite(b_norm):
        raise ValueError('RHS must contain only finite numbers')
    atol_ = _tolerances('lgmres', b_norm, atol, rtol)
    if b_norm == 0:
        return b, 0
    x = V.copy(x0) if x0 is not None else V.o.axpby(V.o.new(b), 0.0, b)
    ptol_max_factor = 1.0
    for k_outer in range(maxiter):
        r_outer = V.comb(1.0, matvec(x), -1.0, b)
        if callback is not None:
            callback(x)
        r_norm = V.nrm(r_outer)
        if r_norm <= max(atol_, rtol * b_norm):
            break
        v0 = V.o.axpby(V.o.new(b), -1.0, psolve(r_outer))
        inner_res_0 = V.nrm(v0)
        if inner_res_0 == 0:
            raise RuntimeError('Preconditioner returned a zero vector; |v| ~ %.1g, |M v| = 0' % r_norm)
        V.scal(v0, 1.0 / inner_res_0)
        ptol = min(ptol_max_factor, max(atol_, rtol * b_norm) / r_norm)
        try:
            Q, R, vs, zs, y, pres = _fgmres(V, matvec, v0, inner_m, ptol, psolve, outer_v, prepend_outer_v)
            y = y * inner_res_0
            if not numpy.isfinite(y).all():
                raise numpy.linalg.LinAlgError()
        except numpy.linalg.LinAlgError:
            return x, k_outer + 1
        if pres > ptol:
            ptol_max_factor = min(1.0, 1.5 * ptol_max_factor)
        else:
            ptol_max_factor = max(1e-16, 0.25 * ptol_max_factor)
        dx = V.o.axpby(V.o.new(b), complex(y[0]), zs[0])
        for w, yc in zip(zs[1:], y[1:]):
            V.axpy(dx, complex(yc), w)
        nx = V.nrm(dx)
        if nx > 0:
            if store_outer_Av:
                q = Q.dot(R.dot(y))
                ax = V.o.axpby(V.o.new(b), complex(q[0]), vs[0])
                for v, qc in zip(vs[1:], q[1:]):
                    V.axpy(ax, complex(qc), v)
                outer_v.append((V.o.axpby(V.o.new(b), 1.0 / nx, dx), V.scal(ax, 1.0 / nx)))
            else:
                outer_v.append((V.o.axpby(V.o.new(b), 1.0 / nx, dx), None))
        while len(outer_v) > outer_k:
            del outer_v[0]
        V.axpy(x, 1.0, dx)
    else:
        return x, maxiter
    return x, 0


def lsqr(ops, matvec, rmatvec, b, n_like, damp=0.0, atol=1e-6, btol=1e-6, conlim=1e8, iter_lim=None, x0=None):
    """Paige & Saunders' LSQR (Golub-Kahan bidiagonalisation); `n_like`: any vector of the solution's size.
    Returns (x, istop, itn, r1norm, r2norm, anorm, acond, arnorm, xnorm)."""
    V = _V(ops)
    sqrt = math.sqrt
    n = n_like.numel()
    if iter_lim is None:
        iter_lim = 2 * n
    itn = istop = 0
    ctol = 1 / conlim if conlim > 0 else 0
    anorm = acond = 0.0
    dampsq = damp ** 2
    ddnorm = res2 = xnorm = xxnorm = z = 0.0
    cs2, sn2 = -1.0, 0.0
    bnorm = V.nrm(b)
    if x0 is None:
        x = V.o.axpby(V.o.new(n_like), 0.0, n_like)
        u = V.copy(b)
        beta = bnorm
    else:
        x = V.copy(x0)
        u = V.comb(1.0, b, -1.0, matvec(x))
        beta = V.nrm(u)
    if beta > 0:
        V.scal(u, 1 / beta)
        v = V.copy(rmatvec(u))
        alfa = V.nrm(v)
    else:
        v = V.copy(x)
        alfa = 0.0
    if alfa > 0:
        V.scal(v, 1 / alfa)
    w = V.copy(v)
    rhobar, phibar = alfa, beta
    rnorm = r1norm = r2norm = beta
    arnorm = alfa * beta
    if arnorm == 0:
        return x, istop, itn, r1norm, r2norm, anorm, acond, arnorm, xnorm
    while itn < iter_lim:
        itn += 1
        V.o.axpby(u, 1.0, matvec(v), -alfa, u)              # u = A v - alfa u
        beta = V.nrm(u)
        if beta > 0:
            V.scal(u, 1 / beta)
            anorm = sqrt(anorm ** 2 + alfa ** 2 + beta ** 2 + dampsq)
            V.o.axpby(v, 1.0, rmatvec(u), -beta, v)          # v = A' u - beta v
            alfa = V.nrm(v)
            if alfa > 0:
                V.scal(v, 1 / alfa)
        if damp > 0:
            rhobar1 = sqrt(rhobar ** 2 + dampsq)
            cs1, sn1 = rhobar / rhobar1, damp / rhobar1
            psi = sn1 * phibar
            phibar = cs1 * phibar
        else:
            rhobar1, psi = rhobar, 0.0
        cs, sn, rho = _sym_ortho(rhobar1, beta)
        theta = sn * alfa
        rhobar = -cs * alfa
        phi = cs * phibar
        phibar = sn * phibar
        tau = sn * phi
        t1 = phi / rho
        t2 = -theta / rho
        ddnorm = ddnorm + (V.nrm(w) / rho) ** 2              # |dk|^2, dk = w / rho
        V.axpy(x, t1, w)
        V.xpay(w, v, t2)
        delta = sn2 * rho
        gambar = -cs2 * rho
        rhs = phi - delta * z
        zbar = rhs / gambar
        xnorm = sqrt(xxnorm + zbar ** 2)
        gamma = sqrt(gambar ** 2 + theta ** 2)
        cs2 = gambar / gamma
        sn2 = theta / gamma
        z = rhs / gamma
        xxnorm = xxnorm + z ** 2
        acond = anorm * sqrt(ddnorm)
        res1 = phibar ** 2
        res2 = res2 + psi ** 2
        rnorm = sqrt(res1 + res2)
        arnorm = alfa * abs(tau)
        if damp > 0:
            r1sq = rnorm ** 2 - dampsq * xxnorm
            r1norm = sqrt(abs(r1sq))
            if r1sq < 0:
                r1norm = -r1norm
        else:
            r1norm = rnorm
        r2norm = rnorm
        test1 = rnorm / bnorm
        test2 = arnorm / (anorm * rnorm + EPS64)
        test3 = 1 / (acond + EPS64)
        t1 = test1 / (1 + anorm * xnorm / bnorm)
        rtol = btol + atol * anorm * xnorm / bnorm
        if itn >= iter_lim:
            istop = 7
        if 1 + test3 <= 1:
            istop = 6
        if 1 + test2 <= 1:
            istop = 5
        if 1 + t1 <= 1:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
        if istop != 0:
            break
    return x, istop, itn, r1norm, r2norm, anorm, acond, arnorm, xnorm


def lsmr(ops, matvec, rmatvec, b, n_like, damp=0.0, atol=1e-6, btol=1e-6, conlim=1e8, maxiter=None, x0=None):
    """Fong & Saunders' LSMR.  Returns (x, istop, itn, normr, normar, normA, condA, normx)."""
    V = _V(ops)
    sqrt = math.sqrt
    m, n = b.numel(), n_like.numel()
    if maxiter is None:
        maxiter = min(m, n)
    normb = V.nrm(b)
    if x0 is None:
        x = V.o.axpby(V.o.new(n_like), 0.0, n_like)
        u = V.copy(b)
        beta = normb
    else:
        x = V.copy(x0)
        u = V.comb(1.0, b, -1.0, matvec(x))
        beta = V.nrm(u)
    if beta > 0:
        V.scal(u, 1 / beta)
        v = V.copy(rmatvec(u))
        alpha = V.nrm(v)
    else:
        v = V.o.axpby(V.o.new(n_like), 0.0, n_like)
        alpha = 0.0
    if alpha > 0:
        V.scal(v, 1 / alpha)
    itn = 0
    zetabar = alpha * beta
    alphabar = alpha
    rho = rhobar = cbar = 1.0
    sbar = 0.0
    h = V.copy(v)
    hbar = V.o.axpby(V.o.new(n_like), 0.0, n_like)
    betadd = beta
    betad = 0.0
    rhodold = 1.0
    tautildeold = thetatilde = zeta = d = 0.0
    normA2 = alpha * alpha
    maxrbar = 0.0
    minrbar = 1e+100
    normA = sqrt(normA2)
    condA = 1.0
    normx = 0.0
    istop = 0
    ctol = 1 / conlim if conlim > 0 else 0
    normr = beta
    normar = alpha * beta
    if normar == 0:
        return x, istop, itn, normr, normar, normA, condA, normx
    if normb == 0:
        V.scal(x, 0.0)
        return x, istop, itn, normr, normar, normA, condA, normx
    while itn < maxiter:
        itn += 1
        V.o.axpby(u, 1.0, matvec(v), -alpha, u)
        beta = V.nrm(u)
        if beta > 0:
            V.scal(u, 1 / beta)
            V.o.axpby(v, 1.0, rmatvec(u), -beta, v)
            alpha = V.nrm(v)
            if alpha > 0:
                V.scal(v, 1 / alpha)
        chat, shat, alphahat = _sym_ortho(alphabar, damp)
        rhoold = rho
        c, s, rho = _sym_ortho(alphahat, beta)
        thetanew = s * alpha
        alphabar = c * alpha
        rhobarold = rhobar
        zetaold = zeta
        thetabar = sbar * rho
        rhotemp = cbar * rho
        cbar, sbar, rhobar = _sym_ortho(cbar * rho, thetanew)
        zeta = cbar * zetabar
        zetabar = -sbar * zetabar
        V.xpay(hbar, h, -(thetabar * rho / (rhoold * rhobarold)))
        V.axpy(x, zeta / (rho * rhobar), hbar)
        V.xpay(h, v, -(thetanew / rho))
        betaacute = chat * betadd
        betacheck = -shat * betadd
        betahat = c * betaacute
        betadd = -s * betaacute
        thetatildeold = thetatilde
        ctildeold, stildeold, rhotildeold = _sym_ortho(rhodold, thetabar)
        thetatilde = stildeold * rhobar
        rhodold = ctildeold * rhobar
        betad = -stildeold * betad + ctildeold * betahat
        tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold
        taud = (zeta - thetatilde * tautildeold) / rhodold
        d = d + betacheck * betacheck
        normr = sqrt(d + (betad - taud) ** 2 + betadd * betadd)
        normA2 = normA2 + beta * beta
        normA = sqrt(normA2)
        normA2 = normA2 + alpha * alpha
        maxrbar = max(maxrbar, rhobarold)
        if itn > 1:
            minrbar = min(minrbar, rhobarold)
        condA = max(maxrbar, rhotemp) / min(minrbar, rhotemp)
        normar = abs(zetabar)
        normx = V.nrm(x)
        test1 = normr / normb
        test2 = normar / (normA * normr) if (normA * normr) != 0 else float('inf')
        test3 = 1 / condA
        t1 = test1 / (1 + normA * normx / normb)
        rtol = btol + atol * normA * normx / normb
        if itn >= maxiter:
            istop = 7
        if 1 + test3 <= 1:
            istop = 6
        if 1 + test2 <= 1:
            istop = 5
        if 1 + t1 <= 1:
            istop = 4
        if test3 <= ctol:
            istop = 3
        if test2 <= atol:
            istop = 2
        if test1 <= rtol:
            istop = 1
        if istop > 0:
            break
    return x, istop, itn, normr, normar, normA, condA, normx
