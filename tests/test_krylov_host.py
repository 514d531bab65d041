"""The device Krylov recurrences (pynufft_b200/krylov.py) driven on CPU tensors against scipy itself: same stopping
rules, same iterates.  scipy is what the reference's CPU solve calls (linalg/solve_cpu.py:226-288); the matrices are
dense complex64, as the reference's operator is."""
import numpy
import pytest
import scipy.sparse.linalg as sla
import torch

from pynufft_b200 import krylov as kr

c64 = numpy.complex64


def _square(n, seed, herm=False):
    rng = numpy.random.default_rng(seed)
    B = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / numpy.sqrt(n)
    A = numpy.eye(n) + 0.35 * B
    if herm:
        A = A.conj().T @ A
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    return A.astype(c64), b.astype(c64)


def _ops(A):
    At = torch.from_numpy(A)
    AH = torch.from_numpy(numpy.ascontiguousarray(A.conj().T))
    return kr.TorchVectorOps(), (lambda v: At @ v), (lambda v: AH @ v)


def rel(a, b):
    a, b = numpy.asarray(a), numpy.asarray(b)
    return numpy.linalg.norm(a - b) / numpy.linalg.norm(b)


@pytest.mark.parametrize('maxiter', [1, 3, 8, 200])
def test_bicgstab(maxiter):
    A, b = _square(60, 1)
    ops, mv, _ = _ops(A)
    x, info = kr.bicgstab(ops, mv, torch.from_numpy(b), maxiter=maxiter)
    xs, infos = sla.bicgstab(A, b, maxiter=maxiter)
    assert info == infos
    assert rel(x.numpy(), xs) < 2e-4


@pytest.mark.parametrize('maxiter', [1, 4, 200])
def test_bicg(maxiter):
    A, b = _square(60, 2)
    ops, mv, rmv = _ops(A)
    x, info = kr.bicg(ops, mv, rmv, torch.from_numpy(b), maxiter=maxiter)
    xs, infos = sla.bicg(A, b, maxiter=maxiter)
    assert info == infos
    assert rel(x.numpy(), xs) < 2e-4


@pytest.mark.parametrize('kw', [dict(maxiter=1, restart=4), dict(maxiter=7, restart=4), dict(maxiter=3), dict(),
                                dict(maxiter=2, restart=5, callback_type='x'), dict(rtol=1e-3)])
def test_gmres(kw):
    A, b = _square(50, 3)
    ops, mv, _ = _ops(A)
    kw = dict(kw)
    if 'callback_type' in kw:
        kw['callback'] = lambda x: None
    x, info = kr.gmres(ops, mv, torch.from_numpy(b), **kw)
    xs, infos = sla.gmres(A, b, **kw)
    assert info == infos
    assert rel(x.numpy(), xs) < 2e-4


@pytest.mark.parametrize('kw', [dict(maxiter=1, inner_m=4), dict(maxiter=3, inner_m=4, outer_k=2), dict(maxiter=50),
                                dict(maxiter=4, inner_m=3, prepend_outer_v=True), dict(maxiter=3, inner_m=5, store_outer_Av=False)])
def test_lgmres(kw):
    A, b = _square(50, 4)
    ops, mv, _ = _ops(A)
    x, info = kr.lgmres(ops, mv, torch.from_numpy(b), **kw)
    xs, infos = sla.lgmres(A, b, **kw)
    assert info == infos
    assert rel(x.numpy(), xs) < 2e-4


def _rect(m, n, seed):
    rng = numpy.random.default_rng(seed)
    A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / numpy.sqrt(m)
    b = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    return A.astype(c64), b.astype(c64)


@pytest.mark.parametrize('kw', [dict(iter_lim=1), dict(iter_lim=6), dict(), dict(damp=0.1, iter_lim=9), dict(atol=1e-3, btol=1e-3)])
def test_lsqr(kw):
    A, b = _rect(80, 50, 5)
    ops, mv, rmv = _ops(A)
    out = kr.lsqr(ops, mv, rmv, torch.from_numpy(b), torch.empty(50, dtype=torch.complex64), **kw)
    ref = sla.lsqr(A, b, **kw)
    assert out[1] == ref[1] and out[2] == ref[2]                      # istop, itn
    assert rel(out[0].numpy(), ref[0]) < 2e-4
    for mine, theirs in zip(out[3:9], ref[3:9]):                      # r1norm, r2norm, anorm, acond, arnorm, xnorm
        assert abs(mine - theirs) <= 2e-3 * abs(theirs) + 1e-5


@pytest.mark.parametrize('kw', [dict(maxiter=1), dict(maxiter=6), dict(), dict(damp=0.1, maxiter=9), dict(atol=1e-3, btol=1e-3)])
def test_lsmr(kw):
    A, b = _rect(80, 50, 6)
    ops, mv, rmv = _ops(A)
    out = kr.lsmr(ops, mv, rmv, torch.from_numpy(b), torch.empty(50, dtype=torch.complex64), **kw)
    ref = sla.lsmr(A, b, **kw)
    assert out[1] == ref[1] and out[2] == ref[2]
    assert rel(out[0].numpy(), ref[0]) < 2e-4
    # normr, normar, normA, normx (condA: scipy's float32 run overflows its 1e100 start value of minrbar)
    for mine, theirs in zip(out[3:6] + out[7:8], ref[3:6] + ref[7:8]):
        assert abs(mine - theirs) <= 2e-3 * abs(theirs) + 1e-5


def test_x0_and_zero_rhs():
    A, b = _square(40, 7)
    ops, mv, rmv = _ops(A)
    x0 = numpy.linspace(0, 1, 40).astype(c64)
    x, info = kr.bicgstab(ops, mv, torch.from_numpy(b), x0=torch.from_numpy(x0), maxiter=2)
    xs, infos = sla.bicgstab(A, b, x0=x0, maxiter=2)
    assert info == infos and rel(x.numpy(), xs) < 2e-4
    x, info = kr.gmres(ops, mv, torch.from_numpy(b), x0=torch.from_numpy(x0), maxiter=3, restart=3)
    xs, infos = sla.gmres(A, b, x0=x0, maxiter=3, restart=3)
    assert info == infos and rel(x.numpy(), xs) < 2e-4
    z = torch.zeros(40, dtype=torch.complex64)
    for fn in (kr.bicgstab, kr.gmres, kr.lgmres):
        x, info = fn(ops, mv, z)
        assert info == 0 and float(x.abs().max()) == 0.0
    with pytest.raises(ValueError):
        kr.bicgstab(ops, mv, torch.from_numpy(b), atol=-1.0)
