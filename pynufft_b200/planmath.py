"""
Host part of plan(): the per-dimension constants of the Fessler-Sutton min-max interpolator.

These are O(N_d + J^2 L^2) numbers per dimension, computed once on the host in numpy exactly as
the reference does (the per-sample O(M) part runs on the GPU, csrc/plan.cu):

  kb_fit_alpha   <- helper.nufft_alpha_kb_fit  (src/_helper/helper.py:910-958), incl. kaiser_bessel
                    (:961-997) and kaiser_bessel_ft (:1000-1012)
  scaling_vector <- helper.nufft_scale1        (:1015-1036), real part (:274, :780)
  interp_T       <- helper.nufft_T             (:1060-1083)

dtype quirks that are part of the behaviour (and therefore kept): the least-squares fit runs in
complex64, so alpha is float32; products alpha_i*alpha_j in T are rounded to float32.
"""
import numpy
import scipy.linalg
import scipy.special

_BEST_ZN = {2: 2.5, 3: 2.27, 4: 2.31, 5: 2.34, 6: 2.32, 7: 2.32, 8: 2.35, 9: 2.34, 10: 2.34,
            11: 2.35, 12: 2.34, 13: 2.35, 14: 2.35, 15: 2.35, 16: 2.33}


def _kb_shape(J, K_over_N):
    if K_over_N != 2:
        return 2.34 * J
    if J in _BEST_ZN:
        return J * _BEST_ZN[J]
    keys = numpy.array(sorted(_BEST_ZN))
    return J * _BEST_ZN[int(keys[numpy.argmin(numpy.abs(keys - J))])]


def _kb_ft(u, J, kb_a):
    """Fourier transform of the Kaiser-Bessel window, order 0, 1-D."""
    z = numpy.sqrt((2 * numpy.pi * (J / 2) * (u + 0j)) ** 2.0 - kb_a ** 2.0)
    y = numpy.sqrt(2 * numpy.pi) * (J / 2) / scipy.special.iv(0, kb_a) * scipy.special.jv(0.5, z) / numpy.sqrt(z)
    return numpy.real(y)


def kb_fit_alpha(N, J, K):
    """alpha (float32, length L+1) and beta=1 of the cosine-series fit to 1/FT(KB)."""
    beta = 1
    L = 13 if N > 40 else int(numpy.ceil(N / 3))
    n0 = numpy.arange(N) * 1.0 - (N - 1.0) / 2.0
    if J > 1:
        target = 1 / _kb_ft(n0 / K, J, _kb_shape(J, K / N))
    else:
        target = numpy.ones(N)
    design = numpy.cos(numpy.outer(beta * (2 * numpy.pi / K) * n0, numpy.arange(L + 1)))
    design = numpy.nan_to_num(design.astype(numpy.complex64))
    target = numpy.nan_to_num(target.reshape(N, 1).astype(numpy.complex64))
    coef = numpy.linalg.lstsq(design, target, rcond=-1)[0]
    if J > 1:
        coef[1:] = coef[1:] / 2.0
    else:
        coef[0] = 1.0
        coef[1:] = 0.0
    return numpy.real(coef).reshape(-1), beta


def scaling_vector(N, K, alpha, beta=1):
    """sn[n] = sum_{l=-L..L} alpha_|l| exp(i gam (n - (N-1)/2) beta l); complex128, real part is used."""
    L = len(alpha) - 1
    if L == 0:
        return alpha[0] * numpy.ones(N)
    phase = 1j * (2 * numpy.pi / K) * (numpy.arange(N) - (N - 1) / 2.0) * beta
    sn = numpy.zeros(N)
    for l in range(-L, L + 1):
        sn = sn + alpha[abs(l)] * numpy.exp(phase * l)
    return sn


def interp_T(N, J, K, alpha, beta=1):
    """T = pinv(CSSC), CSSC[j1,j2] = sum_{l1,l2} alpha_|l1| alpha_|l2| sinc((j2-j1+beta(l1-l2))/(K/N))."""
    L = len(alpha) - 1
    j1, j2 = numpy.mgrid[1:J + 1, 1:J + 1]
    diff = j2 - j1
    cssc = numpy.zeros((J, J))
    for l1 in range(-L, L + 1):
        for l2 in range(-L, L + 1):
            aa = alpha[abs(l1)] * alpha[abs(l2)]           # float32 * float32 -> float32
            cssc = cssc + aa * numpy.sinc((diff + beta * (l1 - l2)) / (1.0 * K / N))
    return scipy.linalg.pinv(cssc)
