"""CPU-only: the product's host plan constants (pynufft_b200/planmath.py) against the reference's
values stored in the golden fixtures and against the oracle."""
import numpy

from oracle import nufft_oracle as orc
from pynufft_b200 import planmath as pm


def test_alpha_and_sn_match_reference(golden):
    Nd, Kd, Jd = golden['Nd'], golden['Kd'], golden['Jd']
    alphas, sns = [], []
    for d in range(len(Nd)):
        a, b = pm.kb_fit_alpha(Nd[d], Jd[d], Kd[d])
        assert a.dtype == numpy.float32 and b == 1
        alphas.append(a)
        sns.append(numpy.real(pm.scaling_vector(Nd[d], Kd[d], a, b)))
    assert numpy.allclose(numpy.concatenate(alphas), golden['alpha'].reshape(-1), rtol=1e-5, atol=1e-7)
    assert numpy.allclose(numpy.concatenate(sns).astype(numpy.float32), golden['tensor_sn'], rtol=1e-6)


def test_T_matches_oracle():
    for (N, J, K) in [(256, 6, 512), (24, 4, 40), (8, 4, 16), (30, 7, 60), (64, 1, 128)]:
        a, b = pm.kb_fit_alpha(N, J, K)
        assert numpy.allclose(pm.interp_T(N, J, K, a, b), orc.T_matrix(N, J, K, a, b), rtol=1e-12, atol=1e-14)
