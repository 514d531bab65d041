"""
Iterates of the device CG in its own complex64 arithmetic AND in exact (complex128) arithmetic (linalg/solve_device.py:351-481, batched twin
linalg/solve_hsa.py:551-682) at configuration 4 (2-D 256^2 / 512^2 / 6^2, golden-angle radial 402 x 512, 32 coils):
oracle.solve_cg(...) with dtype complex64 and complex128 after 10 and after 100 iterations, stored at 60 000 random
entries of the Nd + (B,) result (the full array is 16.8 MB) -> tests/golden/c4_cg_c128.npz.

Why a fixture: the 100-iteration runs take ~5 minutes of CPU, too long for the GPU test suite.  The system is
ill-conditioned: storing the CG vectors in complex64 moves the iterates away from the exact ones by 2.3e-1 after 10 and
1.5e-2 after 100 iterations -- deterministically (the CUDA solver and the numpy restatement round at the same places and
land on the same perturbed trajectory), so parity = agreement with the complex64 restatement; the complex128 iterates
are kept to show how far both are from exact arithmetic.
Test infrastructure only (uses the oracle, which is pinned to the reference goldens).

    python oracle/make_cg_fixture.py
"""
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
from oracle import nufft_oracle as orc  # noqa: E402


def golden_angle_radial(nspokes=402, nread=512):
    s = numpy.arange(nspokes)
    th = s * numpy.pi * (numpy.sqrt(5.0) - 1.0) / 2.0
    r = numpy.pi * (numpy.arange(nread) - nread / 2) / (nread / 2)
    return numpy.stack([numpy.outer(numpy.cos(th), r), numpy.outer(numpy.sin(th), r)], -1).reshape(-1, 2)


def coil_maps(Nd, B, seed=0):
    rng = numpy.random.default_rng(seed)
    grids = numpy.meshgrid(*[numpy.arange(n) for n in Nd], indexing='ij')
    maps = numpy.zeros(Nd + (B,), dtype=numpy.complex64)
    for c in range(B):
        ang = 2 * numpy.pi * c / B
        ctr = [Nd[0] / 2 + 0.45 * Nd[0] * numpy.cos(ang), Nd[1] / 2 + 0.45 * Nd[1] * numpy.sin(ang)]
        r2 = sum((g - c0) ** 2 for g, c0 in zip(grids, ctr))
        maps[..., c] = numpy.exp(-r2 / (2 * (0.6 * Nd[0]) ** 2)) * numpy.exp(1j * rng.uniform(0, 2 * numpy.pi))
    return maps


def main():
    Nd, Kd, Jd, B = (256, 256), (512, 512), (6, 6), 32
    O = orc.NUFFT()
    O.plan(golden_angle_radial(), Nd, Kd, Jd, batch=B)
    O.set_sense(coil_maps(Nd, B))
    rng = numpy.random.default_rng(4)
    s = (rng.standard_normal(Nd) + 1j * rng.standard_normal(Nd)).astype(numpy.complex64)
    y = O.forward_one2many(s).astype(numpy.complex64)
    pick = numpy.sort(numpy.random.default_rng(5).choice(int(numpy.prod(Nd)) * B, 60000, replace=False))
    path = os.path.join(HERE, '..', 'tests', 'golden', 'c4_cg_c128.npz')
    out = dict(pick=pick.astype(numpy.int32))
    if 'c64only' in sys.argv[1:]:                 # keep the (3-minute) complex128 entries of an existing fixture
        out = dict(numpy.load(path))
        assert numpy.array_equal(out['pick'], pick)
    for it in (10, 100):
        if 'c64only' not in sys.argv[1:]:
            x = orc.solve_cg(O, y, it, dtype=numpy.complex128)
            out['x%d' % it] = x.ravel()[pick].astype(numpy.complex64)
            out['norm%d' % it] = numpy.array(numpy.linalg.norm(x))
            print(it, 'iterations (complex128) done, |x| = %.6e' % numpy.linalg.norm(x), flush=True)
        # the complex64 restatement of the device arithmetic (what the reference's device solver computes)
        x = orc.solve_cg(O, y, it)
        out['x%d_c64' % it] = x.ravel()[pick].astype(numpy.complex64)
        out['norm%d_c64' % it] = numpy.array(numpy.linalg.norm(x))
        print(it, 'iterations (complex64) done, |x| = %.6e' % numpy.linalg.norm(x), flush=True)
    numpy.savez_compressed(path, **out)
    print('%.1f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
