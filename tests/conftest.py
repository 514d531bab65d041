import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ['ref_2d_64', 'ref_2d_odd', 'ref_3d_16', 'ref_1d_64', 'ref_3d_small']


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    import numpy
    g = dict(numpy.load(os.path.join(GOLDEN, request.param + '.npz')))
    g['name'] = request.param
    g['Nd'] = tuple(int(v) for v in g['Nd'])
    g['Kd'] = tuple(int(v) for v in g['Kd'])
    g['Jd'] = tuple(int(v) for v in g['Jd'])
    return g


def load_golden(name):
    """a golden file by name (the size / multi-coil fixtures are not part of the parametrised `golden` fixture)"""
    import numpy
    g = dict(numpy.load(os.path.join(GOLDEN, name + '.npz')))
    for k in ('Nd', 'Kd', 'Jd'):
        g[k] = tuple(int(v) for v in g[k])
    return g
