"""Experiment helper: `import variant_env` before `pynufft_b200` makes the package load variants/lib$B200NUFFT_VARIANT.so
(built by tools/build_variant.sh) instead of the product library.  Tools only; the product loader knows nothing of it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
name = os.environ.get('B200NUFFT_VARIANT')
if name:
    import pynufft_b200._lib as L, pynufft_b200._build as B
    L.LIBPATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'variants', 'lib%s.so' % name)
    B.is_current = lambda: True
    assert os.path.exists(L.LIBPATH), L.LIBPATH
