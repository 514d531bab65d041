"""ctypes binding of libb200nufft.so (include/b200nufft.h).  No fallback: if the library is
missing and cannot be built, importing the operator fails loudly."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, 'libb200nufft.so')

MAX_DIM, MAX_J, MAX_L = 3, 16, 32

_c = ctypes
_vp, _i, _i64, _f, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_double

# name -> (restype, argtypes); mirrors include/b200nufft.h one to one
SIGNATURES = {
    'b200nufft_last_error': (_c.c_char_p, []),
    'b200nufft_version': (_i, []),
    'b200nufft_plan_create': (_i, [_c.POINTER(_vp), _i, _i, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    'b200nufft_plan_destroy': (_i, [_vp]),
    'b200nufft_plan_get_kindx': (_i, [_vp, _vp, _vp]),
    'b200nufft_plan_get_udata': (_i, [_vp, _vp, _vp]),
    'b200nufft_plan_get_k0': (_i, [_vp, _vp, _vp]),
    'b200nufft_plan_get_perm': (_i, [_vp, _vp, _vp]),
    'b200nufft_plan_get_tile': (_i, [_vp, _vp]),
    'b200nufft_plan_bytes': (_i64, [_vp]),
    'b200nufft_x2xx': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    'b200nufft_scale_pad': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'b200nufft_fft': (_i, [_vp, _vp, _i, _i, _vp]),
    'b200nufft_interp': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_gridding': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_crop_scale': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'b200nufft_pad_fft': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'b200nufft_ifft_crop': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'b200nufft_forward': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_adjoint': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_forward_one2many': (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    'b200nufft_adjoint_many2one': (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    'b200nufft_forward_host': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_adjoint_host': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_dotc': (_i, [_vp, _vp, _i64, _vp, _vp]),
    'b200nufft_zero_scalars': (_i, [_vp, _i, _vp]),
    'b200nufft_cg_update_xr': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    'b200nufft_cg_update_p': (_i, [_vp, _vp, _vp, _vp, _i64, _vp]),
    'b200nufft_axpby': (_i, [_vp, _d, _d, _vp, _d, _d, _vp, _i64, _vp]),
    'b200nufft_cg_init': (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    'b200nufft_cdiv': (_i, [_vp, _vp, _i64, _vp]),
    'b200nufft_cmul': (_i, [_vp, _vp, _i64, _vp]),
    'b200nufft_tv_rhs': (_i, [_vp, _vp, _vp, _vp, _f, _f, _vp, _vp]),
    'b200nufft_tv_shrink': (_i, [_vp, _vp, _vp, _vp, _f, _vp]),
    'b200nufft_tv_bregman': (_i, [_vp, _vp, _vp, _i64, _vp]),
    'b200nufft_forward_host_async': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    'b200nufft_adjoint_host_async': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    'b200nufft_host_wait': (_i, [_vp, _i, _i]),
    'b200nufft_set_variant': (_i, [_vp, _i, _i]),
    'b200nufft_set_layout_preference': (_i, [_i]),
    'b200nufft_plan_get_layout': (_i, [_vp]),
    'b200nufft_grid_layout': (_i, [_vp, _i]),
    'b200nufft_plan_get_col_perm': (_i, [_vp, _vp, _vp, _vp]),
    'b200nufft_gridding_is_modulated': (_i, [_vp]),
    'b200nufft_kspace_modulated': (_i, [_vp]),
    'b200nufft_kspace_modulated_nb': (_i, [_vp, _i]),
    'b200nufft_interp_modulated': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_gridding_modulated': (_i, [_vp, _vp, _vp, _i, _vp]),
    'b200nufft_ifft_crop_modulated': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'b200nufft_pad_fft_modulated': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    'b200nufft_launch_count': (_i64, []),
}

_lib = None


class B200NufftError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load the C-ABI library, (re)building it first when it is missing or stale (the source hash stored beside the
    library differs from the sources in the tree); raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    from . import _build
    if not _build.is_current():          # missing, or built from other sources than the ones in the tree (source hash)
        if not build_if_missing:
            raise B200NufftError('libb200nufft.so is missing or stale: run `python -m pynufft_b200._build`')
        _build.build()
    lib = ctypes.CDLL(LIBPATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().b200nufft_last_error()
        msg = msg.decode() if msg else 'unknown error'
        if rc == -1:
            raise ValueError(msg)
        raise B200NufftError('libb200nufft error %d: %s' % (rc, msg))
