"""ctypes loader for libgss.so (the C ABI declared in include/gss.h).

The library is the product: if it is missing or fails to load this module
raises -- there is no CPU fallback anywhere in ``pb_chime5_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get('GSS_LIB', _HERE / 'csrc' / 'libgss.so'))
# developer / measurement API (include/gss_dev.h): a superset of the product library, loaded only by
# tests/, tools/ and the side measurements of bench.py -- never by the product modules
DEV_LIB_PATH = Path(os.environ.get('GSS_DEV_LIB', _HERE / 'csrc' / 'libgss_dev.so'))

GSS_ERR_ARG = -1
GSS_ERR_UNSUPPORTED = -2
GSS_ERR_WORKSPACE = -3
GSS_ERR_CUDA = -4

INFO_NOT_POSDEF = 1
INFO_NONFINITE = 2
INFO_NO_CONVERGE = 3
INFO_SINGULAR = 4

(OP_WEIGHTED_COV, OP_CACGMM, OP_BEAMFORM, OP_WPE, OP_STFT, OP_ISTFT, OP_ENHANCE, OP_BF_VECTOR,
 OP_CACGMM_C128, OP_ENHANCE_F64) = range(10)

BF_TYPES = {'mvdrSouden_ban': 0, 'gev_ban': 1, 'ch': 2, 'sum': 3, 'mvdrSouden': 4, 'gev': 5}
POSTFILTERS = {None: 0, 'mask_mul': 1}
BF_CORES = {'mvdr_souden': 0, 'gev': 1, 'wmwf': 2, 'pca': 3, 'pca+mvdr': 4, 'scaled_gev_atf+mvdr': 5, 'ch': 6}
BF_RANK1 = {None: 0, 'rank1_pca': 1, 'rank1_gev': 2}
PCA_SCALING = {None: 0, 'trace': 1, 'eigenvalue': 2}

_p = C.c_void_p
_i = C.c_int
_d = C.c_double
_sz = C.c_size_t

_SIGNATURES = {
    'gss_version': (_i, []),
    'gss_last_error': (C.c_char_p, []),
    'gss_launch_count': (C.c_longlong, []),
    'gss_workspace_bytes': (_i, [_i, _i, _i, _i, _i, _i, _i, C.POINTER(_sz)]),
    'gss_pack_dtf_to_fdt_c64': (_i, [_p, _p, _i, _i, _i, _i, _p]),
    'gss_unpack_fdt_to_dtf_c64': (_i, [_p, _p, _i, _i, _i, _i, _p]),
    'gss_unpack_fkt_to_ktf_f32': (_i, [_p, _p, _i, _i, _i, _i, _p]),
    'gss_pack_ktf_to_fkt_f32': (_i, [_p, _p, _i, _i, _i, _i, _p]),
    'gss_unpack_ft_to_tf_c64': (_i, [_p, _p, _i, _i, _i, _p]),
    'gss_weighted_cov_c64': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    'gss_cacgmm_c64': (_i, [_p, _p, _p, _i, _i, _d, _d, _i, _i, _i, _i, _i, _i, _p,
                            _p, _p, _p, _p, _p, _sz, _p]),
    'gss_beamform_c64': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p,
                              _p, _p, _p, _p, _sz, _p]),
    'gss_beamform_from_posterior_c64': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i,
                                             _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _sz, _p]),
    'gss_bf_vector_c128': (_i, [_p, _p, _p, _i, _i, _i, _i, _d, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    'gss_wpe_c64': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    'gss_wpe_c64_ex': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _d, _p, _p, _p, _p, _sz, _p]),
    'gss_cacgmm_c128': (_i, [_p, _p, _p, _i, _i, _d, _d, _i, _i, _i, _i, _i, _i, _p,
                             _p, _p, _p, _p, _p, _sz, _p]),
    'gss_enhance_c64_ex': (_i, [_p] * 8 + [_i] * 16 + [_p, _p, _sz, _p]),
    'gss_enhance_c64': (_i, [_p] * 8 + [_i] * 15 + [_p, _p, _sz, _p]),
    'gss_stft_f32': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    'gss_istft_f32': (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
}

_DEV_SIGNATURES = {
    'gss_debug_mstep_i8': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    'gss_debug_wpe_gram': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    'gss_debug_fp64_peak_scratch_bytes': (_sz, []),
    'gss_debug_fp64_peak': (_i, [_i, _i, _p, C.POINTER(_d), _p]),
}

WPE_GRAM_AUTO, WPE_GRAM_F64, WPE_GRAM_I8, WPE_GRAM_I8_REDO = -1, 0, 1, 2
WPE_GRAM_MODES = {None: WPE_GRAM_AUTO, 'auto': WPE_GRAM_AUTO, 'f64': WPE_GRAM_F64, 'i8': WPE_GRAM_I8,
                  'i8+redo': WPE_GRAM_I8_REDO}

_lib = None
_dev_lib = None


def exported_symbols():
    """Names include/gss.h declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f'{LIB_PATH} not found: build it with `python __graft_entry__.py` '
                '(or pb_chime5_b200/csrc/build.sh). pb_chime5_b200 has no CPU fallback.')
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def dev_exported_symbols():
    """Names include/gss_dev.h declares (libgss_dev.so only)."""
    return sorted(_DEV_SIGNATURES)


def dev_lib():
    """libgss_dev.so: the product ABI plus the gss_debug_* developer entry points."""
    global _dev_lib
    if _dev_lib is None:
        if not DEV_LIB_PATH.exists():
            raise RuntimeError(f'{DEV_LIB_PATH} not found: build it with pb_chime5_b200/csrc/build.sh')
        handle = C.CDLL(str(DEV_LIB_PATH))
        for name, (res, args) in {**_SIGNATURES, **_DEV_SIGNATURES}.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _dev_lib = handle
    return _dev_lib


def last_error(handle=None):
    return (handle or lib()).gss_last_error().decode('utf-8', 'replace')


def check(rc, handle=None):
    """Map a C return code to the exception type the reference would raise.
    handle: the library the call went to (default: the product library)."""
    if rc == 0:
        return
    msg = last_error(handle)
    if rc == GSS_ERR_ARG:
        raise AssertionError(msg)
    if rc == GSS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f'libgss error {rc}: {msg}')


def workspace_bytes(op, B=1, F=1, D=1, T=1, K=1, L=1):
    out = _sz(0)
    check(lib().gss_workspace_bytes(op, B, F, D, T, K, L, C.byref(out)))
    return int(out.value)
