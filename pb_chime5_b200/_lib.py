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

GSS_ERR_ARG = -1
GSS_ERR_UNSUPPORTED = -2
GSS_ERR_WORKSPACE = -3
GSS_ERR_CUDA = -4

INFO_NOT_POSDEF = 1
INFO_NONFINITE = 2
INFO_NO_CONVERGE = 3
INFO_SINGULAR = 4

OP_WEIGHTED_COV, OP_CACGMM, OP_BEAMFORM, OP_WPE, OP_STFT, OP_ISTFT, OP_ENHANCE = range(7)

BF_TYPES = {'mvdrSouden_ban': 0, 'gev_ban': 1, 'ch': 2, 'sum': 3, 'mvdrSouden': 4, 'gev': 5}
POSTFILTERS = {None: 0, 'mask_mul': 1}

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
    'gss_wpe_c64': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    'gss_enhance_c64': (_i, [_p] * 8 + [_i] * 15 + [_p, _p, _sz, _p]),
    'gss_stft_f32': (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _sz, _p]),
    'gss_istft_f32': (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _sz, _p]),
    'gss_debug_wpe_config': (_i, [_i, _d]),
    'gss_debug_wpe_redo_count': (_i, [_i]),
    'gss_debug_mstep_i8': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    'gss_debug_wpe_gram': (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _sz, _p]),
}

_lib = None


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


def last_error():
    return lib().gss_last_error().decode('utf-8', 'replace')


def check(rc):
    """Map a C return code to the exception type the reference would raise."""
    if rc == 0:
        return
    msg = last_error()
    if rc == GSS_ERR_ARG:
        raise AssertionError(msg)
    if rc == GSS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f'libgss error {rc}: {msg}')


def workspace_bytes(op, B=1, F=1, D=1, T=1, K=1, L=1):
    out = _sz(0)
    check(lib().gss_workspace_bytes(op, B, F, D, T, K, L, C.byref(out)))
    return int(out.value)
