import numpy as np
from oracle import gss_oracle as _o


def wpe_v8(Y, taps=10, delay=3, iterations=3, psd_context=0,
           statistics_mode='full', inplace=False):
    assert statistics_mode == 'full', statistics_mode
    Y = np.asarray(Y)
    if Y.ndim == 2:
        return _o.wpe_bins(Y, taps, delay, iterations, psd_context)
    out = np.empty(Y.shape, dtype=np.complex128)
    for index in np.ndindex(Y.shape[:-2]):     # one frequency bin at a time
        out[index] = _o.wpe_bins(Y[index], taps, delay, iterations, psd_context)
    return out
