import scipy.signal
from oracle import gss_oracle as _o


def stft(time_signal, size, shift, axis=-1, window=scipy.signal.windows.blackman,
         window_length=None, fading=True, pad=True, symmetric_window=False):
    assert axis == -1 and window_length is None and not symmetric_window
    return _o.stft(time_signal, size, shift, fading, window, pad)


def istft(stft_signal, size=1024, shift=256, window=scipy.signal.windows.blackman,
          fading=True, window_length=None, symmetric_window=False):
    assert window_length is None and not symmetric_window
    return _o.istft(stft_signal, size, shift, fading, window)


def _samples_to_stft_frames(samples, size, shift, *, pad=True, fading=False):
    return _o.samples_to_stft_frames(samples, size, shift, pad=pad, fading=fading)
