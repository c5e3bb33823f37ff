"""Device-level operators: thin, typed wrappers over the C ABI (include/gss.h).

All tensors are CUDA tensors in the bin-major layout of the library
(``Y[b, f, d, t]`` complex64, masks / posteriors float32 with the frame axis
last).  PyTorch is only the container for device memory and streams; every
computation below is a libgss kernel.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, ndim, name):
    assert isinstance(t, torch.Tensor) and t.is_cuda, f'{name} must be a CUDA tensor'
    assert t.dtype == dtype, f'{name}: dtype {t.dtype}, expected {dtype}'
    assert t.ndim == ndim, f'{name}: shape {tuple(t.shape)}, expected {ndim} dims'
    return t.contiguous()


_WS = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer per device (caller-provided `ws` of the C ABI)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        _WS[key] = buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    return buf


def new_info(B, device, stages=1):
    """Zeroed device status words: one int32 per (stage, utterance).  Every stage of a pipeline gets its
    own row, so a word of a later stage can never overwrite an earlier, more severe one."""
    shape = (max(B, 1),) if stages == 1 else (stages, max(B, 1))
    return torch.zeros(shape, dtype=torch.int32, device=device)


def check_info(info, what):
    """Translate device-side status words into the reference's exceptions.  `info`: a tensor (any
    device) or a nested list of words (B,) / (stages, B); `what`: a name or one name per stage.
    Reading a device tensor synchronises -- pipelines therefore pass their own `info` tensors to the
    blocks below and call this once per batch, when the results are fetched."""
    if info is None:
        return
    rows = info.tolist() if isinstance(info, torch.Tensor) else info
    if rows and not isinstance(rows[0], (list, tuple)):
        rows = [rows]
    names = [what] * len(rows) if isinstance(what, str) else list(what)
    singular = []
    for name, vals in zip(names, rows):
        for b, v in enumerate(vals):
            code, f = v & 0xFF, v >> 8
            if code == _lib.INFO_NOT_POSDEF:
                # zhegvd INFO > N (get_gev_vector.pyx:139-147)
                raise ValueError(f'{name}: the noise PSD matrix is not positive definite '
                                 f'for utterance {b}, frequency {f}')
            if code == _lib.INFO_NONFINITE:
                raise AssertionError(f'{name}: non-finite SNR in the reference channel search '
                                     f'(beamformer.py:542), utterance {b}')
            if code == _lib.INFO_NO_CONVERGE:
                raise RuntimeError(f'{name}: eigensolver did not converge, utterance {b}, frequency {f}')
            if code == _lib.INFO_SINGULAR:
                singular.append((name, b, f))
    if singular:
        # nara_wpe falls back to a least-squares solve without telling anybody; the device returns the
        # same minimum-norm solution (dead channels are deflated) -- say so once per batch
        import warnings
        name, b, f = singular[0]
        warnings.warn(f'{name}: singular normal equations (dead channel?) in {len(singular)} utterance(s), '
                      f'first: utterance {b}, frequency {f}; the minimum-norm solution was used',
                      RuntimeWarning, stacklevel=2)


# ---- layout glue ------------------------------------------------------------

def pack_dtf_to_fdt(x):
    """(B,D,T,F) complex64 -> (B,F,D,T)."""
    x = _need(x, torch.complex64, 4, 'x')
    B, D, T, F = x.shape
    out = torch.empty((B, F, D, T), dtype=torch.complex64, device=x.device)
    _lib.check(_lib.lib().gss_pack_dtf_to_fdt_c64(_ptr(x), _ptr(out), B, D, T, F, _stream()))
    return out


def unpack_fdt_to_dtf(x):
    """(B,F,D,T) complex64 -> (B,D,T,F)."""
    x = _need(x, torch.complex64, 4, 'x')
    B, F, D, T = x.shape
    out = torch.empty((B, D, T, F), dtype=torch.complex64, device=x.device)
    _lib.check(_lib.lib().gss_unpack_fdt_to_dtf_c64(_ptr(x), _ptr(out), B, D, T, F, _stream()))
    return out


def unpack_fkt_to_ktf(x):
    """(B,F,K,T) float32 -> (B,K,T,F)."""
    x = _need(x, torch.float32, 4, 'x')
    B, F, K, T = x.shape
    out = torch.empty((B, K, T, F), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().gss_unpack_fkt_to_ktf_f32(_ptr(x), _ptr(out), B, K, T, F, _stream()))
    return out


def pack_ktf_to_fkt(x):
    """(B,K,T,F) float32 -> (B,F,K,T)."""
    x = _need(x, torch.float32, 4, 'x')
    B, K, T, F = x.shape
    out = torch.empty((B, F, K, T), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().gss_pack_ktf_to_fkt_f32(_ptr(x), _ptr(out), B, K, T, F, _stream()))
    return out


def unpack_ft_to_tf(x):
    """(B,F,T) complex64 -> (B,T,F)."""
    x = _need(x, torch.complex64, 3, 'x')
    B, F, T = x.shape
    out = torch.empty((B, T, F), dtype=torch.complex64, device=x.device)
    _lib.check(_lib.lib().gss_unpack_ft_to_tf_c64(_ptr(x), _ptr(out), B, T, F, _stream()))
    return out


# ---- numeric blocks ---------------------------------------------------------

def _tper(frames, B, device):
    """Optional valid-frame counts per utterance -> int32 device tensor (or None)."""
    if frames is None:
        return None
    t = torch.as_tensor(frames, dtype=torch.int32).reshape(-1).to(device)
    assert t.numel() == B, (t.numel(), B)
    return t.contiguous()


def weighted_cov(Y, w, normalize=True, frames=None):
    """Phi[b,f,k] = sum_t w'[b,f,k,t] y y^H.  Y (B,F,D,T) c64, w (B,F,K,T) f32."""
    Y = _need(Y, torch.complex64, 4, 'Y')
    w = _need(w, torch.float32, 4, 'w')
    B, F, D, T = Y.shape
    K = w.shape[2]
    assert w.shape == (B, F, K, T), (w.shape, Y.shape)
    out = torch.empty((B, F, K, D, D), dtype=torch.complex64, device=Y.device)
    n = _lib.workspace_bytes(_lib.OP_WEIGHTED_COV, B, F, D, T, K, 0)
    ws = workspace(n, Y.device)
    _lib.check(_lib.lib().gss_weighted_cov_c64(_ptr(Y), _ptr(w), _ptr(out), 1 if normalize else 0,
                                               B, F, D, T, K, _ptr(_tper(frames, B, Y.device)),
                                               _ptr(ws), ws.numel(), _stream()))
    return out


def cacgmm(Y, activity, iterations, iterations_post=1, affiliation_eps=1e-10,
           eigenvalue_floor=1e-10, return_model=False, frames=None, info=None):
    """Guided CACGMM EM.  Y (B,F,D,T) c64, activity (B,K,T_act) bool/uint8 ->
    posterior (B,F,K,T) f32 [, model dict].  info: optional (B,) int32 device tensor; when given the
    status words are left for the caller to check (no host synchronisation here)."""
    f64 = isinstance(Y, torch.Tensor) and Y.dtype == torch.complex128      # float64 hand-off (runtime-shape kernel)
    Y = _need(Y, torch.complex128 if f64 else torch.complex64, 4, 'Y')
    B, F, D, T = Y.shape
    if activity.dtype == torch.bool:
        activity = activity.to(torch.uint8)
    activity = _need(activity, torch.uint8, 3, 'activity')
    assert activity.shape[0] == B, (activity.shape, Y.shape)
    K, T_act = activity.shape[1], activity.shape[2]
    post = torch.empty((B, F, K, T), dtype=torch.float32, device=Y.device)
    own_info = info is None
    if own_info:
        info = new_info(B, Y.device)
    weight = logdet = cov = None
    if return_model:
        weight = torch.empty((B, F, K), dtype=torch.float64, device=Y.device)
        logdet = torch.empty((B, F, K), dtype=torch.float64, device=Y.device)
        cov = torch.empty((B, F, K, D, D), dtype=torch.complex128, device=Y.device)
    ws = workspace(_lib.workspace_bytes(_lib.OP_CACGMM_C128 if f64 else _lib.OP_CACGMM, B, F, D, T, K, 0), Y.device)
    _lib.check((_lib.lib().gss_cacgmm_c128 if f64 else _lib.lib().gss_cacgmm_c64)(
        _ptr(Y), _ptr(activity), _ptr(post), int(iterations), int(iterations_post),
        float(affiliation_eps), float(eigenvalue_floor), B, F, D, T, K, T_act, _ptr(_tper(frames, B, Y.device)),
        _ptr(weight), _ptr(logdet), _ptr(cov), _ptr(info), _ptr(ws), ws.numel(), _stream()))
    if own_info:
        check_info(info, 'cacgmm')
    if return_model:
        return post, dict(weight=weight, log_determinant=logdet, covariance=cov)
    return post


def _bf_call(Y, fn, lead_args, bf, bf_arg, postfilter, K=None, return_aux=False, frames=None, info=None):
    B, F, D, T = Y.shape
    if isinstance(bf, int) and bf & 0x100:
        bf_code = bf                                   # a get_bf_vector program (GSS_BF_PROGRAM, include/gss.h)
    elif bf in _lib.BF_TYPES:
        bf_code = _lib.BF_TYPES[bf]
    else:
        raise NotImplementedError(bf)
    if postfilter not in _lib.POSTFILTERS:
        raise NotImplementedError(postfilter)
    X = torch.empty((B, F, T), dtype=torch.complex64, device=Y.device)
    ref = torch.full((max(B, 1),), -1, dtype=torch.int32, device=Y.device)
    own_info = info is None
    if own_info:
        info = new_info(B, Y.device)
    wts = torch.zeros((B, F, D), dtype=torch.complex128, device=Y.device) if return_aux else None
    n = _lib.workspace_bytes(_lib.OP_BEAMFORM, B, F, D, T, 2, 0)
    ws = workspace(n, Y.device)
    dims = (B, F, D, T) if K is None else (B, F, D, T, K)
    _lib.check(fn(_ptr(Y), *lead_args, _ptr(X), bf_code, int(bf_arg), _lib.POSTFILTERS[postfilter],
                  *dims, _ptr(_tper(frames, B, Y.device)), _ptr(ref), _ptr(wts), _ptr(info), _ptr(ws), ws.numel(),
                  _stream()))
    if own_info:
        check_info(info, 'beamform')
    if return_aux:
        return X, dict(ref_channel=ref, weights=wts)
    return X


def beamform(Y, target_mask, distortion_mask, bf='mvdrSouden_ban', postfilter=None, bf_arg=0,
             return_aux=False, frames=None, info=None):
    """Y (B,F,D,T) c64; masks (B,F,T) f32 -> X_hat (B,F,T) c64."""
    Y = _need(Y, torch.complex64, 4, 'Y')
    B, F, D, T = Y.shape
    tm = _need(target_mask, torch.float32, 3, 'target_mask')
    dm = _need(distortion_mask, torch.float32, 3, 'distortion_mask')
    assert tm.shape == (B, F, T), (tm.shape, B, F, T)
    assert dm.shape == (B, F, T), (dm.shape, B, F, T)
    return _bf_call(Y, _lib.lib().gss_beamform_c64, (_ptr(tm), _ptr(dm)), bf, bf_arg, postfilter,
                    return_aux=return_aux, frames=frames, info=info)


def beamform_from_posterior(Y, posterior, target_index, start_ctx=None, end_ctx=None,
                            bf='mvdrSouden_ban', postfilter=None, bf_arg=0, return_aux=False, frames=None,
                            info=None):
    """Fused core.py:537-564.  posterior (B,F,K,T) f32; target_index/start_ctx/end_ctx (B) int32."""
    Y = _need(Y, torch.complex64, 4, 'Y')
    B, F, D, T = Y.shape
    post = _need(posterior, torch.float32, 4, 'posterior')
    K = post.shape[2]
    assert post.shape == (B, F, K, T), (post.shape, Y.shape)

    def ivec(v):
        if v is None:
            return None
        v = torch.as_tensor(v, dtype=torch.int32, device=Y.device).reshape(-1)
        assert v.numel() == B
        return v.contiguous()

    ti, sc, ec = ivec(target_index), ivec(start_ctx), ivec(end_ctx)
    return _bf_call(Y, _lib.lib().gss_beamform_from_posterior_c64,
                    (_ptr(post), _ptr(ti), _ptr(sc), _ptr(ec)), bf, bf_arg, postfilter, K=K,
                    return_aux=return_aux, frames=frames, info=info)


def wpe(Y, taps=10, delay=3, iterations=3, psd_context=0, frames=None, gram_mode=None, i8_tau=None,
        stats=None, info=None, return_f64=False):
    """Y (B,F,D,T) c64 -> dereverberated (B,F,D,T) c64.

    gram_mode: None / 'auto' (INT8 tensor-core correlation build where it is built, ill-conditioned
    bins re-done in float64), 'f64', 'i8' (diagnostics: no re-do), 'i8+redo'.  stats: optional int32[4]
    device tensor, accumulated by the call ([0] bins, [1] bins that ended on the float64 list,
    [2] float64 re-do builds).  info: optional (B,) int32 device tensor for GSS_INFO_SINGULAR words
    (when omitted a fresh one is allocated and checked here, which synchronises).
    return_f64: also return the result before its rounding to complex64, (B,F,D,T) complex128 -- the
    float64 hand-off to `cacgmm` (which accepts complex128 observations)."""
    Y = _need(Y, torch.complex64, 4, 'Y')
    B, F, D, T = Y.shape
    if gram_mode not in _lib.WPE_GRAM_MODES:
        raise NotImplementedError(gram_mode)
    X = torch.empty_like(Y)
    X64 = torch.empty(Y.shape, dtype=torch.complex128, device=Y.device) if return_f64 else None
    own_info = info is None
    if own_info:
        info = new_info(B, Y.device)
    n = _lib.workspace_bytes(_lib.OP_WPE, B, F, D, T, 0, int(taps))
    ws = workspace(n, Y.device)
    _lib.check(_lib.lib().gss_wpe_c64_ex(_ptr(Y), _ptr(X), int(taps), int(delay), int(iterations),
                                         int(psd_context), B, F, D, T, _ptr(_tper(frames, B, Y.device)),
                                         _lib.WPE_GRAM_MODES[gram_mode], -1.0 if i8_tau is None else float(i8_tau),
                                         _ptr(stats), _ptr(X64), _ptr(info), _ptr(ws), ws.numel(), _stream()))
    if own_info:
        check_info(info, 'wpe')
    return (X, X64) if return_f64 else X


def stft_frames(N, size, shift, fading):
    pad = (size - shift) if fading else 0
    total = N + 2 * pad
    return 1 if total <= size else -(-(total - size) // shift) + 1


def stft(x, size=1024, shift=256, fading=True):
    """x (B,D,N) float32 -> Y (B,F,D,T) complex64 (bin-major)."""
    x = _need(x, torch.float32, 3, 'x')
    B, D, N = x.shape
    F, T = size // 2 + 1, stft_frames(N, size, shift, fading)
    Y = torch.empty((B, F, D, T), dtype=torch.complex64, device=x.device)
    ws = workspace(256, x.device)
    _lib.check(_lib.lib().gss_stft_f32(_ptr(x), _ptr(Y), B, D, N, int(size), int(shift), 1 if fading else 0,
                                       _ptr(ws), ws.numel(), _stream()))
    return Y


def istft(X, size=1024, shift=256, fading=True):
    """X (B,F,T) complex64 -> x (B, T*shift + (size-shift)*(1-2*fading)) float32."""
    X = _need(X, torch.complex64, 3, 'X')
    B, F, T = X.shape
    assert F == size // 2 + 1, (X.shape, size)
    drop = (size - shift) if fading else 0
    nout = T * shift + size - shift - 2 * drop
    out = torch.empty((B, nout), dtype=torch.float32, device=X.device)
    n = _lib.workspace_bytes(_lib.OP_ISTFT, B, F, 1, T, 1, 0)
    ws = workspace(n, X.device)
    _lib.check(_lib.lib().gss_istft_f32(_ptr(X), _ptr(out), B, T, int(size), int(shift), 1 if fading else 0,
                                        _ptr(ws), ws.numel(), _stream()))
    return out


def enhance(Obs, activity, target_index, start_ctx=None, end_ctx=None, frames=None, *, wpe=None,
            em_iterations=20, em_iterations_post=1, bf='mvdrSouden_ban', postfilter=None, bf_arg=0,
            return_posterior=True, handoff='c64'):
    """Whole STFT-domain hot path in ONE library call (gss_enhance_c64), reference layouts:
    Obs (B,D,T,F) c64, activity (B,K,T_act) -> X_hat (B,T,F) c64 [, posterior (B,K,T,F) f32].
    wpe: None or (taps, delay, iterations, psd_context).  handoff: 'c64' (default: the blocks exchange
    complex64 tensors) or 'f64' (gss_enhance_c64_ex with GSS_ENHANCE_F64_HANDOFF: the dereverberated spectrum
    reaches the EM unrounded, as in the float64 reference; slower runtime-shape EM kernel)."""
    Obs = _need(Obs, torch.complex64, 4, 'Obs')
    B, D, T, F = Obs.shape
    if activity.dtype == torch.bool:
        activity = activity.to(torch.uint8)
    activity = _need(activity, torch.uint8, 3, 'activity')
    K, T_act = activity.shape[1], activity.shape[2]
    if bf not in _lib.BF_TYPES:
        raise NotImplementedError(bf)
    if postfilter not in _lib.POSTFILTERS:
        raise NotImplementedError(postfilter)
    taps, delay, its, ctx = wpe if wpe is not None else (0, 0, 0, 0)
    ivec = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.int32).reshape(-1).to(Obs.device).contiguous()
    ti, sc, ec = ivec(target_index), ivec(start_ctx), ivec(end_ctx)
    assert ti is not None and ti.numel() == B
    X = torch.empty((B, T, F), dtype=torch.complex64, device=Obs.device)
    post = torch.empty((B, K, T, F), dtype=torch.float32, device=Obs.device) if return_posterior else None
    info = torch.zeros((max(B, 1),), dtype=torch.int32, device=Obs.device)
    assert handoff in ('c64', 'f64'), handoff
    f64 = handoff == 'f64'
    ws = workspace(_lib.workspace_bytes(_lib.OP_ENHANCE_F64 if f64 else _lib.OP_ENHANCE, B, F, D, T, K, int(taps)), Obs.device)
    _lib.check(_lib.lib().gss_enhance_c64_ex(
        _ptr(Obs), _ptr(activity), _ptr(ti), _ptr(sc), _ptr(ec), _ptr(_tper(frames, B, Obs.device)),
        _ptr(X), _ptr(post), int(taps), int(delay), int(its), int(ctx), int(em_iterations), int(em_iterations_post),
        _lib.BF_TYPES[bf], int(bf_arg), _lib.POSTFILTERS[postfilter], 1 if f64 else 0, B, F, D, T, K, T_act,
        _ptr(info), _ptr(ws), ws.numel(), _stream()))
    check_info(info, 'enhance')
    return (X, post) if return_posterior else X
