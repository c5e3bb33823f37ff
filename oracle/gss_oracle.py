"""NumPy restatement of the pb_chime5 Enhancer hot path (float64 / complex128).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function cites the
reference lines it restates (paths relative to ``/root/reference``).  All
functions accept arbitrary leading "independent" axes so that one call can
process all frequency bins (fast, used by the tests) or a single bin (the
reference's own loop structure, used for the CPU timing baseline).

Shapes:  D channels, T frames, F bins, K classes, N samples.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
import scipy.signal

F64_TINY = np.finfo(np.float64).tiny


# ---------------------------------------------------------------------------
# STFT / iSTFT  (nara_wpe.utils, third-party, not vendored; call sites
# pb_chime5/core.py:305-321, 224-237; framing pinned by the doctest at
# pb_chime5/database/chime5/database.py:417-453)
# ---------------------------------------------------------------------------

def samples_to_stft_frames(samples, size, shift, *, pad=True, fading=False):
    """nara_wpe.utils._samples_to_stft_frames (call: core.py:224-237)."""
    if fading:
        samples = samples + 2 * (size - shift)
    frames = (samples - size + shift) / shift
    return int(np.ceil(frames)) if pad else int(np.floor(frames))


def _frame(x, length, shift, pad):
    """Chop the last axis into frames (segment_axis_v2 semantics,
    pb_chime5/utils/numpy_utils.py:10-222, end='pad' | 'cut')."""
    n = x.shape[-1]
    if pad:
        if n < length:
            extra = length - n
        else:
            extra = (-(n - length)) % shift
        if extra:
            widths = [(0, 0)] * (x.ndim - 1) + [(0, extra)]
            x = np.pad(x, widths, mode='constant')
        n = x.shape[-1]
    count = (n - length) // shift + 1
    idx = np.arange(length)[None, :] + shift * np.arange(count)[:, None]
    return x[..., idx]


def analysis_window(size, window='blackman'):
    """Periodic (DFT-even) window: scipy window of size+1 without last tap."""
    if callable(window):
        return np.asarray(window(size + 1)[:-1], dtype=np.float64)
    return np.asarray(scipy.signal.get_window(window, size + 1, fftbins=False)[:-1],
                      dtype=np.float64)


def stft(x, size=1024, shift=256, fading=True, window='blackman', pad=True):
    """(..., N) real -> (..., T, size//2+1) complex128.  core.py:305-312."""
    x = np.asarray(x, dtype=np.float64)
    if fading:
        widths = [(0, 0)] * (x.ndim - 1) + [(size - shift, size - shift)]
        x = np.pad(x, widths, mode='constant')
    w = analysis_window(size, window)
    frames = _frame(x, size, shift, pad)
    return np.fft.rfft(frames * w, n=size, axis=-1)


def synthesis_window(size, shift, window='blackman'):
    """Biorthogonal synthesis window w / sum_i w[i*shift + n mod shift]^2."""
    w = analysis_window(size, window)
    assert size % shift == 0, (size, shift)
    ssq = (w.reshape(size // shift, shift) ** 2).sum(axis=0)
    return w / np.tile(ssq, size // shift)


def istft(X, size=1024, shift=256, fading=True, window='blackman'):
    """(..., T, size//2+1) -> (..., T*shift - (size-shift)) if fading.
    core.py:314-321."""
    X = np.asarray(X)
    assert X.shape[-1] == size // 2 + 1, X.shape
    T = X.shape[-2]
    ws = synthesis_window(size, shift, window)
    seg = np.fft.irfft(X, n=size, axis=-1) * ws
    out = np.zeros(X.shape[:-2] + (T * shift + size - shift,))
    for t in range(T):
        out[..., t * shift:t * shift + size] += seg[..., t, :]
    if fading:
        out = out[..., size - shift:out.shape[-1] - (size - shift)]
    return out


def activity_time_to_frequency(time_activity, stft_window_length, stft_shift,
                               stft_fading, stft_pad=True):
    """Sample-level bool activity -> frame-level ("any sample in the frame").
    pb_chime5/database/chime5/database.py:409-472."""
    a = np.asarray(time_activity)
    if stft_fading:
        p = stft_window_length - stft_shift
        widths = [(0, 0)] * (a.ndim - 1) + [(p, p)]
        a = np.pad(a, widths, mode='constant')
    return _frame(a, stft_window_length, stft_shift, stft_pad).any(axis=-1)


# ---------------------------------------------------------------------------
# WPE  (nara_wpe.wpe.wpe_v8 -> wpe_v6; call sites core.py:52-58, 63-69, 72-78)
# PARITY UNPINNED: third-party arithmetic restated from the published
# algorithm (Nakatani 2010; Drude 2018), SURVEY.md appendix A.
# ---------------------------------------------------------------------------

def wpe_power_inverse(X, psd_context=0):
    """1 / max(mean_d |X|^2, 1e-10 * max_t) ; (..., D, T) -> (..., T)."""
    power = np.mean(X.real ** 2 + X.imag ** 2, axis=-2)
    if psd_context > 0:
        c = int(psd_context)
        T = power.shape[-1]
        csum = np.cumsum(np.pad(power, [(0, 0)] * (power.ndim - 1) + [(c + 1, c)]),
                         axis=-1)
        ones = np.cumsum(np.pad(np.ones(T), (c + 1, c)))
        num = csum[..., 2 * c + 1:] - csum[..., :-(2 * c + 1)]
        den = ones[2 * c + 1:] - ones[:-(2 * c + 1)]
        power = num / den
    eps = 1e-10 * np.max(power, axis=-1, keepdims=True)
    return 1.0 / np.maximum(power, eps)


def wpe_tap_matrix(Y, taps, delay):
    """(..., D, T) -> (..., taps*D, T); row k*D+d at frame t is Y[d, t-delay-k]
    (zero history).  Tap ordering is irrelevant for the result."""
    *lead, D, T = Y.shape
    out = np.zeros((*lead, taps, D, T), dtype=Y.dtype)
    for k in range(taps):
        s = delay + k
        if s < T:
            out[..., k, :, s:] = Y[..., :, :T - s]
    return out.reshape(*lead, taps * D, T)


def _solve_with_fallback(A, B):
    """np.linalg.solve, per-matrix fallback to lstsq on singular input
    (pb_bss/pb_bss/math/solve.py:20-114; nara_wpe uses the same pattern)."""
    A = np.asarray(A)
    B = np.asarray(B)
    try:
        return np.linalg.solve(A, B)
    except np.linalg.LinAlgError:
        a = A.reshape((-1,) + A.shape[-2:])
        b = B.reshape((-1,) + B.shape[-2:])
        c = np.zeros_like(b)
        for i in range(a.shape[0]):
            try:
                c[i] = np.linalg.solve(a[i], b[i])
            except np.linalg.LinAlgError:
                c[i] = np.linalg.lstsq(a[i], b[i], rcond=None)[0]
        return c.reshape(B.shape)


def stable_solve(A, B):
    """pb_bss/pb_bss/math/solve.py:20-114."""
    return _solve_with_fallback(A, B)


def wpe_bins(Y, taps=10, delay=3, iterations=3, psd_context=0):
    """WPE on (..., D, T) complex; all leading axes are independent bins."""
    Y = np.asarray(Y, dtype=np.complex128)
    X = Y.copy()
    Yt = wpe_tap_matrix(Y, taps, delay)
    YtH = np.conj(np.swapaxes(Yt, -1, -2))
    YH = np.conj(np.swapaxes(Y, -1, -2))
    for _ in range(iterations):
        inv = wpe_power_inverse(X, psd_context)
        Yw = Yt * inv[..., None, :]
        R = Yw @ YtH
        P = Yw @ YH
        G = _solve_with_fallback(R, P)
        X = Y - np.conj(np.swapaxes(G, -1, -2)) @ Yt
    return X


def wpe_dtf(Obs, taps, delay, iterations, psd_context=0, loop_over_bins=False):
    """WPE.__call__ for 3-D input (D, T, F) -> (D, T, F).  core.py:48-58."""
    Y = np.transpose(Obs, (2, 0, 1))
    if loop_over_bins:
        out = np.empty(Y.shape, dtype=np.complex128)
        for f in range(Y.shape[0]):
            out[f] = wpe_bins(Y[f], taps, delay, iterations, psd_context)
    else:
        out = wpe_bins(Y, taps, delay, iterations, psd_context)
    return np.transpose(out, (1, 2, 0))


# ---------------------------------------------------------------------------
# CACGMM  (pb_bss/pb_bss/distribution/*)
# ---------------------------------------------------------------------------

def unit_norm_frames(y):
    """(..., T, D) -> (..., D, T) unit-norm frames; all-zero frames stay zero.
    complex_angular_central_gaussian.py:34-55, distribution/utils.py:228-261
    (eps_style='where')."""
    norm = np.linalg.norm(y, axis=-1, keepdims=True)
    norm = np.where(norm == 0, np.finfo(y.dtype).tiny, norm)
    return np.ascontiguousarray(np.swapaxes(y / norm, -2, -1))


def cacg_quadratic_form(y, eigvec, eigval):
    """q[..., k, t] = max(|y^H V diag(1/lambda) V^H y|, tiny).
    y (..., D, T); eigvec (..., K, D, D); eigval (..., K, D).
    complex_angular_central_gaussian.py:185-199."""
    q = np.einsum('...dt,...kde,...ke,...kge,...gt->...kt',
                  y.conj(), eigvec, 1 / eigval, eigvec.conj(), y,
                  optimize='optimal')
    return np.maximum(np.abs(q), F64_TINY)


def cacg_log_pdf(y, eigvec, eigval):
    """complex_angular_central_gaussian.py:166-203."""
    D = y.shape[-2]
    q = cacg_quadratic_form(y, eigvec, eigval)
    log_pdf = -D * np.log(q)
    log_pdf -= np.sum(np.log(eigval), axis=-1)[..., None]
    return log_pdf, q


def posterior_from_log_pdf(weight, log_pdf, source_activity_mask=None,
                           affiliation_eps=0.0):
    """mixture_model_utils.py:7-55."""
    aff = log_pdf - np.amax(log_pdf, axis=-2, keepdims=True)
    np.exp(aff, out=aff)
    aff *= weight
    if source_activity_mask is not None:
        aff *= source_activity_mask
    aff /= np.maximum(np.sum(aff, axis=-2, keepdims=True), F64_TINY)
    if affiliation_eps != 0:
        aff = np.clip(aff, affiliation_eps, 1 - affiliation_eps)
    return aff


def cacg_from_covariance(cov, eigenvalue_floor):
    """eigh -> normalise by the largest eigenvalue -> floor.
    complex_angular_central_gaussian.py:81-131 (covariance_norm='eigenvalue')."""
    eigval, eigvec = np.linalg.eigh(cov)
    eigval = eigval / np.maximum(np.amax(eigval, axis=-1, keepdims=True), F64_TINY)
    eigval = np.maximum(eigval, eigenvalue_floor)
    return eigvec, eigval


def cacgmm_m_step(y, q, aff, eigenvalue_floor=1e-10):
    """cacgmm.py:313-343 + complex_angular_central_gaussian.py:253-310 +
    mixture_model_utils.py:187-190 (weight_constant_axis=(-1,)).
    y (..., D, T); q, aff (..., K, T)."""
    D = y.shape[-2]
    weight = np.mean(aff, axis=-1, keepdims=True)
    denom = np.sum(aff, axis=-1)[..., None, None]
    cov = D * np.einsum('...kt,...dt,...et->...kde', aff / q, y, y.conj(),
                        optimize='greedy')
    cov = cov / denom
    cov = (cov + np.conj(np.swapaxes(cov, -1, -2))) / 2   # utils.py:323-334
    eigvec, eigval = cacg_from_covariance(cov, eigenvalue_floor)
    return weight, eigvec, eigval


def cacgmm_fit(y_td, initialization, iterations, source_activity_mask=None,
               affiliation_eps=1e-10, eigenvalue_floor=1e-10):
    """CACGMMTrainer.fit with an affiliation initialisation.  cacgmm.py:141-278.
    y_td (..., T, D); initialization / mask (..., K, T).
    Returns the model (weight (...,K,1), eigvec (...,K,D,D), eigval (...,K,D))."""
    y = unit_norm_frames(np.asarray(y_td, dtype=np.complex128))
    aff = np.asarray(initialization, dtype=np.float64)
    q = np.ones(aff.shape, dtype=np.float64)
    model = None
    for _ in range(iterations):
        if model is not None:
            weight, eigvec, eigval = model
            log_pdf, q = cacg_log_pdf(y, eigvec, eigval)
            aff = posterior_from_log_pdf(weight, log_pdf, source_activity_mask,
                                         affiliation_eps)
        model = cacgmm_m_step(y, q, aff, eigenvalue_floor)
    return model


def cacgmm_predict(model, y_td):
    """CACGMM.predict: final E-step, no activity mask, no clipping.
    cacgmm.py:63-94."""
    weight, eigvec, eigval = model
    y = unit_norm_frames(np.asarray(y_td, dtype=np.complex128))
    log_pdf, _ = cacg_log_pdf(y, eigvec, eigval)
    return posterior_from_log_pdf(weight, log_pdf)


def gss_init(activity_freq):
    """core.py:156-163 (the reference hard-codes 513 copies; values are the
    same for every bin so one copy is kept here)."""
    init = np.asarray(activity_freq, dtype=np.float64)
    init = np.where(init == 0, 1e-10, init)
    init = init / np.sum(init, keepdims=True, axis=0)
    mask = np.asarray(activity_freq, dtype=bool)
    return init, mask


def gss_posteriors(Obs, activity_freq, iterations, iterations_post=1,
                   loop_over_bins=False, return_models=False):
    """GSS.__call__: Obs (D, T, F) complex, activity (K, T_act) bool
    -> posterior (K, T, F) float64.  core.py:154-214."""
    assert iterations_post >= 1, 'iterations_post=0 raises TypeError in the reference'
    init, mask = gss_init(activity_freq)
    Y = np.asarray(Obs).T                       # (F, T, D), as Obs.T in core.py:181
    F, T, D = Y.shape
    init = init[..., :T]
    mask = mask[..., :T]

    def one(y):
        model = cacgmm_fit(y, np.broadcast_to(init, y.shape[:-2] + init.shape),
                           iterations,
                           np.broadcast_to(mask, y.shape[:-2] + mask.shape))
        if iterations_post > 1:
            # unguided refinement, initialised from the model (core.py:188-194)
            yy = unit_norm_frames(np.asarray(y, dtype=np.complex128))
            for _ in range(iterations_post - 1):
                weight, eigvec, eigval = model
                log_pdf, q = cacg_log_pdf(yy, eigvec, eigval)
                aff = posterior_from_log_pdf(weight, log_pdf, None, 1e-10)
                model = cacgmm_m_step(yy, q, aff, 1e-10)
        return model, cacgmm_predict(model, y)

    if loop_over_bins:
        outs = [one(Y[f]) for f in range(F)]
        post = np.array([o[1] for o in outs])
        models = [o[0] for o in outs]
    else:
        models, post = one(Y)
    post = post.transpose(1, 2, 0)              # (K, T, F)
    return (post, models) if return_models else post


# ---------------------------------------------------------------------------
# Beamforming  (pb_bss/pb_bss/extraction/beamformer.py,
#               pb_chime5/speech_enhancement/beamforming_wrapper.py)
# ---------------------------------------------------------------------------

def psd_matrix(Y, mask, normalize=True):
    """Y (..., D, T), mask (..., T) -> (..., D, D).  beamformer.py:61-145."""
    mask = np.array(mask, dtype=np.float64)
    if normalize:
        mask = mask / np.maximum(np.sum(mask, axis=-1, keepdims=True), 1e-10)
    return np.einsum('...dt,...et->...de', mask[..., None, :] * Y, Y.conj())


def optimal_reference_channel(w_mat, psd_x, psd_n, eps):
    """beamformer.py:524-543."""
    num = np.einsum('...FdR,...FdD,...FDR->...R', w_mat.conj(), psd_x, w_mat)
    den = np.einsum('...FdR,...FdD,...FDR->...R', w_mat.conj(), psd_n, w_mat)
    snr = num / np.maximum(den, eps)
    assert np.all(np.isfinite(snr)), snr
    return int(np.argmax(snr.real))


def mvdr_souden(psd_x, psd_n, ref_channel=None, eps=None, return_ref_channel=False):
    """beamformer.py:546-617."""
    psd_x = np.asarray(psd_x)
    psd_n = np.asarray(psd_n)
    phi = _solve_with_fallback(psd_n, psd_x)
    lam = np.trace(phi, axis1=-1, axis2=-2)[..., None, None]
    if eps is None:
        eps = np.finfo(lam.real.dtype).tiny
    mat = phi / np.maximum(lam.real, eps)
    if ref_channel is None:
        ref_channel = optimal_reference_channel(mat, psd_x, psd_n, eps)
    w = mat[..., ref_channel]
    return (w, ref_channel) if return_ref_channel else w


def blind_analytic_normalization(w, psd_n):
    """beamformer.py:396-418."""
    num = np.sqrt(np.einsum('...a,...ab,...bc,...c->...', w.conj(), psd_n, psd_n, w))
    den = np.einsum('...a,...ab,...b->...', w.conj(), psd_n, w)
    den = np.sqrt(den * den.conj())
    norm = np.divide(num, den, out=np.zeros_like(num), where=den != 0)
    return w * np.abs(norm[..., None])


def gev_vector(psd_x, psd_n):
    """Principal generalised eigenvector per bin (scipy eigh(A, B) fallback of
    beamformer.py:317-348; the Cython zhegvd path get_gev_vector.pyx:42-150
    yields the same vector up to a unit-modulus factor)."""
    psd_x = np.asarray(psd_x, dtype=np.complex128)
    psd_n = np.asarray(psd_n, dtype=np.complex128)
    shape = psd_x.shape
    D = shape[-1]
    a = psd_x.reshape(-1, D, D)
    b = psd_n.reshape(-1, D, D)
    out = np.empty((a.shape[0], D), dtype=np.complex128)
    for f in range(a.shape[0]):
        vals, vecs = scipy.linalg.eigh(a[f], b[f])
        out[f] = vecs[:, np.argmax(vals)]
    return out.reshape(shape[:-1])


def apply_beamforming_vector(w, Y):
    """beamformer.py:502-510."""
    return np.einsum('...a,...at->...t', w.conj(), Y)


def canonical_phase(w, psd_n, ref=0):
    """Deterministic per-bin phase for GEV comparisons (the reference compares
    GEV vectors by cosine similarity only, test_beamformer.py:17-21,140).
    Makes (Phi_NN w)[ref] real and non-negative."""
    z = np.einsum('...ab,...b->...a', psd_n, w)[..., ref]
    ph = np.where(np.abs(z) > 0, z / np.maximum(np.abs(z), F64_TINY), 1.0)
    return w * np.conj(ph)[..., None]


def beamform(Obs, target_mask, distortion_mask, bf='mvdrSouden_ban',
             postfilter=None, return_aux=False, ref_channel=None):
    """Beamformer.__call__: Obs (D,T,F), masks (T,F) -> X_hat (T,F).
    core.py:246-278; beamforming_wrapper.py:11-124, 192-208."""
    Obs = np.asarray(Obs)
    aux = {}
    if bf in ('mvdrSouden_ban', 'gev_ban', 'mvdrSouden', 'gev'):
        Y = np.transpose(Obs, (2, 0, 1))                 # 'DTF->FDT'
        Xm = np.transpose(target_mask, (1, 0))           # 'TF->FT'
        Nm = np.transpose(distortion_mask, (1, 0))
        F, D, T = Y.shape
        assert D < 30, (D, Y.shape)
        cov_x = psd_matrix(Y, Xm)
        cov_n = psd_matrix(Y, Nm)
        if bf.startswith('mvdrSouden'):
            # ref_channel=None: the reference's arg-max over ALL bins of the call
            # (beamformer.py:535-543); a caller that checks a subset of bins passes the
            # channel the full-band run selected
            w, ref = mvdr_souden(cov_x, cov_n, ref_channel=ref_channel, eps=1e-10,
                                 return_ref_channel=True)
            aux['ref_channel'] = ref
        else:
            w = canonical_phase(gev_vector(cov_x, cov_n), cov_n)
        if bf.endswith('_ban'):
            w = blind_analytic_normalization(w, cov_n)
        aux.update(cov_x=cov_x, cov_n=cov_n, w=w)
        X_hat = apply_beamforming_vector(w, Y).T
    elif bf == 'ch2':
        X_hat = Obs[2]
    elif bf == 'sum':
        X_hat = np.sum(Obs, axis=0)
    else:
        raise NotImplementedError(bf)
    if postfilter is None:
        pass
    elif postfilter == 'mask_mul':
        X_hat = X_hat * target_mask
    else:
        raise NotImplementedError(postfilter)
    return (X_hat, aux) if return_aux else X_hat


# ---------------------------------------------------------------------------
# Whole path  (core.py:514-571)
# ---------------------------------------------------------------------------

def enhance_stft(Obs, activity_freq, target_index, *, wpe=None, gss_iterations=20,
                 gss_iterations_post=1, bf='mvdrSouden_ban', postfilter=None,
                 start_context_frames=0, end_context_frames=0, bf_drop_context=True,
                 loop_over_bins=False, ref_channel=None):
    """STFT-domain part of Enhancer.enhance_observation (core.py:524-564).
    wpe: None or dict(taps, delay, iterations, psd_context).
    Returns dict with Obs (after WPE), masks, X_hat."""
    Obs = np.asarray(Obs, dtype=np.complex128)
    if wpe is not None:
        Obs = wpe_dtf(Obs, loop_over_bins=loop_over_bins, **wpe)
    masks = gss_posteriors(Obs, activity_freq, gss_iterations, gss_iterations_post,
                           loop_over_bins=loop_over_bins)
    if bf_drop_context:
        masks[:, :start_context_frames, :] = 0
        if end_context_frames > 0:
            masks[:, -end_context_frames:, :] = 0
    target_mask = masks[target_index]
    distortion_mask = np.sum(np.delete(masks, target_index, axis=0), axis=0)
    X_hat = beamform(Obs, target_mask, distortion_mask, bf, postfilter,
                     ref_channel=ref_channel)
    return dict(Obs=Obs, masks=masks, target_mask=target_mask,
                distortion_mask=distortion_mask, X_hat=X_hat)


def enhance_observation(obs, sample_activity, target_index, *, stft_size=1024,
                        stft_shift=256, stft_fading=True, start_context_samples=0,
                        end_context_samples=0, **kwargs):
    """Enhancer.enhance_observation on raw audio (D, N) with per-class sample
    activity (K, N) bool.  core.py:514-571."""
    Obs = stft(obs, stft_size, stft_shift, stft_fading)           # (D, T, F)
    act = activity_time_to_frequency(sample_activity, stft_size, stft_shift,
                                     stft_fading, True)
    sc = samples_to_stft_frames(start_context_samples, stft_size, stft_shift,
                                fading=stft_fading)
    ec = samples_to_stft_frames(end_context_samples, stft_size, stft_shift,
                                fading=stft_fading)
    out = enhance_stft(Obs, act, target_index, start_context_frames=sc,
                       end_context_frames=ec, **kwargs)
    out['x_hat'] = istft(out['X_hat'], stft_size, stft_shift, stft_fading)
    out['activity_freq'] = act
    return out


# ---------------------------------------------------------------------------
# get_bf_vector DSL  (pb_bss/pb_bss/extraction/beamformer_wrapper.py:108-227 and the
# beamformer.py functions it dispatches to); SURVEY.md section 8 row f4
# ---------------------------------------------------------------------------

def pca_vector(psd, scaling=None):
    """beamformer.py:148-202: principal eigenvector (np.linalg.eigh, largest eigenvalue)."""
    psd = np.asarray(psd)
    vals, vecs = np.linalg.eigh(psd)
    v, lam = vecs[..., -1], vals[..., -1]
    if scaling is None:
        return v
    if scaling == 'trace':
        return v * (np.sqrt(np.trace(psd, axis1=-1, axis2=-2)) / np.linalg.norm(v, axis=-1))[..., None]
    if scaling == 'eigenvalue':
        return v * (lam / np.linalg.norm(v, axis=-1))[..., None]
    raise ValueError(scaling)


def mvdr_vector(atf, psd_n):
    """beamformer.py:205-235: Phi_N^-1 a / (a^H Phi_N^-1 a), Phi_N hermitised first."""
    psd_n = 0.5 * (psd_n + np.conj(np.swapaxes(psd_n, -1, -2)))
    num = np.linalg.solve(psd_n, atf[..., None])[..., 0]
    den = np.einsum('...d,...d->...', atf.conj(), num)
    return num / den[..., None]


def gev_atf_vector(psd_x, psd_n):
    """beamformer_wrapper.py:24-44: Phi_N w_gev."""
    return np.einsum('...dD,...D->...d', psd_n, gev_vector(psd_x, psd_n))


def rank1_approximation(kind, psd_x, psd_n):
    """beamformer_wrapper.py:11-21, 47-62, 88-101: trace-preserving rank-1 model of Phi_X."""
    if kind == 'rank1_pca':
        a = pca_vector(psd_x)
    elif kind == 'rank1_gev':
        a = gev_atf_vector(psd_x, psd_n)
    else:
        raise ValueError(kind, 'use either rank1_pca or rank1_gev')
    r1 = np.einsum('...d,...D->...dD', a, a.conj())
    scale = np.trace(psd_x, axis1=-1, axis2=-2) / np.trace(r1, axis1=-1, axis2=-2)
    return scale[..., None, None] * r1


def wmwf_vector(psd_x, psd_n, reference_channel=None, distortion_weight=1.0, return_ref_channel=False):
    """beamformer.py:620-672 (no channel_selection_vector)."""
    phi = _solve_with_fallback(psd_n, psd_x)
    lam = np.trace(phi, axis1=-1, axis2=-2)[..., None, None]
    if isinstance(distortion_weight, str):
        assert distortion_weight == 'frequency_dependent', distortion_weight
        filt = phi / np.sqrt(psd_x[..., 0:1, 0:1] * lam)
    else:
        filt = phi / (distortion_weight + lam)
    if reference_channel is None:
        # beamformer.py:524-543 with its default eps
        reference_channel = optimal_reference_channel(filt, psd_x, psd_n, F64_TINY)
    w = filt[..., reference_channel]
    return (w, reference_channel) if return_ref_channel else w


def get_bf_vector(beamformer, psd_x, psd_n=None, **bf_kwargs):
    """beamformer_wrapper.py:108-227.  psd_* (F, D, D) -> (F, D)."""
    assert isinstance(beamformer, str) and 'lcmv' not in beamformer, beamformer
    psd_x = np.asarray(psd_x, dtype=np.complex128)
    psd_n = None if psd_n is None else np.asarray(psd_n, dtype=np.complex128)
    ban = beamformer.endswith('+ban')
    core = beamformer[:-len('+ban')] if ban else beamformer
    if core == 'pca':
        w = pca_vector(psd_x, **bf_kwargs)
    elif core in ('pca+mvdr', 'scaled_gev_atf+mvdr'):
        atf = pca_vector(psd_x) if core.startswith('pca') else gev_atf_vector(psd_x, psd_n)
        w = mvdr_vector(atf, psd_n)
    elif core in ('mvdr_souden', 'rank1_pca+mvdr_souden', 'rank1_gev+mvdr_souden'):
        if core != 'mvdr_souden':
            psd_x = rank1_approximation(core.split('+')[0], psd_x, psd_n)
        w = mvdr_souden(psd_x, psd_n, **bf_kwargs)
    elif core in ('gev', 'rank1_pca+gev', 'rank1_gev+gev'):
        if core != 'gev':
            psd_x = rank1_approximation(core.split('+')[0], psd_x, psd_n)
        w = gev_vector(psd_x, psd_n)
    elif core in ('wmwf', 'rank1_pca+wmwf', 'rank1_gev+wmwf'):
        if core != 'wmwf':
            psd_x = rank1_approximation(core.split('+')[0], psd_x, psd_n)
        w = wmwf_vector(psd_x, psd_n, **bf_kwargs)
    elif core.startswith('ch') and core[2:].isdigit():
        w = np.zeros(psd_x.shape[:-1], dtype=np.complex128)
        w[..., int(core[2:])] = 1
    else:
        raise ValueError(f'Could not find implementation for {core}.\nOriginal call contained {beamformer}.')
    if ban:
        w = blind_analytic_normalization(w, psd_n)
    return w
