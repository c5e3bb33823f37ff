"""CPU: the oracle restatement against (a) the golden fixtures produced by the
unmodified reference (oracle/make_golden.py) and (b) the reference's own
known-answer tests for this path."""
import numpy as np
import pytest

from oracle import gss_oracle as oracle


@pytest.mark.parametrize('name', ['gss_d4_k3', 'gss_d8_k4', 'gss_d24_k5'])
def test_gss_and_beamformer_fixture(golden_dir, name):
    g = np.load(golden_dir / f'{name}.npz')
    Obs = g['Obs'].astype(np.complex128)
    post = oracle.gss_posteriors(Obs, g['activity'], int(g['iterations']))
    assert np.abs(post - g['posterior']).max() < 1e-9
    Y = np.transpose(Obs, (2, 0, 1))
    cov_x = oracle.psd_matrix(Y, g['target_mask'].T)
    cov_n = oracle.psd_matrix(Y, g['distortion_mask'].T)
    np.testing.assert_allclose(cov_x, g['cov_x'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(cov_n, g['cov_n'], rtol=1e-10, atol=1e-14)
    w, ref = oracle.mvdr_souden(cov_x, cov_n, eps=1e-10, return_ref_channel=True)
    assert ref == int(g['ref_channel'])
    np.testing.assert_allclose(w, g['w_mvdr'], rtol=1e-8, atol=1e-12)
    X = oracle.beamform(Obs, g['target_mask'], g['distortion_mask'])
    np.testing.assert_allclose(X, g['X_mvdr_ban'], rtol=1e-7, atol=1e-10)
    Xg = oracle.beamform(Obs, g['target_mask'], g['distortion_mask'], bf='gev_ban')
    np.testing.assert_allclose(np.abs(Xg), g['X_gev_ban_abs'], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize('name', ['enh_nowpe', 'enh_wpe'])
def test_whole_path_fixture(golden_dir, name):
    g = np.load(golden_dir / f'{name}.npz')
    taps, delay, its, ctx = (int(v) for v in g['wpe'])
    wpe = dict(taps=taps, delay=delay, iterations=its, psd_context=ctx) if taps else None
    out = oracle.enhance_observation(g['obs'].astype(np.float64), g['sample_activity'], 0,
                                     start_context_samples=int(g['start_context']),
                                     end_context_samples=int(g['end_context']),
                                     wpe=wpe, gss_iterations=10)
    assert np.array_equal(out['activity_freq'], g['activity_freq'])
    assert np.abs(out['masks'] - g['masks']).max() < 1e-6
    assert np.abs(out['x_hat'] - g['x_hat']).max() / np.abs(g['x_hat']).max() < 1e-5


# ---- pb_bss/tests/test_extraction/test_beamformer.py:182-371 -------------------
def _phis():
    obs = np.array([[0, 0, 1], [0, 0.1, 1], [0.1, 0, 1]])
    return obs.T.conj() @ obs, np.eye(3)


def test_mvdr_souden_well_behaviour():
    PhiXX, PhiNN = _phis()
    np.testing.assert_allclose(PhiXX, [[0.01, 0., 0.1], [0., 0.01, 0.1], [0.1, 0.1, 3.]])
    w, = oracle.mvdr_souden(PhiXX[None], PhiNN[None])
    assert repr(w) == 'array([0.03311258, 0.03311258, 0.99337748])', repr(w)
    w3 = oracle.mvdr_souden(np.array([PhiXX] * 3), np.array([PhiNN] * 3))
    np.testing.assert_allclose([w] * 3, w3)


def test_mvdr_souden_difficulties():
    PhiXX, PhiNN = _phis()
    for a, b in [(PhiXX[None] * 0, PhiNN[None]), (PhiXX[None], PhiNN[None] * 0), (PhiXX[None] * 0, PhiNN[None] * 0)]:
        w = oracle.mvdr_souden(a, b)
        assert repr(w) == 'array([[0., 0., 0.]])', repr(w)
    with np.errstate(all='ignore'):
        for a, b in [(PhiXX[None] * np.inf, PhiNN[None]), (PhiXX[None], PhiNN[None] * np.inf)]:
            with pytest.raises((AssertionError, np.linalg.LinAlgError)):
                oracle.mvdr_souden(a, b)


def test_mvdr_souden_eps_multi():
    PhiXX, PhiNN = _phis()
    well, = oracle.mvdr_souden(PhiXX[None], PhiNN[None])
    for a, b in [([PhiXX * 0, PhiXX], [PhiNN, PhiNN]), ([PhiXX, PhiXX], [PhiNN * 0, PhiNN]),
                 ([PhiXX * 0, PhiXX], [PhiNN * 0, PhiNN])]:
        w, ref = oracle.mvdr_souden(np.array(a), np.array(b), return_ref_channel=True)
        assert ref == 2, ref
        np.testing.assert_allclose(w, np.array([[0., 0., 0.], well]))


def test_gev_matches_scipy_by_cosine():
    """test_beamformer.py:120-144 compares implementations by cosine similarity."""
    rng = np.random.default_rng(0)

    def posdef(F, D):
        a = rng.uniform(-1, 1, (F, D, D)) + 1j * rng.uniform(-1, 1, (F, D, D))
        return a @ a.conj().swapaxes(-1, -2) + 0.1 * np.eye(D)

    A, B = posdef(20, 6), posdef(20, 6)
    v = oracle.gev_vector(A, B)
    # v is the principal generalised eigenvector: A v = lambda B v
    Av = np.einsum('fab,fb->fa', A, v)
    Bv = np.einsum('fab,fb->fa', B, v)
    lam = np.einsum('fa,fa->f', v.conj(), Av) / np.einsum('fa,fa->f', v.conj(), Bv)
    np.testing.assert_allclose(Av, lam[:, None] * Bv, atol=1e-8)
    vc = oracle.canonical_phase(v, B)
    cos = np.abs(np.einsum('fd,fd->f', v, vc.conj())) / np.linalg.norm(v, axis=-1) / np.linalg.norm(vc, axis=-1)
    np.testing.assert_allclose(cos, 1.0, atol=1e-6)


# ---- doctests of the reference pinned for the STFT framing ----------------------
def test_stft_framing_doctest():
    """pb_chime5/database/chime5/database.py:417-453."""
    signal = np.array([0, 0, 0, 0, 0, 1, -3, 0, 5, 0, 0, 0, 0, 0])
    vad = np.array([0, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0, 0, 0])
    S = oracle.stft(signal, size=4, shift=2, fading=True, window=np.ones)
    expect = np.array([[0, 0, 0], [0, 0, 0], [1, 1j, -1], [-2, 3 - 1j, -4], [2, -8, 2], [5, 5, 5], [0, 0, 0], [0, 0, 0]])
    np.testing.assert_allclose(S, expect, atol=1e-12)
    a = oracle.activity_time_to_frequency(vad, 4, 2, True)
    assert a.tolist() == [False, False, True, True, True, True, False, False]
    a = oracle.activity_time_to_frequency(vad, 4, 2, False)
    assert a.tolist() == [False, True, True, True, True, False]
    assert oracle.activity_time_to_frequency(np.zeros(200000), 1024, 256, False, False).shape == (778,)
    assert oracle.stft(np.zeros(200000), 1024, 256, fading=False, pad=False).shape == (778, 513)


def test_istft_perfect_reconstruction():
    x = np.random.default_rng(1).standard_normal((2, 5000))
    X = oracle.stft(x)
    assert X.shape == (2, oracle.samples_to_stft_frames(5000, 1024, 256, fading=True), 513)
    xr = oracle.istft(X)
    assert np.abs(xr[:, :5000] - x).max() < 1e-12


def test_stft_istft_match_scipy():
    """nara_wpe.utils.stft / istft are un-vendored (core.py:305-321 only calls them): the oracle's
    periodic Blackman analysis window, the framing and the biorthogonal synthesis are cross-checked
    against scipy.signal's independent STFT / least-squares ISTFT."""
    import warnings
    import scipy.signal
    size, shift = 1024, 256
    rng = np.random.default_rng(3)
    x = rng.standard_normal(5000)
    w = scipy.signal.get_window('blackman', size, fftbins=True)
    xp = np.pad(x, (size - shift, size - shift))
    xp = np.pad(xp, (0, (-(len(xp) - size)) % shift))
    _, _, Z = scipy.signal.stft(xp, window=w, nperseg=size, noverlap=size - shift, boundary=None, padded=False)
    X = oracle.stft(x, size, shift, fading=True)
    assert X.shape == Z.T.shape
    np.testing.assert_allclose(X, Z.T * w.sum(), atol=1e-11)
    # synthesis on spectra that are NOT the STFT of a signal (so that nothing cancels)
    S = rng.standard_normal((23, 513)) + 1j * rng.standard_normal((23, 513))
    S[:, 0] = S[:, 0].real
    S[:, -1] = S[:, -1].real
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')                   # NOLA warning: the edges are cut below
        _, xr = scipy.signal.istft((S / w.sum()).T, window=w, nperseg=size, noverlap=size - shift, boundary=False)
    xo = oracle.istft(S, size, shift, fading=True)
    np.testing.assert_allclose(xo, xr[size - shift:size - shift + len(xo)], atol=1e-13)


def test_wpe_normal_equations():
    """WPE is unpinned by the reference tree; self-check: the filter of the last
    iteration satisfies its normal equations and the output is Y - G^H Yt."""
    rng = np.random.default_rng(2)
    Y = rng.standard_normal((3, 4, 200)) + 1j * rng.standard_normal((3, 4, 200))
    Y[..., 4:] += 0.6 * Y[..., :-4]
    X = oracle.wpe_bins(Y, taps=5, delay=2, iterations=3)
    X2 = oracle.wpe_bins(Y, taps=5, delay=2, iterations=2)
    inv = oracle.wpe_power_inverse(X2)
    Yt = oracle.wpe_tap_matrix(Y, 5, 2)
    R = (Yt * inv[..., None, :]) @ Yt.conj().swapaxes(-1, -2)
    P = (Yt * inv[..., None, :]) @ Y.conj().swapaxes(-1, -2)
    G = np.linalg.solve(R, P)
    np.testing.assert_allclose(X, Y - G.conj().swapaxes(-1, -2) @ Yt, atol=1e-10)
    assert np.mean(np.abs(X) ** 2) < np.mean(np.abs(Y) ** 2)
    # tap matrix layout: row k*D+d at frame t is Y[d, t-delay-k]
    assert np.array_equal(Yt[0, 1 * 4 + 2, 10], Y[0, 2, 10 - 2 - 1])
    assert np.all(Yt[:, :, 0] == 0)


def test_wpe_matches_independent_weighted_least_squares():
    """Second, independent statement of the published WPE iteration (Yoshioka & Nakatani 2012; what
    nara_wpe.wpe_v8 implements): per bin, lambda_t = mean_d |x_d(t)|^2 floored at 1e-10 of its maximum,
    G = argmin_G sum_t |y(t) - G^H ytilde(t)|^2 / lambda_t solved by QR/SVD least squares on the
    whitened design matrix -- no Gram matrix, no normal equations -- written with explicit loops."""
    rng = np.random.default_rng(11)
    F, D, T, taps, delay = 2, 3, 160, 4, 2
    Y = rng.standard_normal((F, D, T)) + 1j * rng.standard_normal((F, D, T))
    Y[..., 3:] += 0.5 * Y[..., :-3]
    Y[..., 7:] -= 0.3j * Y[..., :-7]
    want = np.empty_like(Y)
    for f in range(F):
        x = Y[f].copy()
        for _ in range(3):
            lam = np.mean(np.abs(x) ** 2, axis=0)
            lam = np.maximum(lam, 1e-10 * lam.max())
            A = np.zeros((T, taps * D), complex)                  # row t: ytilde(t)^H / sqrt(lambda_t)
            for t in range(T):
                for k in range(taps):
                    if t - delay - k >= 0:
                        A[t, k * D:(k + 1) * D] = Y[f][:, t - delay - k].conj() / np.sqrt(lam[t])
            B = Y[f].conj().T / np.sqrt(lam)[:, None]             # row t: y(t)^H / sqrt(lambda_t)
            G = np.linalg.lstsq(A, B, rcond=None)[0]
            x = Y[f] - np.stack([sum(G[k * D + e, :].conj() * Y[f][e, t - delay - k]
                                     for k in range(taps) for e in range(D) if t - delay - k >= 0)
                                 if t >= delay else np.zeros(D, complex) for t in range(T)], axis=1)
        want[f] = x
    got = oracle.wpe_bins(Y, taps=taps, delay=delay, iterations=3)
    np.testing.assert_allclose(got, want, atol=1e-9)


def test_bf_vector_dsl_oracle_matches_reference_fixture(golden_dir):
    """get_bf_vector DSL (beamformer_wrapper.py:108-227): the oracle restatement against vectors the
    unmodified reference produced (oracle/make_golden.py) -- same LAPACK calls, so bit for bit."""
    g = np.load(golden_dir / 'bf_dsl_d8.npz')
    cx, cn = g['cov_x'], g['cov_n']
    for key in g.files:
        if key in ('cov_x', 'cov_n'):
            continue
        kw = {}
        name = key.replace('__', '+')
        if key == 'wmwf_mu0p25':
            name, kw = 'wmwf', dict(distortion_weight=0.25)
        elif key == 'wmwf_fd':
            name, kw = 'wmwf', dict(distortion_weight='frequency_dependent')
        elif key in ('pca_trace', 'pca_eigenvalue'):
            name, kw = 'pca', dict(scaling=key[4:])
        w = oracle.get_bf_vector(name, cx, cn, **kw)
        assert w.shape == g[key].shape and np.abs(w - g[key]).max() <= 1e-12 * np.abs(g[key]).max(), key
    with pytest.raises(ValueError):
        oracle.get_bf_vector('music', cx, cn)
    # 'pca+mvdr' / 'scaled_gev_atf+mvdr' go through get_mvdr_vector, which does not run under numpy >= 2
    # in the reference: restated from the formula, distortionless by construction
    for name, atf in (('pca+mvdr', oracle.pca_vector(cx)), ('scaled_gev_atf+mvdr', oracle.gev_atf_vector(cx, cn))):
        w = oracle.get_bf_vector(name, cx, cn)
        assert np.abs(np.einsum('fd,fd->f', w.conj(), atf) - 1).max() < 1e-10
