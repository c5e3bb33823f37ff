"""GPU parity: the CUDA path (through the C ABI) against the golden fixtures
produced by the unmodified reference and against the CPU oracle on seeded
inputs.  Tolerances: north_star says 1e-4 relative on float32 outputs."""
import numpy as np
import pytest
import torch

from oracle import gss_oracle as oracle
from pb_chime5_b200 import ops, synth

pytestmark = pytest.mark.gpu

MASK_ATOL = 1e-4          # posteriors live in [0, 1]
REL_TOL = 1e-4            # beamformed spectra, relative to the largest magnitude


def to_fdt(obs_dtf, dev):
    x = torch.from_numpy(np.ascontiguousarray(obs_dtf)).to(dev)
    return ops.pack_dtf_to_fdt(x[None])


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('name', ['gss_d4_k3', 'gss_d8_k4', 'gss_d24_k5'])
def test_cacgmm_matches_reference_fixture(cuda, golden_dir, name):
    g = np.load(golden_dir / f'{name}.npz')
    Y = to_fdt(g['Obs'], cuda)
    act = torch.from_numpy(g['activity'])[None].to(cuda)
    post = ops.cacgmm(Y, act, int(g['iterations']))
    post = ops.unpack_fkt_to_ktf(post)[0].cpu().numpy()          # (K,T,F)
    assert post.shape == g['posterior'].shape
    err = np.abs(post - g['posterior']).max()
    assert err < MASK_ATOL, err


@pytest.mark.parametrize('name', ['gss_d4_k3', 'gss_d8_k4', 'gss_d24_k5'])
@pytest.mark.parametrize('bf', ['mvdrSouden_ban', 'gev_ban'])
def test_beamformer_matches_reference_fixture(cuda, golden_dir, name, bf):
    g = np.load(golden_dir / f'{name}.npz')
    Y = to_fdt(g['Obs'], cuda)
    tm = torch.from_numpy(g['target_mask'].T.astype(np.float32).copy())[None].to(cuda)    # (1,F,T)
    dm = torch.from_numpy(g['distortion_mask'].T.astype(np.float32).copy())[None].to(cuda)
    X, aux = ops.beamform(Y, tm, dm, bf=bf, return_aux=True)
    X = ops.unpack_ft_to_tf(X)[0].cpu().numpy()
    if bf == 'mvdrSouden_ban':
        assert int(aux['ref_channel'][0]) == int(g['ref_channel'])
        assert rel_err(X, g['X_mvdr_ban']) < REL_TOL
        w = aux['weights'][0].cpu().numpy()
        assert rel_err(w, g['w_mvdr_ban']) < REL_TOL
    else:
        assert rel_err(np.abs(X), g['X_gev_ban_abs']) < REL_TOL


def test_mvdr_known_answer(cuda):
    """pb_bss/tests/test_extraction/test_beamformer.py:182-204."""
    obs = np.array([[0, 0, 1], [0, 0.1, 1], [0.1, 0, 1]])
    # one bin, three frames with equal weights reproduce PhiXX up to scale; use
    # the covariance path directly: Phi_X = obs^H obs, Phi_N = I via masks is not
    # expressible, so check through the oracle on a synthetic observation instead.
    Obs, act = synth.make_utterance(3, D=3, T=64, F=4, K=3)
    rng = np.random.default_rng(0)
    tm = rng.random((64, 4)); dm = rng.random((64, 4))
    ref = oracle.beamform(Obs.astype(np.complex128), tm.astype(np.float32).astype(np.float64),
                          dm.astype(np.float32).astype(np.float64))
    Y = to_fdt(Obs, cuda)
    X = ops.beamform(Y, torch.from_numpy(tm.T.astype(np.float32).copy())[None].to(cuda),
                     torch.from_numpy(dm.T.astype(np.float32).copy())[None].to(cuda))
    X = ops.unpack_ft_to_tf(X)[0].cpu().numpy()
    assert rel_err(X, ref) < REL_TOL


@pytest.mark.parametrize('D,T,F,taps,delay', [(4, 150, 5, 4, 2), (8, 200, 3, 10, 3), (24, 300, 2, 10, 2)])
def test_wpe_matches_oracle(cuda, D, T, F, taps, delay):
    Obs, _ = synth.make_utterance(11, D=D, T=T, F=F, K=3)
    # add some reverberation-like temporal correlation
    Obs[:, 3:, :] += 0.5 * Obs[:, :-3, :]
    Obs[:, 5:, :] += 0.25 * Obs[:, :-5, :]
    ref = oracle.wpe_dtf(Obs.astype(np.complex128), taps, delay, 3)
    Y = to_fdt(Obs, cuda)
    X = ops.wpe(Y, taps, delay, 3)
    X = ops.unpack_fdt_to_dtf(X)[0].cpu().numpy()
    assert rel_err(X, ref) < REL_TOL


def test_weighted_cov_matches_oracle(cuda):
    Obs, _ = synth.make_utterance(5, D=24, T=333, F=7, K=3)
    rng = np.random.default_rng(1)
    w = rng.random((1, 7, 3, 333)).astype(np.float32)
    Y = to_fdt(Obs, cuda)
    Phi = ops.weighted_cov(Y, torch.from_numpy(w).to(cuda), normalize=True)[0].cpu().numpy()
    Yn = np.transpose(Obs.astype(np.complex128), (2, 0, 1))
    for k in range(3):
        ref = oracle.psd_matrix(Yn, w[0, :, k].astype(np.float64))
        assert rel_err(Phi[:, k], ref) < 1e-5


def test_layout_roundtrip(cuda):
    x = torch.randn(2, 3, 50, 17, dtype=torch.complex64, device=cuda)
    y = ops.pack_dtf_to_fdt(x)
    assert torch.equal(y, x.permute(0, 3, 1, 2).contiguous())
    assert torch.equal(ops.unpack_fdt_to_dtf(y), x)
    m = torch.rand(2, 17, 4, 50, device=cuda)
    assert torch.equal(ops.unpack_fkt_to_ktf(m), m.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.pack_ktf_to_fkt(ops.unpack_fkt_to_ktf(m)), m)


@pytest.mark.parametrize('N,size,shift,fading', [(5000, 1024, 256, True), (20000, 1024, 256, True),
                                                  (4096, 512, 128, False), (700, 1024, 256, True)])
def test_stft_istft_match_oracle(cuda, N, size, shift, fading):
    rng = np.random.default_rng(N)
    x = rng.standard_normal((2, 3, N)).astype(np.float32)
    ref = oracle.stft(x.astype(np.float64), size, shift, fading)          # (B,D,T,F)
    Y = ops.stft(torch.from_numpy(x).to(cuda), size, shift, fading)       # (B,F,D,T)
    got = Y.permute(0, 2, 3, 1).cpu().numpy()
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-6
    # iSTFT of channel 0
    X = Y[:, :, 0, :].contiguous()
    back = ops.istft(X, size, shift, fading).cpu().numpy()
    ref_back = oracle.istft(ref[:, 0], size, shift, fading)
    assert back.shape == ref_back.shape
    assert rel_err(back, ref_back) < 1e-5
    if fading:
        assert np.abs(back[:, :N] - x[:, 0]).max() < 1e-4      # perfect reconstruction


def _dsl_cases(g):
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
        yield key, name, kw


def _same_up_to_phase(w, ref, tol=1e-8):
    """per-bin vectors equal up to a unit-modulus factor (the reference compares eigenvector based
    beamformers by cosine similarity, test_beamformer.py:17-21) and in norm"""
    num = np.abs(np.einsum('...d,...d->...', ref.conj(), w))
    nw, nr = np.linalg.norm(w, axis=-1), np.linalg.norm(ref, axis=-1)
    return float(np.abs(1 - num / np.maximum(nw * nr, 1e-300)).max()) < tol and float(np.abs(nw / nr - 1).max()) < tol


def test_get_bf_vector_dsl_matches_reference_fixture(cuda, golden_dir):
    """row f4: every `get_bf_vector` string on the device against the vectors of the unmodified reference
    (tests/golden/bf_dsl_d8.npz).  Closed-form vectors must agree entry by entry; eigenvector-based ones
    (pca, gev, and rank-1 models feeding gev) up to the phase LAPACK leaves open."""
    from pb_chime5_b200.extraction import get_bf_vector
    torch.cuda.set_device(cuda)
    g = np.load(golden_dir / 'bf_dsl_d8.npz')
    cx, cn = g['cov_x'], g['cov_n']
    for key, name, kw in _dsl_cases(g):
        w = get_bf_vector(name, cx, cn, **kw)
        ref = g[key]
        assert w.shape == ref.shape and w.dtype == np.complex128, key
        phase_free = ('gev' in name and 'mvdr_souden' not in name and 'wmwf' not in name) or name.startswith('pca')
        if phase_free:
            assert _same_up_to_phase(w, ref), key
        else:
            assert np.abs(w - ref).max() <= 1e-8 * np.abs(ref).max(), (key, np.abs(w - ref).max())
    # ATF-based MVDR (not runnable in the reference under numpy 2): against the oracle's formula
    for name in ('pca+mvdr', 'scaled_gev_atf+mvdr', 'scaled_gev_atf+mvdr+ban'):
        assert _same_up_to_phase(get_bf_vector(name, cx, cn), oracle.get_bf_vector(name, cx, cn)), name
    # utterance batches: the reference channel is chosen per utterance; torch in -> torch out
    cxb = torch.from_numpy(np.stack([cx, cx[::-1].copy()])).cuda()
    cnb = torch.from_numpy(np.stack([cn, cn[::-1].copy()])).cuda()
    wb, refs = get_bf_vector('mvdr_souden', cxb, cnb, return_ref_channel=True)
    assert wb.is_cuda and wb.shape == (2,) + cx.shape[:-1]
    for b, (x, n) in enumerate(((cx, cn), (cx[::-1], cn[::-1]))):
        wo, ro = oracle.mvdr_souden(x, n, return_ref_channel=True)
        assert int(refs[b]) == ro and np.abs(wb[b].cpu().numpy() - wo).max() <= 1e-8 * np.abs(wo).max()
    with pytest.raises(ValueError):
        get_bf_vector('music', cx, cn)
    with pytest.raises(AssertionError):
        get_bf_vector('lcmv', cx, cn)
    with pytest.raises(AssertionError):
        get_bf_vector('mvdr_souden', cx, None)


@pytest.mark.parametrize('bf', ['rank1_gev+mvdr_souden+ban', 'wmwf+ban', 'rank1_pca+gev+ban', 'pca+ban', 'ch1'])
def test_beamformer_block_accepts_dsl_types(cuda, golden_dir, bf):
    """extension: core.Beamformer(type=<get_bf_vector string>) = PSD matrices + DSL vector + w^H y on the
    device, against the oracle chain (magnitudes for the eigenvector based ones)"""
    from pb_chime5_b200 import core
    torch.cuda.set_device(cuda)
    g = np.load(golden_dir / 'gss_d8_k4.npz')
    Obs = g['Obs'].astype(np.complex128)
    tm, dm = g['target_mask'], g['distortion_mask']
    X = core.Beamformer(bf, None)(Obs, tm, dm)
    Y = np.transpose(Obs, (2, 0, 1))
    w = oracle.get_bf_vector(bf, oracle.psd_matrix(Y, tm.T), oracle.psd_matrix(Y, dm.T))
    refX = oracle.apply_beamforming_vector(w, Y).T
    assert X.shape == refX.shape
    if 'gev' in bf.split('+')[-2:] or bf.startswith('pca'):
        assert np.abs(np.abs(X) - np.abs(refX)).max() < 1e-4 * np.abs(refX).max()
    else:
        assert np.abs(X - refX).max() < 1e-4 * np.abs(refX).max()
