"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/refboot.py) on seeded synthetic
inputs.  Run here (CPU container) only:  python -m oracle.make_golden

The fixtures store the exact complex64 / float32 inputs and the reference's
float64 outputs, so the GPU box (no /root/reference) can check both the oracle
and the CUDA path against the reference itself.
"""
import warnings
from pathlib import Path

import numpy as np

from oracle import refboot

OUT = Path(__file__).resolve().parent.parent / 'tests' / 'golden'

# DSL strings the reference evaluates under numpy 2 ('pca+mvdr' / 'scaled_gev_atf+mvdr' go through
# get_mvdr_vector, whose np.linalg.solve(a, vector-stack) call breaks on numpy >= 2.0: SURVEY appendix B)
BF_DSL_NAMES = ['mvdr_souden', 'mvdr_souden+ban', 'rank1_pca+mvdr_souden', 'rank1_gev+mvdr_souden',
                'rank1_gev+mvdr_souden+ban', 'gev', 'gev+ban', 'rank1_pca+gev', 'rank1_pca+gev+ban',
                'wmwf', 'wmwf+ban', 'rank1_pca+wmwf', 'rank1_gev+wmwf', 'pca', 'pca+ban', 'ch2', 'ch1+ban']


def main():
    warnings.filterwarnings('ignore')
    refboot.boot()
    import pb_chime5.core as core
    from pb_chime5.speech_enhancement.beamforming_wrapper import (
        beamform_gev_from_masks, _Beamformer)
    from pb_bss.extraction import beamformer as rbf
    from pb_chime5_b200 import synth

    OUT.mkdir(parents=True, exist_ok=True)

    # ---- GSS + beamformer on STFT input -------------------------------
    for name, kw, iters in (
        ('gss_d4_k3', dict(D=4, T=150, F=9, K=3), 20),
        ('gss_d8_k4', dict(D=8, T=220, F=5, K=4), 12),
        ('gss_d24_k5', dict(D=24, T=300, F=3, K=5), 15),
    ):
        Obs32, act = synth.make_utterance(100 + iters, **kw)
        Obs = Obs32.astype(np.complex128)
        post = core.GSS(iterations=iters, iterations_post=1, verbose=False)(Obs, act)
        masks = post.copy()
        masks[:, :3, :] = 0
        masks[:, -3:, :] = 0
        tm = masks[0]
        dm = masks[1:].sum(0)
        bf = _Beamformer(Obs, tm, dm)
        w_mvdr, ref_ch = rbf.get_mvdr_vector_souden(bf._Cov_X, bf._Cov_N, eps=1e-10,
                                                    return_ref_channel=True)
        X_mvdr = core.Beamformer('mvdrSouden_ban', None)(Obs, tm, dm)
        X_gev = beamform_gev_from_masks(Obs, tm, dm, ban=True)
        np.savez_compressed(
            OUT / f'{name}.npz', Obs=Obs32, activity=act, iterations=iters,
            posterior=post, target_mask=tm, distortion_mask=dm,
            cov_x=bf._Cov_X, cov_n=bf._Cov_N, w_mvdr=w_mvdr, ref_channel=ref_ch,
            w_mvdr_ban=bf._w_mvdr_souden_ban, X_mvdr_ban=X_mvdr,
            X_gev_ban_abs=np.abs(X_gev))
        print(name, post.shape, 'ref_channel', ref_ch)

    # ---- get_bf_vector DSL (beamformer_wrapper.py:108-227) on the PSD matrices of a fixture ----
    from pb_bss.extraction.beamformer_wrapper import get_bf_vector
    g = np.load(OUT / 'gss_d8_k4.npz')
    dsl = {}
    for name in BF_DSL_NAMES:
        dsl[name.replace('+', '__')] = get_bf_vector(name, g['cov_x'].copy(), g['cov_n'].copy())
    dsl['wmwf_mu0p25'] = get_bf_vector('wmwf', g['cov_x'].copy(), g['cov_n'].copy(), distortion_weight=0.25)
    dsl['wmwf_fd'] = get_bf_vector('wmwf', g['cov_x'].copy(), g['cov_n'].copy(), distortion_weight='frequency_dependent')
    dsl['pca_trace'] = get_bf_vector('pca', g['cov_x'].copy(), g['cov_n'].copy(), scaling='trace')
    dsl['pca_eigenvalue'] = get_bf_vector('pca', g['cov_x'].copy(), g['cov_n'].copy(), scaling='eigenvalue')
    np.savez_compressed(OUT / 'bf_dsl_d8.npz', cov_x=g['cov_x'], cov_n=g['cov_n'], **dsl)
    print('bf_dsl_d8', sorted(dsl))

    # ---- whole path on raw audio through Enhancer.enhance_observation ---
    for name, wpe in (('enh_nowpe', None), ('enh_wpe', dict(taps=4, delay=2, iterations=3, psd_context=0))):
        obs, sact = synth.make_audio(5, D=4, N=20000, K=3)
        enh = core.Enhancer(
            wpe_block=core.WPE(**wpe) if wpe else None, activity=None,
            gss_block=core.GSS(iterations=10, iterations_post=1, verbose=False),
            bf_block=core.Beamformer(type='mvdrSouden_ban', postfilter=None),
            bf_drop_context=True, stft_size=1024, stft_shift=256, stft_fading=True,
            context_samples=4000, multiarray=False, reference_array='U01')
        ex = {'start': {'original': 0}, 'start_orig': {'original': 4000},
              'end': {'original': 20000}, 'end_orig': {'original': 16000}}
        ex_act = {'P01': sact[0], 'P02': sact[1], 'Noise': sact[2]}
        x_hat = enh.enhance_observation(obs.astype(np.float64), ex_act, 'P01', ex=ex,
                                        debug=True)
        loc = enh.enhance_observation_locals
        np.savez_compressed(
            OUT / f'{name}.npz', obs=obs, sample_activity=sact,
            wpe=np.array([wpe[k] for k in ('taps', 'delay', 'iterations', 'psd_context')]
                         if wpe else [0, 0, 0, 0]),
            activity_freq=loc['acitivity_freq'], masks=loc['masks'].astype(np.float32),
            X_hat=loc['X_hat'].astype(np.complex64), x_hat=x_hat.astype(np.float32),
            start_context=4000, end_context=4000)
        print(name, x_hat.shape)


if __name__ == '__main__':
    main()
