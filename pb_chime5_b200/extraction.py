"""`get_bf_vector`: the beamformer DSL of pb_bss (pb_bss/extraction/beamformer_wrapper.py:108-227)
on the device.  Same strings, same keyword arguments:

    'mvdr_souden', 'mvdr_souden+ban', 'rank1_pca+mvdr_souden', 'rank1_gev+mvdr_souden+ban',
    'gev', 'gev+ban', 'rank1_pca+gev', 'wmwf', 'rank1_gev+wmwf', 'pca', 'pca+mvdr',
    'scaled_gev_atf+mvdr', 'ch3', ...

Every step is a libgss kernel (`gss_bf_vector_c128`): rank-1 models (principal eigenvector / GEV based
ATF estimate), the MVDR-Souden / weighted-MWF solve with the SNR-optimal reference channel, the
generalised eigenvector, the ATF-based MVDR and the blind analytic normalisation.  NumPy in ->
NumPy out (complex128), torch CUDA in -> CUDA out.  There is no CPU path.

Eigenvector phases (which LAPACK leaves unspecified) are fixed: 'pca' vectors have their largest
component real and positive, 'gev' vectors make (Phi_N w)[0] real and non-negative; compare with
the reference through |w^H y| or the cosine similarity, as the reference's own tests do
(pb_bss/tests/test_extraction/test_beamformer.py:17-21).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops

__all__ = ['get_bf_vector', 'parse_beamformer']


def parse_beamformer(beamformer):
    """'rank1_gev+mvdr_souden+ban' -> dict(rank1='rank1_gev', core='mvdr_souden', ban=True, channel=0).
    Raises like the reference: AssertionError for lcmv / non-strings, ValueError for unknown names."""
    assert isinstance(beamformer, str), beamformer
    assert 'lcmv' not in beamformer, (
        'Since the LCMV beamformer and its variants sufficiently differ from all other beamforming '
        'approaches, the reference provides a separate wrapper function `get_multi_source_bf_vector()`.')
    ban = beamformer.endswith('+ban')
    core = beamformer[:-len('+ban')] if ban else beamformer
    rank1, channel = None, 0
    if core in ('pca', 'pca+mvdr', 'scaled_gev_atf+mvdr', 'mvdr_souden', 'gev', 'wmwf'):
        pass
    elif core in ('rank1_pca+mvdr_souden', 'rank1_gev+mvdr_souden', 'rank1_pca+gev', 'rank1_gev+gev',
                  'rank1_pca+wmwf', 'rank1_gev+wmwf'):
        rank1, core = core.split('+')
    elif 'ch' in core and core[2:].isdigit():
        channel, core = int(core[2:]), 'ch'
    else:
        raise ValueError(f'Could not find implementation for {core}.\nOriginal call contained {beamformer}.')
    return dict(rank1=rank1, core=core, ban=ban, channel=channel)


def get_bf_vector(beamformer, target_psd_matrix, noise_psd_matrix=None, **bf_kwargs):
    """Beamforming vectors (..., F, D) from PSD matrices (..., F, D, D); the leading axes are
    independent utterances (the SNR-optimal reference channel of mvdr_souden / wmwf is chosen per
    utterance over its F bins, beamformer.py:524-543).  bf_kwargs as in the reference:
    `ref_channel` / `eps` (mvdr_souden), `reference_channel` / `distortion_weight` (wmwf),
    `scaling` (pca); `return_ref_channel=True` also returns the chosen channel(s)."""
    prog = parse_beamformer(beamformer)
    core = prog['core']
    was_np = not isinstance(target_psd_matrix, torch.Tensor)
    dev = torch.device('cuda', torch.cuda.current_device())
    X = torch.as_tensor(target_psd_matrix).to(device=dev, dtype=torch.complex128).contiguous()
    assert X.ndim >= 3 and X.shape[-1] == X.shape[-2], X.shape
    lead, (F, D) = X.shape[:-3], X.shape[-3:-1]
    B = int(np.prod(lead)) if lead else 1
    N = None
    if noise_psd_matrix is not None:
        N = torch.as_tensor(noise_psd_matrix).to(device=dev, dtype=torch.complex128).contiguous()
        assert N.shape == X.shape, (N.shape, X.shape)
    elif prog['ban'] or prog['rank1'] == 'rank1_gev' or core not in ('pca', 'ch'):
        raise AssertionError('noise_psd_matrix is None')            # the reference asserts the same
    kw = dict(bf_kwargs)
    kw.pop('atf_kwargs', None)
    ref, mu, scaling = -1, 1.0, None
    return_ref = bool(kw.pop('return_ref_channel', False))
    if core == 'mvdr_souden':
        r = kw.pop('ref_channel', None)
        ref = -1 if r is None else int(r)
        eps = kw.pop('eps', None)
        assert eps is None, 'eps: only the reference default (float64 tiny) is built'
    elif core == 'wmwf':
        r = kw.pop('reference_channel', None)
        ref = -1 if r is None else int(r)
        mu = kw.pop('distortion_weight', 1.0)
        mu = -1.0 if mu == 'frequency_dependent' else float(mu)
        assert kw.pop('channel_selection_vector', None) is None, 'channel_selection_vector is not built'
    elif core == 'pca':
        scaling = kw.pop('scaling', None)
        if scaling not in _lib.PCA_SCALING:
            raise ValueError(scaling)
    assert not kw, f'unexpected keyword arguments {sorted(kw)}'
    w = torch.empty((B, F, D), dtype=torch.complex128, device=dev)
    ref_out = torch.full((max(B, 1),), ref, dtype=torch.int32, device=dev)
    info = ops.new_info(B, dev)
    ws = ops.workspace(_lib.workspace_bytes(_lib.OP_BF_VECTOR, B, F, D, 1, 1, 0), dev)
    _lib.check(_lib.lib().gss_bf_vector_c128(
        ops._ptr(X), ops._ptr(N), ops._ptr(w), _lib.BF_CORES[core], _lib.BF_RANK1[prog['rank1']], int(prog['ban']),
        ref, mu, _lib.PCA_SCALING[scaling], prog['channel'], B, F, D, ops._ptr(ref_out), ops._ptr(info),
        ops._ptr(ws), ws.numel(), ops._stream()))
    ops.check_info(info, f'get_bf_vector({beamformer!r})')
    w = w.reshape(*lead, F, D)
    out = w.cpu().numpy() if was_np else w
    if return_ref:
        r = ref_out.reshape(lead) if lead else ref_out[0]
        return out, (r.cpu().numpy() if was_np and lead else (int(r) if not lead else r))
    return out
