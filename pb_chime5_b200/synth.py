"""Seeded synthetic CHiME-5-shaped segments.

STFT-domain mixture of K-1 point sources plus spatially white noise; frame
activity per speaker from a two-state Markov chain; last class is the
always-active 'Noise' garbage class (pb_chime5/activity.py:150-156).  The same
complex64 values go to the GPU and (up-cast, bit-identical) to the oracle.
"""
from __future__ import annotations

import numpy as np


def make_activity(rng, K, T, D, p_off=0.01, p_on=0.005, min_active=None):
    """(K, T) bool; rows 0..K-2 speakers, row K-1 'Noise' (all True)."""
    if min_active is None:
        min_active = min(2 * D, T)
    act = np.ones((K, T), dtype=bool)
    for k in range(K - 1):
        while True:
            state = rng.random() < 0.5
            u = rng.random(T)
            row = np.empty(T, dtype=bool)
            for t in range(T):
                if state:
                    state = not (u[t] < p_off)
                else:
                    state = u[t] < p_on
                row[t] = state
            if row.sum() >= min_active:
                break
        act[k] = row
    return act


def make_utterance(seed, D=24, T=941, F=513, K=5, noise_std=0.1,
                   dtype=np.complex64):
    """Returns Obs (D, T, F) complex64 and activity (K, T) bool."""
    rng = np.random.default_rng(seed)
    act = make_activity(rng, K, T, D)
    S = K - 1

    def cn(*shape):
        return (rng.standard_normal(shape, dtype=np.float32)
                + 1j * rng.standard_normal(shape, dtype=np.float32)) * np.float32(0.5 ** 0.5)

    a = cn(F, S, D)
    a /= np.linalg.norm(a, axis=-1, keepdims=True)
    sigma = (0.5 + rng.random(S)).astype(np.float32)
    s = cn(F, S, T) * sigma[None, :, None] * act[None, :S, :]
    obs = np.einsum('fsd,fst->dtf', a, s) + np.float32(noise_std) * cn(D, T, F)
    return np.ascontiguousarray(obs.astype(dtype)), act


def make_batch(seed0, B, **kw):
    obs, act = zip(*(make_utterance(seed0 + b, **kw) for b in range(B)))
    return np.stack(obs), np.stack(act)


def make_audio(seed, D=4, N=48000, K=3, stft_size=1024, stft_shift=256):
    """Raw multichannel audio (D, N) float32 + per-class sample activity
    (K, N) bool, for the STFT -> ... -> iSTFT end-to-end path."""
    rng = np.random.default_rng(seed)
    S = K - 1
    act = np.ones((K, N), dtype=bool)
    seg = max(N // 8, 1)
    for k in range(S):
        on = rng.random(8 + 1) < 0.6
        on[k % 8] = True
        act[k] = np.repeat(on, seg)[:N]
    src = rng.standard_normal((S, N)).astype(np.float32) * act[:S]
    # short random FIR per (source, mic) as a toy room
    h = rng.standard_normal((S, D, 32)).astype(np.float32) * np.exp(-np.arange(32) / 6.0).astype(np.float32)
    obs = np.zeros((D, N), dtype=np.float32)
    for k in range(S):
        for d in range(D):
            obs[d] += np.convolve(src[k], h[k, d])[:N]
    obs += 0.05 * rng.standard_normal((D, N)).astype(np.float32)
    return obs, act


def make_reverberant_audio(seed, D=8, N=64000, K=3, rir_len=2048, t60_taps=900.0, noise=0.02, fast=False):
    """Speech-like (AR(2)-coloured, amplitude-modulated) sources through long, exponentially
    decaying random room responses: strongly time-correlated, reverberant, low-noise input --
    the regime where the WPE normal equations are ill conditioned.  Returns obs (D, N) float32
    and sample activity (K, N) bool (last class = noise).  fast: FFT convolution and a vectorised
    AR filter (same model, different rounding; for the long benchmark signals)."""
    rng = np.random.default_rng(seed)
    S = K - 1
    act = np.ones((K, N), dtype=bool)
    src = np.zeros((S, N), dtype=np.float64)
    for k in range(S):
        e = rng.standard_normal(N)
        x = np.zeros(N)
        a1, a2 = 1.6 - 0.2 * k, -0.8                      # resonant AR(2)
        if fast:
            from scipy.signal import lfilter
            e2 = e.copy()
            e2[:2] = 0
            x = lfilter([1.0], [1.0, -a1, -a2], e2)
        else:
            for n in range(2, N):
                x[n] = a1 * x[n - 1] + a2 * x[n - 2] + e[n]
        env = (np.sin(2 * np.pi * (1.5 + k) * np.arange(N) / 16000.0 + k) > -0.2)
        gate = np.ones(N, dtype=bool)
        seg = N // 6
        off = rng.integers(0, 6)
        gate[off * seg:(off + 1) * seg] = False          # one silent sixth per speaker
        act[k] = gate
        src[k] = x * env * gate / (np.std(x) + 1e-12)
    h = rng.standard_normal((S, D, rir_len)) * np.exp(-np.arange(rir_len) / (t60_taps / 6.9))
    h[:, :, 0] += 3.0                                     # direct path
    obs = np.zeros((D, N))
    if fast:
        from scipy.signal import fftconvolve
        for k in range(S):
            obs += fftconvolve(np.broadcast_to(src[k], (D, N)), h[k], axes=-1)[:, :N]
    else:
        for k in range(S):
            for d in range(D):
                obs[d] += np.convolve(src[k], h[k, d])[:N]
    obs += noise * np.std(obs) * rng.standard_normal((D, N))
    return obs.astype(np.float32), act


# ---------------------------------------------------------------------------
# dev-shaped work lists (BASELINE.json configs[2] / configs[3])
# ---------------------------------------------------------------------------

def make_work_list(seed, n, context_s=15.0, sample_rate=16000, n_bases=4, K=5, median_s=2.0, sigma=0.8, min_s=0.3):
    """n utterance descriptors with dev-shaped durations: LogNormal(median 2 s, sigma 0.8) clipped to
    [min_s = 0.3, 20] s, plus `context_s` seconds of context on both sides (the reference's default
    context_samples=240000, core.py:576).  Plain dicts of ints (cheap to broadcast): the audio itself
    is cut from one of `n_bases` long base recordings every rank synthesises locally
    (`make_base_recording`), starting at `offset`."""
    rng = np.random.default_rng(seed)
    dur = np.clip(rng.lognormal(np.log(median_s), sigma, size=n), min_s, 20.0)
    ctx = int(round(context_s * sample_rate))
    items = []
    for i in range(n):
        ns = int(round(dur[i] * sample_rate))
        items.append(dict(index=i, base=int(rng.integers(n_bases)), offset=int(rng.integers(0, sample_rate)),
                          num_samples_orig=ns, context=ctx, total=ns + 2 * ctx, target=int(rng.integers(K - 1))))
    return items


def make_base_recording(seed, D=24, seconds=51.0, K=5, sample_rate=16000, noise_std=0.1):
    """One long synthetic multichannel recording (D, N) float32 + per-class sample activity
    (K, N) bool (last class = always-on noise): K-1 white sources with on/off activity in blocks of
    0.25-3 s, instantaneous random mixing, white sensor noise."""
    rng = np.random.default_rng(seed)
    N = int(seconds * sample_rate)
    S = K - 1
    act = np.ones((K, N), dtype=bool)
    for k in range(S):
        t, on = 0, bool(rng.random() < 0.5)
        row = np.zeros(N, dtype=bool)
        while t < N:
            n = int(rng.uniform(0.25, 3.0) * sample_rate)
            row[t:t + n] = on
            on = not on
            t += n
        act[k] = row
    src = rng.standard_normal((S, N), dtype=np.float32) * act[:S]
    mix = rng.standard_normal((D, S)).astype(np.float32)
    mix /= np.linalg.norm(mix, axis=0, keepdims=True)
    obs = mix @ src + np.float32(noise_std) * rng.standard_normal((D, N), dtype=np.float32)
    return np.ascontiguousarray(obs.astype(np.float32)), act


def work_item_example(item, speakers):
    """example dict in the CHiME-5 json layout (the keys Enhancer / SessionScheduler read) for a
    work-list item"""
    ctx, ns, tot = item['context'], item['num_samples_orig'], item['total']
    return {'example_id': f"utt{item['index']:06d}", 'session_id': 'S02', 'reference_array': 'U01',
            'speaker_id': speakers[item['target']],
            'start': {'original': 0, 'observation': {'U01': 0}}, 'end': {'original': tot, 'observation': {'U01': tot}},
            'start_orig': {'original': ctx, 'observation': {'U01': ctx}},
            'end_orig': {'original': ctx + ns, 'observation': {'U01': ctx + ns}},
            'num_samples_orig': {'original': ns, 'observation': {'U01': ns}}, 'item': item}
