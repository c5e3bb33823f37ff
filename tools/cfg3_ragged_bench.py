"""Side measurement for BASELINE configs[2]/[3] ("dev-shaped" utterances): a ragged batch of
24-channel utterances whose lengths follow the SURVEY 8d recipe (utterance ~ LogNormal(median
2.0 s, sigma 0.8) clipped to [0.3, 20] s, plus `context` seconds on both sides), full
WPE + GSS + GEV(+BAN), through Enhancer.enhance_stft_batch with per-utterance frame counts.
Prints one JSON line (utterances/s, processed audio seconds per second).  Not the bench metric.
    python tools/cfg3_ragged_bench.py [--batch 16] [--context 15] [--steps 2]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import core, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--context', type=float, default=15.0, help='seconds of context on each side (reference default: 15)')
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--bf', default='gev_ban')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(3)
    D, F, K = 24, 513, 5
    dur = np.clip(rng.lognormal(np.log(2.0), 0.8, size=a.batch), 0.3, 20.0)
    total = dur + 2 * a.context
    frames = sorted((int(np.ceil((t * 16000 + 2 * 768 - 1024 + 256) / 256)) for t in total), reverse=True)
    Tmax = frames[0]
    ctx = int(np.ceil((a.context * 16000 + 768) / 256)) if a.context > 0 else 0
    enh = core.get_enhancer(wpe_tabs=10, wpe_delay=2, wpe_iterations=3, bss_iterations=20, bf=a.bf)
    Y = torch.zeros((a.batch, F, D, Tmax), dtype=torch.complex64, device=dev)
    A = torch.zeros((a.batch, K, Tmax), dtype=torch.uint8, device=dev)
    for b, T in enumerate(frames):
        obs, act = synth.make_utterance(5000 + b, D=D, T=T, F=F, K=K)
        Y[b, :, :, :T] = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)[None])[0]
        A[b, :, :T] = torch.from_numpy(act.astype(np.uint8)).to(dev)
    iv = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)   # noqa: E731
    ti, sc, fr = iv([0] * a.batch), iv([min(ctx, T // 3) for T in frames]), iv(frames)

    def step():
        return enh.enhance_stft_batch(Y, A, ti, sc, sc, frames=fr)
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        X = step()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / a.steps
    assert bool(torch.isfinite(torch.view_as_real(X)).all())
    print(json.dumps({'workload': f'cfg3-like: {a.batch} dev-shaped utterances, D=24, K=5, context {a.context} s per side, '
                                  f'WPE 10/2/3 + 20 EM iterations + {a.bf}',
                      'frames_min_median_max': [frames[-1], int(np.median(frames)), Tmax],
                      'padding_efficiency': float(sum(frames) / (a.batch * Tmax)),
                      'seconds_per_batch': sec, 'utterances_per_s': a.batch / sec,
                      'audio_seconds_per_s': float(total.sum() / sec),
                      'enhanced_speech_seconds_per_s': float(dur.sum() / sec)}))


if __name__ == '__main__':
    main()
