#!/usr/bin/env python
"""Benchmark of the Enhancer hot path (WPE -> guided CACGMM EM -> MVDR-Souden+BAN).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): utterances/sec for 15 s, 24-channel, 513-bin STFT
segments.  Workload = BASELINE.json configs[1] ("cfg2"): D=24, T=941, F=513,
K=5 classes, WPE taps=10 delay=2 iterations=3, 100 EM iterations, MVDR-Souden
+ BAN -- `--batch` such utterances per GPU and step (default 8).

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the meaning
of every field.  A "step" is one pass of the hot path over one batch.
"""
import os

os.environ.setdefault('OMP_NUM_THREADS', '1')      # the reference pins BLAS to 1 thread (pb_chime5/__init__.py:4-14)
os.environ.setdefault('MKL_NUM_THREADS', '1')
os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')

import argparse
import json
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # BASELINE.json configs[1] -- the configuration the metric is quoted on (default)
    'cfg2': (dict(D=24, T=941, F=513, K=5, taps=10, delay=2, wpe_iterations=3, em_iterations=100,
                  bf='mvdrSouden_ban', ctx_frames=3),
             'cfg2: 15 s utterances, 6 arrays x 4 mics (D=24), T=941 frames, F=513 bins, K=5 classes, '
             'WPE taps=10 delay=2 it=3 + CACGMM 100 EM it + MVDR-Souden+BAN'),
    # configs[0] -- the reference's own CPU-runnable case (1 array, WPE off, 20 EM iterations)
    'cfg1': (dict(D=4, T=941, F=513, K=3, taps=0, delay=0, wpe_iterations=0, em_iterations=20,
                  bf='mvdrSouden_ban', ctx_frames=3),
             'cfg1: 15 s utterances, 1 array x 4 mics (D=4), T=941, F=513, K=3, WPE off, '
             'CACGMM 20 EM it + MVDR-Souden+BAN'),
    # configs[4] -- stress shape
    'cfg5': (dict(D=24, T=3753, F=513, K=6, taps=20, delay=2, wpe_iterations=3, em_iterations=200,
                  bf='mvdrSouden_ban', ctx_frames=3),
             'cfg5: 60 s utterances, D=24, T=3753, F=513, K=6 (5 speakers + noise), WPE taps=20 it=3 + '
             'CACGMM 200 EM it + MVDR-Souden+BAN'),
}
# configs[2] / configs[3]: dev-shaped ragged work lists through the session driver (raw audio in, audio out)
SESSION_WORKLOADS = {
    'cfg3': (dict(D=24, K=5, F=513, taps=10, delay=2, wpe_iterations=3, em_iterations=20, bf='gev_ban',
                  n_utts=512, context_s=15.0, batch_size=8, min_s=0.5),
             'cfg3: 512 dev-shaped utterances (LogNormal(2 s, 0.8) in [0.5, 20] s -- GEV needs more utterance frames than '
             'channels, zhegvd fails on a singular noise PSD in the reference too -- + 15 s context per side = 30-50 s '
             'segments, T ~ 1900-3100 frames), D=24, K=5, WPE taps=10 it=3 + CACGMM 20 EM it + GEV+BAN, '
             'raw audio -> STFT -> ... -> iSTFT through the session driver on one GPU'),
    'cfg4': (dict(D=24, K=5, F=513, taps=10, delay=2, wpe_iterations=3, em_iterations=20, bf='mvdrSouden_ban',
                  n_utts=20000, context_s=15.0, batch_size=8, min_s=0.3),
             'cfg4: 20 000 dev-shaped utterances (as cfg3, default context_samples=240000) sharded over the '
             'GPUs of one box by the task-farm work queue, reference defaults (20 EM it, MVDR-Souden+BAN)'),
}
CFG, WORKLOAD = WORKLOADS['cfg2']
METRIC = 'utterances/sec (15 s, 24-ch, 513-bin STFT)'


# ---------------------------------------------------------------------------
# CPU side: the oracle port timed like the reference runs (one process per core,
# one BLAS thread each, a Python loop over frequency bins)
# ---------------------------------------------------------------------------

def _cpu_worker(args):
    seed, bins = args
    from oracle import gss_oracle as oracle
    from pb_chime5_b200 import synth
    c = CFG
    Obs, act = synth.make_utterance(seed, D=c['D'], T=c['T'], F=bins, K=c['K'])
    Obs = Obs.astype(np.complex128)
    t0 = time.perf_counter()
    oracle.enhance_stft(Obs, act, 0,
                        wpe=(dict(taps=c['taps'], delay=c['delay'], iterations=c['wpe_iterations'], psd_context=0)
                             if c['taps'] else None),
                        gss_iterations=c['em_iterations'], bf=c['bf'],
                        start_context_frames=c['ctx_frames'], end_context_frames=c['ctx_frames'],
                        loop_over_bins=True)
    return time.perf_counter() - t0


def cpu_sample(pool, procs, bins_per_proc, seed0):
    """One bounded CPU sample: every process enhances `bins_per_proc` bins of one
    utterance over ALL iterations.  Returns utterances/sec of the whole host."""
    times = pool.map(_cpu_worker, [(seed0 + i, bins_per_proc) for i in range(procs)])
    wall = max(times)                 # synthetic-data generation is outside the workers' timers
    return procs * bins_per_proc / CFG['F'] / wall, wall


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python
    reference cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    import multiprocessing as mp
    procs = os.cpu_count() or 1
    bins = args.cpu_bins or 4
    ctx = mp.get_context('fork')
    with ctx.Pool(procs) as pool:
        for i in range(args.warmup):
            cpu_sample(pool, procs, bins, 10_000 + 100 * i)
        wall = 0.0
        for i in range(args.steps):
            wall += cpu_sample(pool, procs, bins, 20_000 + 100 * i)[1]
    value = args.steps * procs * bins / CFG['F'] / wall
    sample = (f'{procs} processes x {bins} of 513 bins per step, all iterations of WPE+CACGMM+MVDR, '
              f'scaled by 513/{bins}; per-bin Python loop as pb_chime5/core.py:172')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'utterances/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * wall / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': None},
        'cpu_baseline': {'value': value, 'unit': 'utterances/s', 'cores': procs, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'utterances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------

class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.thread, self.index = [], None, None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            p = [x.strip() for x in r.split(',')]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return None
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------

def algorithmic_bytes_em(B):
    """SURVEY.md section 8d, per launch of the fused EM kernel: per-iteration
    covariance + E-step streams x iterations (c = 8 B complex64, r = 4 B float32)."""
    c, r = 8, 4
    D, T, F, K, it = CFG['D'], CFG['T'], CFG['F'], CFG['K'], CFG['em_iterations']
    cov = F * (D * T * c + 2 * K * T * r + K * D * D * c)
    est = F * (D * T * c + K * D * D * c + K * D * r + K * T * 1 + 2 * K * T * r)
    fused_min = F * D * T * c + K * T + F * K * T * r
    return B * it * (cov + est), B * fused_min


def algorithmic_flops_em(B):
    """SURVEY.md section 8d flop counts of the same launch: cov F*K*T*8*D^2 + E-step
    F*K*T*(8*D^2 + 8*D) per iteration (full D x D complex contractions).  The kernel executes
    fewer: Hermitian half + outer products shared by the K classes = 4*(4 + 2K) FMA per pair-frame."""
    D, T, F, K, it = CFG['D'], CFG['T'], CFG['F'], CFG['K'], CFG['em_iterations']
    alg = B * it * (F * K * T * 8 * D * D + F * K * T * (8 * D * D + 8 * D))
    pairs = D * (D + 1) // 2
    executed = B * it * F * T * 2 * (2 * pairs * (4 + 2 * K))      # E + M, 2 flops per FMA
    return alg, executed


def measure_fp64_peak(torch, _lib):
    """FP64 throughput of THIS box (the denominator of `roofline.fp64`): the probe kernels of the
    developer library (csrc/probe.cu, include/gss_dev.h), timed with CUDA events on the current stream.
    Returns TFLOP/s for the CUDA-core DFMA stream and for DMMA (mma.sync.m8n8k4.f64)."""
    import ctypes
    dl = _lib.dev_lib()
    scratch = torch.empty(int(dl.gss_debug_fp64_peak_scratch_bytes()), dtype=torch.uint8, device='cuda')
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for name, mode in (('dfma', 0), ('dmma', 1), ('dfma_3operand', 2), ('dfma_3operand_16warps', 3),
                       ('dfma_3operand_8warps', 4), ('dfma_3operand_4warps', 5)):
        fl = ctypes.c_double(0.0)
        _lib.check(dl.gss_debug_fp64_peak(mode, 2000, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(fl), stream), dl)
        best = 0.0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(dl.gss_debug_fp64_peak(mode, 20000, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(fl), stream), dl)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out[name] = best
    return out


def measure_wpe_reverberant(torch, core, ops, synth, c):
    """Side line on speech-like REVERBERANT input (synth.make_reverberant_audio: AR(2)-coloured sources
    through 2048-tap decaying room responses, 2 % noise) at the cfg2 shape: the INT8 tensor-core
    correlation build with its float64 re-do against the plain float64 build, and what the adaptive
    policy of the WPE block (core.WPE) makes of it over a few batches.  The headline generator is
    white in time, flags nothing and so says nothing about this regime."""
    dev = torch.device('cuda', torch.cuda.current_device())
    obs, _ = synth.make_reverberant_audio(7, D=c['D'], N=240000, K=3, fast=True)
    Y = ops.stft(torch.from_numpy(obs).to(dev)[None]).contiguous()                    # (1,F,D,T)
    taps, delay, its = c['taps'], c['delay'], c['wpe_iterations']

    def timed(fn, n=2):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, r

    st = torch.zeros(4, dtype=torch.int32, device=dev)
    ms_f64, x64 = timed(lambda: ops.wpe(Y, taps, delay, its, gram_mode='f64'))
    ms_i8, x8 = timed(lambda: ops.wpe(Y, taps, delay, its, gram_mode='i8+redo', stats=st))
    bins, on_list, builds = [int(v) for v in st.tolist()[:3]]
    blk = core.WPE(taps=taps, delay=delay, iterations=its, psd_context=0)
    per_call = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        blk._run(Y, info=ops.new_info(1, dev))
        e1.record()
        torch.cuda.synchronize()
        per_call.append(e0.elapsed_time(e1))
    return {'input': 'synth.make_reverberant_audio(7, D=24, N=240000): one 15 s utterance, F=513, T=%d' % Y.shape[3],
            'ms_per_utterance_f64': ms_f64, 'ms_per_utterance_i8_redo': ms_i8,
            'redo_fraction_bins': on_list / max(bins, 1), 'redo_builds_per_bin': builds / max(bins, 1),
            'max_abs_diff_i8_redo_vs_f64': float((x8 - x64).abs().max()),
            'adaptive_block_ms_per_call': per_call,
            'adaptive_block_float64_fraction_seen': blk.last_float64_fraction,
            'note': 'adaptive block: call 1 probes INT8 + re-do, later calls go straight to float64 once the '
                    'statistics of call 1 have landed (re-probed every %d calls)' % core.WPE.REPROBE}


def measure_wpe_gram(torch, _lib, ops, obs, c):
    """One WPE correlation build (all bins of the batch) through gss_debug_wpe_gram: the INT8
    tensor-core path (digit planes + tcgen05 kind::i8 GEMM) and the float64 DMMA kernel."""
    Y = ops.pack_dtf_to_fdt(obs)
    B, F, D, T = Y.shape
    taps, delay, LD = c['taps'], c['delay'], c['taps'] * c['D']
    inv = 1.0 / (Y.real.double() ** 2 + Y.imag.double() ** 2).mean(dim=2).clamp_min(1e-30)
    out = torch.empty((B, F, LD + D, LD), dtype=torch.complex128, device=Y.device)
    ws = ops.workspace(_lib.workspace_bytes(_lib.OP_WPE, B, F, D, T, 0, taps), Y.device)
    ms = {}
    for mode in (0, 1):
        def call():
            _lib.check(_lib.dev_lib().gss_debug_wpe_gram(ops._ptr(Y), ops._ptr(inv), ops._ptr(out), mode, 0, B, F, D, T,
                                                     taps, delay, None, ops._ptr(ws), ws.numel(), ops._stream()))
        for _ in range(2):
            call()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms[mode] = e0.elapsed_time(e1) / 3
    # algorithmic work: 8 (LD)^2 T + 8 LD D T real flops per bin for R and P (full matrices);
    # executed INT8 work: 15 digit-pair GEMMs over the real-stacked lower trapezoid tiles
    alg_flops = B * F * (8.0 * LD * LD * T + 8.0 * LD * D * T)
    tiles = -(-2 * LD // 128)
    cols = sum(min(2 * 24 + 128 * (i + 1), 2 * 24 + 2 * LD) for i in range(tiles)) if D == 24 else None   # D = 24: no row padding
    int8_ops = B * F * 15 * 2.0 * 128 * cols * (-(-T // 32) * 32) if cols else None
    peaks_file = ROOT / 'MEASURED_PEAKS.json'
    bf16 = json.loads(peaks_file.read_text()).get('bf16_tflops') if peaks_file.exists() else 1590.0
    res = {'kernel': 'wpe_i8_scale + wpe_i8_slice + wpe_gram_i8_kernel (tcgen05.mma kind::i8, exact integer Gram)',
           'bound': 'tensor', 'ms_per_build': ms[1], 'ms_per_build_float64_dmma': ms[0],
           'algorithmic_tflops': alg_flops / (ms[1] * 1e-3) / 1e12,
           'algorithmic_tflops_float64_dmma': alg_flops / (ms[0] * 1e-3) / 1e12,
           'unit': 'TOP/s', 'peak': 2 * bf16,
           'peak_source': '2 x measured dense bf16 (MEASURED_PEAKS.json; INT8 nominal = 2 x bf16 on B200)',
           'note': 'whole build (row scales + digit planes + GEMM); GEMM tiles are 128 x 80 (TMEM holds five accumulators + the A planes), epilogue not overlapped; see DESIGN.md 4.3 and profiles/r1_mma_rate_probe.txt'}
    if int8_ops:
        res['achieved'] = int8_ops / (ms[1] * 1e-3) / 1e12
        res['frac'] = res['achieved'] / res['peak']
    return res


def parity_check(torch, ops, enh, obs_dtf, act, c, bins=(0, 256)):
    """Outside every timed region: the tensors the bench times, checked against the CPU oracle.
    One utterance of the timed batch goes through the same device calls as `step_device`
    (all F bins); `bins` of it are compared with the oracle (a) block by block on identical
    inputs and (b) end to end (float64 oracle from the raw STFT).  `sens32` = how far the float64
    oracle itself moves when its dereverberated spectrum is rounded to complex64 (the storage
    format of the WPE -> EM hand-off): (b) cannot be tighter than that."""
    from oracle import gss_oracle as oracle
    dev = obs_dtf.device
    bins = [b for b in bins if b < c['F']]
    ctx = c['ctx_frames']
    iv = lambda v: torch.tensor([v], dtype=torch.int32, device=dev)
    Y = ops.pack_dtf_to_fdt(obs_dtf[None])
    Yw = enh.wpe_block._run(Y) if enh.wpe_block is not None else Y
    post = enh.gss_block._run(Yw, act[None])
    bf, arg = enh.bf_block._bf_args()
    X, aux = ops.beamform_from_posterior(Yw, post, iv(0), iv(ctx), iv(ctx), bf=bf, bf_arg=arg, return_aux=True)
    ref_ch = int(aux['ref_channel'][0])
    X_tf = ops.unpack_ft_to_tf(X)[0][:, bins].cpu().numpy()
    m_dev = ops.unpack_fkt_to_ktf(post)[0][:, :, bins].cpu().numpy().astype(np.float64)
    W64 = ops.unpack_fdt_to_dtf(Yw[:, bins].contiguous())[0].cpu().numpy().astype(np.complex128)
    O64 = obs_dtf[:, :, bins].cpu().numpy().astype(np.complex128)
    a = act.cpu().numpy().astype(bool)

    def drop(m):
        m = m.copy()
        m[:, :ctx] = 0
        m[:, -ctx:] = 0
        return m

    def rel(x, y):
        return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))

    wpe_kw = (dict(taps=c['taps'], delay=c['delay'], iterations=c['wpe_iterations'], psd_context=0)
              if c['taps'] else None)
    ref = oracle.enhance_stft(O64, a, 0, wpe=wpe_kw, gss_iterations=c['em_iterations'], bf=c['bf'],
                              start_context_frames=ctx, end_context_frames=ctx, ref_channel=ref_ch)
    m_stage = oracle.gss_posteriors(W64, a, c['em_iterations'])
    md = drop(m_dev)
    x_stage = oracle.beamform(W64, md[0], md[1:].sum(0), bf=c['bf'], ref_channel=ref_ch)
    out = {'utterance': 'first utterance of the first timed batch', 'bins': bins, 'ref_channel': ref_ch,
           'mask_max_abs': float(np.abs(md - ref['masks']).max()), 'xhat_rel': rel(X_tf, ref['X_hat']),
           'stagewise': {'wpe_rel': rel(W64, ref['Obs']) if wpe_kw else 0.0,
                         'mask_max_abs': float(np.abs(m_dev - m_stage).max()), 'xhat_rel': rel(X_tf, x_stage)},
           'tolerance': 1e-4}
    if wpe_kw:
        m32 = drop(oracle.gss_posteriors(ref['Obs'].astype(np.complex64).astype(np.complex128), a, c['em_iterations']))
        out['oracle_self_sensitivity_c64_handoff'] = float(np.abs(m32 - ref['masks']).max())
        # the same bins with the float64 hand-off (gss_enhance_c64_ex, GSS_ENHANCE_F64_HANDOFF): end to end
        # against the float64 oracle with no allowance (reference channel: each side's own arg-max on these bins)
        ref2 = oracle.enhance_stft(O64, a, 0, wpe=wpe_kw, gss_iterations=c['em_iterations'], bf=c['bf'],
                                   start_context_frames=ctx, end_context_frames=ctx)
        Xf, pf = ops.enhance(obs_dtf[:, :, bins].contiguous()[None], act[None], iv(0), iv(ctx), iv(ctx),
                             wpe=(c['taps'], c['delay'], c['wpe_iterations'], 0), em_iterations=c['em_iterations'],
                             bf=c['bf'], handoff='f64')
        out['float64_handoff'] = {'mask_max_abs': float(np.abs(drop(pf[0].cpu().numpy().astype(np.float64)) - ref2['masks']).max()),
                                  'xhat_rel': rel(Xf[0].cpu().numpy(), ref2['X_hat'])}
    st = out['stagewise']
    out['ok'] = bool(st['wpe_rel'] < 1e-4 and st['mask_max_abs'] < 1e-4 and st['xhat_rel'] < 1e-4 and out['xhat_rel'] < 1e-4
                     and out['mask_max_abs'] < max(1e-4, 2 * out.get('oracle_self_sensitivity_c64_handoff', 0.0)))
    return out


def run_session_job(torch, enh, items, c, rank, world, schedule, device_resident=False, bases=None, seed=77):
    """One pass over a dev-shaped work list through the session driver (session.py): every rank holds
    the base recordings in host memory (pageable, like freshly read wav data), a loader thread cuts the
    segments, pins them and frames the activity, the main thread runs one ragged batch per hot-path
    pass, finished utterances are cut back to their own samples and handed to a sink (host float64
    arrays; wav encoding is outside the hot path).  Returns this rank's SessionReport + byte counts."""
    from pb_chime5_b200 import sharding, synth
    from pb_chime5_b200.session import SessionScheduler, run_distributed
    speakers = [f'P{k:02d}' for k in range(c['K'] - 1)]
    if bases is None:
        bases = [synth.make_base_recording(seed + i, D=c['D'], K=c['K'], seconds=c['context_s'] * 2 + 21.5)
                 for i in range(4)]
    if device_resident:
        bases = [(torch.from_numpy(o).cuda() if not isinstance(o, torch.Tensor) else o, a) for o, a in bases]
    exs = [synth.work_item_example(it, speakers) for it in items]
    counts = {'h2d': 0, 'd2h': 0, 'utt_samples': 0}

    def load(ex):
        it = ex['item']
        obs, act = bases[it['base']]
        a, b = it['offset'], it['offset'] + it['total']
        acts = {spk: act[k, a:b] for k, spk in enumerate(speakers)}
        acts['Noise'] = act[-1, a:b]
        return obs[:, a:b], acts, ex['speaker_id']

    def finish(ex, x):
        it = ex['item']
        return x[..., it['context']:it['context'] + it['num_samples_orig']]

    def sink(ex, x):                                   # writer thread (single): byte counts per finished utterance
        counts['h2d'] += c['D'] * ex['item']['total'] * 4      # float32 samples of the segment, all channels
        counts['d2h'] += ex['item']['total'] * 4               # enhanced float32 samples of the segment (before the context cut)
        counts['utt_samples'] += x.shape[-1]

    sched = SessionScheduler(enh, load, None, finish, batch_size=c['batch_size'], window=64,
                             max_batch_samples=c['batch_size'] * 52 * 16000, prefetch=1, skip_existing=False,
                             strict=True, sink_fn=sink)
    rep = run_distributed(sched, exs, schedule)
    return rep, counts, bases


def session_summary(reports, counts, n_items, items, sec, world):
    """rank-0 view of a session pass: reports / counts = one entry per rank"""
    busy = [r.busy_seconds for r in reports]
    seg_s = sum(it['total'] for it in items) / 16000.0
    utt_s = sum(it['num_samples_orig'] for it in items) / 16000.0
    padded = sum(r.padded_samples for r in reports)
    valid = sum(r.valid_samples for r in reports)
    return {'utterances': n_items, 'seconds': sec, 'utterances_per_s': n_items / sec,
            'segment_audio_s_per_s': seg_s / sec, 'utterance_audio_s_per_s': utt_s / sec,
            'real_time_factor_segments': sec / seg_s,
            'per_rank': {'done': [r.done for r in reports], 'batches': [r.batches for r in reports],
                         'busy_s': busy, 'wall_s': [r.seconds for r in reports],
                         'starved_s': [r.wait_seconds for r in reports], 'loader_s': [r.load_seconds for r in reports]},
            'busy_imbalance_max_over_min': (max(busy) / min(busy)) if min(busy) > 0 else None,
            'padding_efficiency': valid / padded if padded else None,
            'h2d_bytes': sum(cn['h2d'] for cn in counts), 'd2h_bytes': sum(cn['d2h'] for cn in counts),
            'failed': sum(len(r.failed) for r in reports)}


def run_session_bench(args):
    """--workload cfg3 | cfg4: the metric through the session driver on dev-shaped ragged work lists."""
    import torch
    from pb_chime5_b200 import _lib, core, sharding, synth
    rank, world = sharding.init_process_group()
    dev = torch.device('cuda', torch.cuda.current_device())
    c = CFG
    n_utts = args.utterances or (c['n_utts'] if args.workload == 'cfg3' else 96 * world)
    enh = core.get_enhancer(wpe=True, wpe_tabs=c['taps'], wpe_delay=c['delay'], wpe_iterations=c['wpe_iterations'],
                            bss_iterations=c['em_iterations'], bf=c['bf'], context_samples=int(c['context_s'] * 16000))
    # rank 0 draws the work list; everybody gets the same copy (the one collective of the job)
    items = sharding.broadcast_work_list(synth.make_work_list(4242, n_utts, context_s=c['context_s'], K=c['K'],
                                                              min_s=c['min_s']) if rank == 0 else None)
    warm = items[:min(len(items), 2 * c['batch_size'] * world)]
    bases = None
    results = {}
    sampler = None
    for name, resident, steps, warmup in (('e2e', False, args.steps, args.warmup), ('device', True, args.steps, 1)):
        for _ in range(warmup):
            _, _, bases = run_session_job(torch, enh, warm, c, rank, world, args.schedule, resident, bases)
        secs, summaries = [], []
        l0 = _lib.lib().gss_launch_count()
        if name == 'e2e' and rank == 0:
            sampler = ClockSampler(torch.cuda.current_device())
            sampler.start()
        for _ in range(steps):
            sharding.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rep, cnt, bases = run_session_job(torch, enh, items, c, rank, world, args.schedule, resident, bases)
            torch.cuda.synchronize()
            sharding.barrier()
            sec = sharding.max_over_ranks(time.perf_counter() - t0, dev)
            reps, cnts = sharding.gather_objects(rep), sharding.gather_objects(cnt)
            if rank == 0:
                secs.append(sec)
                summaries.append(session_summary(reps, cnts, len(items), items, sec, world))
        if name == 'e2e' and sampler is not None:
            results['clocks'] = sampler.stop()
        results[name] = (secs, summaries, _lib.lib().gss_launch_count() - l0)
        if resident:
            bases = None
    if rank == 0:
        secs, summ, launches = results['e2e']
        dsecs, dsumm, _ = results['device']
        best = summ[int(np.argmin(secs))]
        value = len(items) * len(dsecs) / sum(dsecs)
        e2e_value = len(items) * len(secs) / sum(secs)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'utterances/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * sum(dsecs) / len(dsecs), 'higher_is_better': True,
            'scaling': 'strong' if args.utterances or args.workload == 'cfg3' else 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'utterances_per_step': len(items), 'schedule': args.schedule,
                       'batch_size': c['batch_size'],
                       'value_region': 'base recordings resident in HBM, results stay on the device; timed: '
                                       'segment cut, STFT, WPE, EM, beamformer, iSTFT, context cut (host clock '
                                       'around whole passes, max over ranks)',
                       'l2': 'every batch is > 1 GB of STFT data, far beyond the 126 MB L2'},
            'e2e': {'value': e2e_value, 'unit': 'utterances/s', 'h2d_bytes_per_step': best['h2d_bytes'],
                    'd2h_bytes_per_step': best['d2h_bytes'], 'ms_per_step': 1e3 * sum(secs) / len(secs),
                    'api': 'SessionScheduler + Enhancer.enhance_prepared_batch: raw float32 audio in pageable host '
                           'memory in, float64 utterance samples in host memory out'},
            'gpu_launches': int(launches),
            'session': best, 'session_device_resident': dsumm[int(np.argmin(dsecs))],
            'wpe_float64_fraction': enh.wpe_block.last_float64_fraction,
            'clocks': results.get('clocks'),
        }
        print(json.dumps(line), flush=True)
    sharding.barrier()
    sharding.shutdown()


def run_gpu(args):
    import torch
    from pb_chime5_b200 import _lib, core, ops, sharding, synth

    rank, world = sharding.init_process_group()
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    c = CFG
    B = args.batch
    enh = core.get_enhancer(wpe=bool(c['taps']), wpe_tabs=max(c['taps'], 1), wpe_delay=c['delay'],
                            wpe_iterations=c['wpe_iterations'],
                            bss_iterations=c['em_iterations'], bf=c['bf'])

    # two distinct host batches (pinned), alternated between steps; each is B*93 MB >> L2
    nsets = 2
    host_obs, host_act = [], []
    for s in range(nsets):
        obs, act = synth.make_batch(1000 * (rank + 1) + 100 * s, B, D=c['D'], T=c['T'], F=c['F'], K=c['K'])
        host_obs.append(torch.from_numpy(obs).pin_memory())
        host_act.append(torch.from_numpy(act.astype(np.uint8)))
    dev_obs = [h.to(dev) for h in host_obs]
    dev_act = [h.to(dev) for h in host_act]
    ti = torch.zeros(B, dtype=torch.int32, device=dev)
    ctx = torch.full((B,), c['ctx_frames'], dtype=torch.int32, device=dev)
    out_host = {'X_hat': torch.empty((B, c['T'], c['F']), dtype=torch.complex64, pin_memory=True),
                'masks': torch.empty((B, c['K'], c['T'], c['F']), dtype=torch.float32, pin_memory=True)}

    em_events = []
    info_dev = ops.new_info(B, dev, stages=3)        # status words of the three stages, read after the timed region

    def step_device(i, time_em=False):
        """Whole hot path, inputs resident in HBM in the reference layout (B,D,T,F); the body of
        Enhancer.enhance_stft_batch with CUDA events around the EM kernel.  No host synchronisation."""
        s = i % nsets
        Y = ops.pack_dtf_to_fdt(dev_obs[s])
        if enh.wpe_block is not None:
            Y = enh.wpe_block._run(Y, info=info_dev[0])
        if time_em:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        post = enh.gss_block._run(Y, dev_act[s], info=info_dev[1])
        if time_em:
            e1.record()
            em_events.append((e0, e1))
        X = enh.bf_block._run_from_posterior(Y, post, ti, ctx, ctx, info=info_dev[2])
        return ops.unpack_ft_to_tf(X), ops.unpack_fkt_to_ktf(post)

    def step_e2e(i):
        s = i % nsets
        return enh.enhance_stft_host(host_obs[s], host_act[s], ti, ctx, ctx, out=out_host)

    def timed(fn, steps, **kw):
        sharding.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i, **kw)
        e1.record()
        torch.cuda.synchronize()
        sharding.barrier()
        return sharding.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)

    for i in range(args.warmup):
        step_device(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _lib.lib().gss_launch_count()
    sec = timed(step_device, args.steps, time_em=True)
    launches = _lib.lib().gss_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    em_ms = statistics.mean(a.elapsed_time(b) for a, b in em_events)
    ops.check_info(info_dev, enh.STAGES)

    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    sec_e2e = timed(step_e2e, args.steps)

    # the same through the pipelined public API (copies of neighbouring steps overlap the kernels)
    def run_stream(nsteps):
        gen = enh.enhance_stft_host_stream(
            ((host_obs[i % nsets], host_act[i % nsets], ti, ctx, ctx) for i in range(nsteps)), reuse_outputs=True)
        for res in gen:
            last = res
        return last
    run_stream(2)
    sharding.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_stream(args.steps)
    e1.record()
    torch.cuda.synchronize()
    sharding.barrier()
    sec_e2e_pipe = sharding.max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)

    # side measurement: the WPE correlation build alone (INT8 tensor cores vs the float64 DMMA build)
    wpe_gram = None
    if c['taps'] and rank == 0:
        try:
            wpe_gram = measure_wpe_gram(torch, _lib, ops, dev_obs[0], c)
        except Exception as ex:  # noqa: BLE001  (diagnostic only, never fails the bench)
            wpe_gram = {'error': repr(ex)}

    parity = None
    if rank == 0 and not args.no_parity:
        try:
            parity = parity_check(torch, ops, enh, dev_obs[0][0], dev_act[0][0], c)
        except Exception as ex:  # noqa: BLE001  (reported, never hidden)
            parity = {'ok': False, 'error': repr(ex)}

    # N > 1: a real sharded job after the headline timing -- a cfg4-shaped work list (dev-shaped ragged
    # lengths + 15 s context per side, raw audio from host memory) drawn by rank 0, broadcast over NCCL and
    # farmed out over the ranks by the work queue; exposes load imbalance, padding and host-side loading
    sharded = None
    if world > 1 and not args.no_sharded:
        try:
            c4, _ = SESSION_WORKLOADS['cfg4']
            enh4 = core.get_enhancer(wpe=True, wpe_tabs=c4['taps'], wpe_delay=c4['delay'], wpe_iterations=c4['wpe_iterations'],
                                     bss_iterations=c4['em_iterations'], bf=c4['bf'], context_samples=int(c4['context_s'] * 16000))
            items = sharding.broadcast_work_list(
                synth.make_work_list(4242, (args.utterances or 64) * world, context_s=c4['context_s'], K=c4['K'])
                if rank == 0 else None)
            _, _, bases = run_session_job(torch, enh4, items[:2 * c4['batch_size'] * world], c4, rank, world, args.schedule)
            sharding.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rep, cnt, _ = run_session_job(torch, enh4, items, c4, rank, world, args.schedule, bases=bases)
            torch.cuda.synchronize()
            sharding.barrier()
            sec4 = sharding.max_over_ranks(time.perf_counter() - t0, dev)
            reps, cnts = sharding.gather_objects(rep), sharding.gather_objects(cnt)
            if rank == 0:
                sharded = session_summary(reps, cnts, len(items), items, sec4, world)
                sharded['workload'] = SESSION_WORKLOADS['cfg4'][1]
                sharded['schedule'] = 'dynamic (task farm)' if args.schedule == 'auto' else args.schedule
        except Exception as ex:  # noqa: BLE001  (reported, never hidden)
            sharded = {'error': repr(ex)}

    wpe_reverb = None
    if c['taps'] and rank == 0 and args.workload == 'cfg2':
        try:
            wpe_reverb = measure_wpe_reverberant(torch, core, ops, synth, c)
        except Exception as ex:  # noqa: BLE001
            wpe_reverb = {'error': repr(ex)}

    total_utts = world * B * args.steps
    value = total_utts / sec
    e2e_value = total_utts / sec_e2e
    h2d = host_obs[0].numel() * 8 + host_act[0].numel() + 3 * 4 * B
    d2h = out_host['X_hat'].numel() * 8 + out_host['masks'].numel() * 4

    peaks_file = ROOT / 'MEASURED_PEAKS.json'
    if peaks_file.exists():
        peak, peak_src = json.loads(peaks_file.read_text())['hbm_gbs'], 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    alg_bytes, fused_min = algorithmic_bytes_em(B)
    achieved = alg_bytes / (em_ms * 1e-3) / 1e9
    alg_flops, exe_flops = algorithmic_flops_em(B)
    traffic = traffic_src = None
    tfile = ROOT / 'profiles' / 'em_kernel_traffic.json'
    if tfile.exists():
        try:
            tj = json.loads(tfile.read_text())
            traffic = tj.get('dram_bytes_per_utterance') * B
            traffic_src = (f"profiles/em_kernel_traffic.json: ncu --set full of {tj.get('kernel', 'the EM kernel')[:60]} at git "
                           f"{tj.get('git')} ({tj.get('utterances_in_launch')} utterance(s), {tj.get('em_iterations')} iterations; "
                           f"bytes per utterance x {B})")
        except Exception:
            traffic = None
    fp64_peak = None
    if rank == 0:
        try:
            fp64_peak = measure_fp64_peak(torch, _lib)
        except Exception as ex:  # noqa: BLE001
            fp64_peak = {'error': repr(ex)}

    # headline end-to-end number: host buffers in, host buffers out, every step copied in full;
    # the faster of the two public host APIs (one synchronous call per batch / the pipelined stream)
    e2e_sync = {'value': e2e_value, 'ms_per_step': 1e3 * sec_e2e / args.steps,
                'api': 'Enhancer.enhance_stft_host (pinned host STFT in, X_hat + masks out, one synchronous call per batch)'}
    e2e_pipe = {'value': total_utts / sec_e2e_pipe, 'ms_per_step': 1e3 * sec_e2e_pipe / args.steps,
                'api': 'Enhancer.enhance_stft_host_stream (same copies per batch; those of neighbouring batches overlap the kernels)'}
    best = e2e_pipe if e2e_pipe['value'] > e2e_sync['value'] else e2e_sync
    e2e = {'value': best['value'], 'unit': 'utterances/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
           'ms_per_step': best['ms_per_step'], 'api': best['api'], 'synchronous': e2e_sync, 'pipelined': e2e_pipe}

    if rank != 0:
        sharding.barrier()
        sharding.shutdown()
        return
    line = {
        'metric': METRIC, 'value': value, 'unit': 'utterances/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sec / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': B,
                   'l2': f"inputs {B * c['D'] * c['T'] * c['F'] * 8 // 1000000} MB per step per GPU vs 126 MB L2, two input sets alternated",
                   'value_region': 'inputs resident in HBM in the reference layout (B,D,T,F); timed: pack, WPE, EM, beamformer, unpack',
                   'arithmetic': 'complex64 storage, float64 arithmetic'},
        'e2e': e2e,
        'parity': parity,
        'sharded': sharded,
        'gpu_launches': int(launches),
        'roofline': {'kernel': 'cacgmm_em_kernel (fused EM, all iterations)', 'bound': 'hbm',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': alg_bytes,
                     'fused_minimum_bytes_per_launch': fused_min,
                     'kernel_ms': em_ms,
                     'fp64': {'algorithmic_tflops': alg_flops / (em_ms * 1e-3) / 1e12,
                              'executed_tflops': exe_flops / (em_ms * 1e-3) / 1e12,
                              'peak_tflops_measured': (fp64_peak or {}).get('dfma'),
                              'peak_tflops_measured_dmma': (fp64_peak or {}).get('dmma'),
                              'peak_tflops_measured_dfma_3_register_operands': (fp64_peak or {}).get('dfma_3operand'),
                              'peak_tflops_measured_dfma_3_register_operands_16_warps_per_sm': (fp64_peak or {}).get('dfma_3operand_16warps'),
                              'frac_of_measured_peak': (exe_flops / (em_ms * 1e-3) / 1e12 / fp64_peak['dfma'])
                              if fp64_peak and fp64_peak.get('dfma') else None,
                              'peak_source': 'measured in this run: gss_debug_fp64_peak (csrc/probe.cu, libgss_dev.so), '
                                             'best of 3, CUDA events'},
                     'note': 'declared variant: per-iteration covariance+E-step streams x EM iterations (SURVEY 8d); '
                             'the kernel is FP64-pipe bound at D=24, see DESIGN.md'},
        'clocks': clocks,
    }
    if wpe_gram is not None:
        line['roofline_wpe_gram'] = wpe_gram
    if wpe_reverb is not None:
        line['wpe_reverberant'] = wpe_reverb
    if world == 1 and not args.no_cpu:
        import multiprocessing as mp
        procs = os.cpu_count() or 1
        with mp.get_context('fork').Pool(procs) as pool:
            cbins = args.cpu_bins or 8
            v, wall = cpu_sample(pool, procs, cbins, 30_000)
        line['cpu_baseline'] = {
            'value': v, 'unit': 'utterances/s', 'cores': procs, 'kind': 'port',
            'sample': f'{procs} processes x {cbins} of 513 bins, all iterations, scaled by 513/{cbins}; '
                      f'{wall:.1f} s wall'}
    print(json.dumps(line), flush=True)
    sharding.barrier()
    sharding.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=0, help='utterances per GPU per step (0: 8 for cfg2, 32 for cfg1, 1 for cfg5)')
    ap.add_argument('--cpu-bins', type=int, default=0, help='frequency bins per process in a CPU sample (0 = 8 for the cpu_baseline of the GPU arm, 4 per step for --impl reference)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-parity', action='store_true', help='skip the oracle spot check of the timed tensors')
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS) + sorted(SESSION_WORKLOADS),
                    help='cfg2 = the metric configuration (default); cfg1 / cfg5 are side measurements; cfg3 / cfg4 = '
                         'dev-shaped ragged work lists through the session driver (raw audio in and out)')
    ap.add_argument('--utterances', type=int, default=0,
                    help='cfg3 / cfg4: utterances per step (0: 512 for cfg3, 96 per GPU for cfg4)')
    ap.add_argument('--schedule', default='auto', choices=['auto', 'dynamic', 'lpt', 'strided'],
                    help='cfg3 / cfg4 and the sharded section: work distribution over the ranks')
    ap.add_argument('--no-sharded', action='store_true', help='skip the sharded work-list section at N > 1')
    args = ap.parse_args()
    global CFG, WORKLOAD
    if args.workload in SESSION_WORKLOADS:
        CFG, WORKLOAD = SESSION_WORKLOADS[args.workload]
        if args.impl == 'reference':
            raise SystemExit('--impl reference times cfg1 / cfg2 / cfg5')
        run_session_bench(args)
        return
    CFG, WORKLOAD = WORKLOADS[args.workload]
    if args.batch <= 0:
        args.batch = {'cfg2': 8, 'cfg1': 32, 'cfg5': 1}[args.workload]
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
