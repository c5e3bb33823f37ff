"""Session driver: length-bucketed batches, audio prefetch, asynchronous wav writing, resume
and per-example failure isolation.

The reference enhances one example at a time inside an MPI task farm
(``Enhancer.enhance_session``, ``pb_chime5/core.py:333-394``:
``for ex in dlp_mpi.split_managed(it): x_hat = self.enhance_example(ex); dump_audio(...)``).
On a GPU one utterance does not fill the machine (513 CTAs per kernel on 148 SMs), and reading
/ writing wav files would leave it idle, so here

* the examples of this rank's shard are grouped into batches of similar length (sorted inside
  a sliding window, so the order on disk stays roughly sequential) and enhanced with ONE pass of
  the hot path per batch (``Enhancer.enhance_observation_batch``, ragged frame counts);
* a loader thread reads the audio of the next batches while the GPU works, a writer thread
  dumps the finished wavs (``audio_io.dump_audio``, the reference's peak normalisation);
* examples whose output file exists are skipped (resume); if a batch fails, its examples are
  retried one by one and the failures are reported at the end instead of killing the run
  (the reference's RTTM front door prints ``ERROR: Failed example`` and re-raises,
  ``core_chime6_rttm.py:169-185``; ``strict=True`` keeps that behaviour).

Nothing here computes: numerics are the libgss kernels behind ``Enhancer``.
"""
from __future__ import annotations

import queue
import threading
import time
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import audio_io


def plan_batches(lengths, keys=None, batch_size=8, window=64, max_batch_samples=None):
    """Group example indices into batches.

    lengths[i]: samples of example i; keys[i]: examples may only share a batch when their key
    (channel count, class count, ...) is equal.  Examples are taken `window` at a time in their
    original order, sorted by decreasing length inside the window (stable), and cut into batches
    of at most `batch_size` examples whose padded size ``len(batch) * max(length)`` does not
    exceed `max_batch_samples`.  Deterministic; every index appears exactly once."""
    n = len(lengths)
    keys = [None] * n if keys is None else list(keys)
    assert len(keys) == n and batch_size >= 1 and window >= 1
    batches = []
    for w0 in range(0, n, window):
        idx = sorted(range(w0, min(w0 + window, n)), key=lambda i: (-lengths[i], i))
        groups = {}
        for i in idx:
            groups.setdefault(keys[i], []).append(i)
        for key in sorted(groups, key=lambda k: min(groups[k])):
            cur = []
            for i in groups[key]:
                longest = lengths[cur[0]] if cur else lengths[i]
                if cur and (len(cur) >= batch_size or
                            (max_batch_samples is not None and (len(cur) + 1) * longest > max_batch_samples)):
                    batches.append(cur)
                    cur = []
                cur.append(i)
            if cur:
                batches.append(cur)
    return batches


@dataclass
class SessionReport:
    done: int = 0
    skipped: int = 0
    failed: list = field(default_factory=list)      # (example_id, repr(exception))
    batches: int = 0
    seconds: float = 0.0
    audio_seconds: float = 0.0

    def __str__(self):
        rtf = self.seconds / self.audio_seconds if self.audio_seconds else float('nan')
        return (f'{self.done} enhanced in {self.batches} batches, {self.skipped} skipped (exist), '
                f'{len(self.failed)} failed, {self.seconds:.1f} s wall, real-time factor {rtf:.4f}')


class SessionScheduler:
    """Drives ``enhancer.enhance_observation_batch`` over a list of examples.

    load_fn(ex)   -> (obs (D, N) float array, ex_array_activity {speaker: (N,) bool}, speaker_id)
    finish_fn(ex, x_hat) -> the samples to write (context cut; default: identity)
    path_fn(ex)   -> output wav path
    """

    def __init__(self, enhancer, load_fn, path_fn, finish_fn=None, *, batch_size=8, window=64,
                 max_batch_samples=8 * 60 * 16000, prefetch=2, skip_existing=True, strict=False,
                 sample_rate=16000, verbose=False):
        self.enhancer, self.load_fn, self.path_fn = enhancer, load_fn, path_fn
        self.finish_fn = finish_fn or (lambda ex, x: x)
        self.batch_size, self.window, self.max_batch_samples = batch_size, window, max_batch_samples
        self.prefetch, self.skip_existing, self.strict = prefetch, skip_existing, strict
        self.sample_rate, self.verbose = sample_rate, verbose

    # -- planning ---------------------------------------------------------------------------
    @staticmethod
    def example_length(ex):
        """samples of the segment (metadata only, no file access): CHiME-6 style flat indices,
        CHiME-5 style per-array indices, else 'num_samples', else 0 (unknown: keeps the order)"""
        try:
            start, end = ex['start'], ex['end']
            if isinstance(start, dict):
                arr = ex.get('reference_array') or sorted(start['observation'])[0]
                start, end = start['observation'][arr], end['observation'][arr]
            return int(end) - int(start)
        except (KeyError, TypeError, AttributeError, ValueError):
            n = ex.get('num_samples', 0) if isinstance(ex, dict) else 0
            return int(n) if isinstance(n, (int, float)) else 0

    def plan(self, examples):
        todo, skipped = [], 0
        for i, ex in enumerate(examples):
            if self.skip_existing and Path(self.path_fn(ex)).exists():
                skipped += 1
            else:
                todo.append(i)
        lengths = [self.example_length(examples[i]) for i in todo]
        keys = [(examples[i].get('session_id'),) for i in todo]      # same session: same arrays and speakers
        batches = plan_batches(lengths, keys, self.batch_size, self.window, self.max_batch_samples)
        return [[todo[j] for j in b] for b in batches], skipped

    # -- execution --------------------------------------------------------------------------
    def run(self, examples):
        report = SessionReport()
        t0 = time.perf_counter()
        batches, report.skipped = self.plan(examples)
        loaded = queue.Queue(maxsize=max(1, self.prefetch))
        to_write = queue.Queue(maxsize=4 * self.batch_size)
        errors = []

        def loader():
            for b in batches:
                items = []
                for i in b:
                    ex = examples[i]
                    try:
                        items.append((ex, self.load_fn(ex), None))
                    except Exception as e:  # noqa: BLE001  (isolated per example)
                        items.append((ex, None, e))
                loaded.put(items)
            loaded.put(None)

        def writer():
            while True:
                job = to_write.get()
                if job is None:
                    return
                ex, x = job
                try:
                    path = Path(self.path_fn(ex))
                    path.parent.mkdir(parents=True, exist_ok=True)
                    audio_io.dump_audio(x, path, sample_rate=self.sample_rate)
                except Exception as e:  # noqa: BLE001
                    errors.append((ex.get('example_id'), repr(e)))

        lt = threading.Thread(target=loader, daemon=True)
        wt = threading.Thread(target=writer, daemon=True)
        lt.start()
        wt.start()
        try:
            while True:
                items = loaded.get()
                if items is None:
                    break
                report.batches += 1
                good = []
                for ex, data, err in items:
                    if err is not None:
                        self._fail(report, ex, err)
                    else:
                        good.append((ex, data))
                if not good:
                    continue
                try:
                    outs = self._enhance([d for _, d in good], [e for e, _ in good])
                except Exception as e:  # noqa: BLE001
                    if self.strict or len(good) == 1:
                        outs = None
                        first_err = e
                    else:
                        outs, first_err = [], None          # retry one by one: isolate the bad example
                        for ex, d in good:
                            try:
                                outs.append(self._enhance([d], [ex])[0])
                            except Exception as e1:  # noqa: BLE001
                                outs.append(e1)
                    if outs is None:
                        for ex, _ in good:
                            self._fail(report, ex, first_err)
                        continue
                for (ex, d), x in zip(good, outs):
                    if isinstance(x, Exception):
                        self._fail(report, ex, x)
                        continue
                    x = self.finish_fn(ex, x)
                    report.audio_seconds += x.shape[-1] / self.sample_rate
                    to_write.put((ex, x))
                    report.done += 1
        finally:
            to_write.put(None)
            wt.join()
        for eid, msg in errors:
            report.failed.append((eid, msg))
        report.seconds = time.perf_counter() - t0
        if self.strict and report.failed:
            raise RuntimeError(f'ERROR: Failed example: {report.failed[0][0]}: {report.failed[0][1]}')
        return report

    def _enhance(self, datas, exs):
        obs = [d[0] for d in datas]
        acts = [d[1] for d in datas]
        spk = [d[2] for d in datas]
        return self.enhancer.enhance_observation_batch(obs, acts, spk, exs)

    def _fail(self, report, ex, err):
        print('ERROR: Failed example:', ex.get('example_id'), repr(err), flush=True)
        report.failed.append((ex.get('example_id'), repr(err)))
        if self.strict:
            raise err


def stack_arrays(arrays, multiarray):
    """Channel selection of Enhancer.enhance_example (core.py:464-498): list of (C, N_a) arrays
    (sorted by array name) -> (D, N); arrays may differ in length by a few samples."""
    selectors = {True: slice(None), 'outer_array_mics': (0, -1), 'first_array_mics': (0,)}
    if multiarray not in selectors:
        raise ValueError(multiarray)
    assert {v.ndim for v in arrays} == {2}, [v.shape for v in arrays]
    n = min(v.shape[-1] for v in arrays)
    sel = selectors[multiarray]
    return np.concatenate([v[sel, :n] for v in arrays], axis=0)
