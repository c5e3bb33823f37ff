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

import queue as queue_mod
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
    busy_seconds: float = 0.0                       # time inside the enhancement calls (this rank)
    wait_seconds: float = 0.0                       # main thread blocked on the loader (GPU starved)
    load_seconds: float = 0.0                       # loader thread: reading / cutting / pinning / framing
    padded_samples: int = 0                         # samples the batches occupied after padding ...
    valid_samples: int = 0                          # ... and the samples that were real input

    @property
    def padding_efficiency(self):
        return self.valid_samples / self.padded_samples if self.padded_samples else float('nan')

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
                 sample_rate=16000, verbose=False, sink_fn=None, loader_threads=4):
        """sink_fn(ex, x): where a finished utterance goes instead of a wav file (in-memory runs,
        benchmarks); path_fn may then be None.  loader_threads: the examples of a batch are read /
        cut / pinned / framed by this many threads (NumPy and the copies release the GIL): one thread
        needs ~75 ms per 32 s, 24-channel segment, the GPU ~65 ms."""
        self.enhancer, self.load_fn, self.path_fn = enhancer, load_fn, path_fn
        self.sink_fn = sink_fn
        self.finish_fn = finish_fn or (lambda ex, x: x)
        self.batch_size, self.window, self.max_batch_samples = batch_size, window, max_batch_samples
        self.prefetch, self.skip_existing, self.strict = prefetch, skip_existing, strict
        self.loader_threads = max(1, int(loader_threads))
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
            if self.skip_existing and self.path_fn is not None and Path(self.path_fn(ex)).exists():
                skipped += 1
            else:
                todo.append(i)
        lengths = [self.example_length(examples[i]) for i in todo]
        keys = [(examples[i].get('session_id'),) for i in todo]      # same session: same arrays and speakers
        batches = plan_batches(lengths, keys, self.batch_size, self.window, self.max_batch_samples)
        return [[todo[j] for j in b] for b in batches], skipped

    # -- execution --------------------------------------------------------------------------
    def run(self, examples, queue=None):
        """Enhance `examples`.  queue=None: all planned batches, in plan order (the list is this
        rank's own shard).  queue = an iterable of batch indices, or a callable n_batches -> iterable
        (`sharding.WorkQueue`): `examples`
        is the WHOLE job on every rank, every rank makes the same plan (deterministic, longest
        batches first) and takes the batches the queue hands it -- the task farm of
        `dlp_mpi.split_managed` (pb_chime5/core.py:381)."""
        report = SessionReport()
        t0 = time.perf_counter()
        batches, report.skipped = self.plan(examples)
        if queue is not None:
            order = sorted(range(len(batches)),
                           key=lambda j: (-len(batches[j]) * max(self.example_length(examples[i]) for i in batches[j]), j))
            batches = [batches[j] for j in order]
            if callable(queue):
                queue = queue(len(batches))
            batch_iter = (batches[j] for j in queue)
        else:
            batch_iter = iter(batches)
        loaded = queue_mod.Queue(maxsize=max(1, self.prefetch))
        to_write = queue_mod.Queue(maxsize=4 * self.batch_size)
        errors = []
        stop = threading.Event()
        slots = threading.Semaphore(max(1, self.prefetch))
        prepare = getattr(self.enhancer, 'prepare_observation', None)
        cuda_device = None
        if prepare is not None:
            try:
                import torch
                if torch.cuda.is_available():
                    cuda_device = torch.cuda.current_device()
            except ImportError:
                pass

        def put(q, item):
            """blocking put that gives up when the run was stopped (no thread left hanging on a full queue)"""
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue_mod.Full:
                    continue
            return False

        def load_one(i):
            ex = examples[i]
            try:
                data = self.load_fn(ex)
                if prepare is not None:
                    # host half of the hot path (pinned float32 samples, upload, activity framing)
                    # here, off the main thread, while the GPU works on the previous batch
                    data = data + (prepare(data[0], data[1], data[2], ex),)
                return (ex, data, None)
            except Exception as e:  # noqa: BLE001  (isolated per example)
                return (ex, None, e)

        def bind_device():
            if cuda_device is not None:
                import torch
                torch.cuda.set_device(cuda_device)            # the current device is thread-local

        def loader():
            pool = None
            try:
                bind_device()
                if self.loader_threads > 1:
                    from concurrent.futures import ThreadPoolExecutor
                    pool = ThreadPoolExecutor(self.loader_threads, initializer=bind_device)
                while True:
                    # a batch is claimed (from the shared work queue, when there is one) only when a
                    # prefetch slot is free: at most `prefetch` claimed batches wait behind the one in
                    # flight, so the tail of a task-farmed job stays short
                    while not slots.acquire(timeout=0.1):
                        if stop.is_set():
                            return
                    if stop.is_set():
                        return
                    b = next(batch_iter, None)
                    if b is None:
                        return
                    tl = time.perf_counter()
                    items = list(pool.map(load_one, b)) if pool is not None else [load_one(i) for i in b]
                    report.load_seconds += time.perf_counter() - tl
                    if not put(loaded, items):
                        return
            finally:
                if pool is not None:
                    pool.shutdown(wait=False)
                put(loaded, None)

        def writer():
            while True:
                job = to_write.get()
                if job is None:
                    return
                ex, x = job
                try:
                    if self.sink_fn is not None:
                        self.sink_fn(ex, x)
                    else:
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
                tw = time.perf_counter()
                items = loaded.get()
                report.wait_seconds += time.perf_counter() - tw
                if items is None:
                    break
                slots.release()
                report.batches += 1
                good = []
                for ex, data, err in items:
                    if err is not None:
                        self._fail(report, ex, err)
                    else:
                        good.append((ex, data))
                # one pass of the hot path per (channels, classes) signature: a recording with a
                # missing array or an extra speaker does not take its batch mates down
                groups = {}
                for ex, d in good:
                    groups.setdefault((int(d[0].shape[0]), len(d[1])), []).append((ex, d))
                for group in groups.values():
                    self._run_group(report, group, to_write)
        finally:
            stop.set()
            while True:                                   # unblock and drop whatever the loader prefetched
                try:
                    loaded.get_nowait()
                except queue_mod.Empty:
                    break
            to_write.put(None)
            wt.join()
            lt.join(timeout=5)
        for eid, msg in errors:
            report.failed.append((eid, msg))
        report.seconds = time.perf_counter() - t0
        if self.strict and report.failed:
            raise RuntimeError(f'ERROR: Failed example: {report.failed[0][0]}: {report.failed[0][1]}')
        return report

    def _run_group(self, report, good, to_write):
        tb = time.perf_counter()
        try:
            outs = self._enhance([d for _, d in good], [e for e, _ in good])
        except Exception as e:  # noqa: BLE001
            if self.strict or len(good) == 1:
                for ex, _ in good:
                    self._fail(report, ex, e)
                return
            outs = []                                     # retry one by one: isolate the bad example
            for ex, d in good:
                try:
                    outs.append(self._enhance([d], [ex])[0])
                except Exception as e1:  # noqa: BLE001
                    outs.append(e1)
        finally:
            report.busy_seconds += time.perf_counter() - tb
        lens = [int(d[0].shape[-1]) for _, d in good]
        report.padded_samples += len(lens) * max(lens)
        report.valid_samples += sum(lens)
        for (ex, d), x in zip(good, outs):
            if isinstance(x, Exception):
                self._fail(report, ex, x)
                continue
            x = self.finish_fn(ex, x)
            report.audio_seconds += x.shape[-1] / self.sample_rate
            to_write.put((ex, x))
            report.done += 1

    def _enhance(self, datas, exs):
        if len(datas[0]) == 4:                                # prepared by the loader thread
            return self.enhancer.enhance_prepared_batch([d[3] for d in datas])
        obs = [d[0] for d in datas]
        acts = [d[1] for d in datas]
        spk = [d[2] for d in datas]
        return self.enhancer.enhance_observation_batch(obs, acts, spk, exs)

    def _fail(self, report, ex, err):
        print('ERROR: Failed example:', ex.get('example_id'), repr(err), flush=True)
        report.failed.append((ex.get('example_id'), repr(err)))
        if self.strict:
            raise err


def run_distributed(sched, examples, schedule='auto'):
    """Run `examples` (the WHOLE job, identical on every rank) through `sched` on this rank's share.

    schedule: 'dynamic' -- task farm (`sharding.WorkQueue`, the semantics of dlp_mpi.split_managed,
              core.py:381): every rank plans the same length-sorted batches and pulls the next one
              when it is free; 'lpt' -- static shards balanced by the known segment lengths;
              'strided' -- i % world == rank (kaldi_run.py:73-76); 'auto' -- dynamic when a
              process group with more than one rank is up, else lpt."""
    from . import sharding
    rank, world = sharding.init_process_group()
    if schedule == 'auto':
        schedule = 'dynamic' if world > 1 else 'lpt'
    if world == 1:
        return sched.run(examples)
    if schedule == 'dynamic':
        return sched.run(examples, queue=sharding.WorkQueue)
    lengths = [sched.example_length(ex) for ex in examples] if schedule == 'lpt' else None
    if schedule not in ('lpt', 'strided'):
        raise ValueError(schedule)
    mine = sharding.shard_indices(len(examples), rank, world, lengths=lengths)
    return sched.run([examples[i] for i in mine])


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
