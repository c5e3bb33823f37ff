"""CHiME-6 RTTM front door: segments and speaker activity from RTTM files, audio from a list of
single-channel wavs per session (row f4 of SURVEY.md section 8).

Mirrors ``pb_chime5/core_chime6_rttm.py`` (Activity :31-69, Enhancer :72-282, get_database
:286-357, get_enhancer :360-422) and the parts of ``pb_chime5/database/chime5/rttm.py`` it
drives (``RTTMDatabase`` :285-547, ``get_chime6_files`` :21-117), which
``scripts/kaldi_run_rttm.py`` launches.  The numeric blocks are the classes of
``pb_chime5_b200.core``; everything here is plumbing.

Third-party pieces the reference takes from ``paderbox`` / ``lazy_dataset`` (not vendored, not
installed) are restated minimally: RTTM parsing into per-speaker sample intervals
(``paderbox.array.intervall.from_rttm``: ``SPEAKER <file> <chan> <onset s> <duration s> ... <name>``,
16 kHz) and a sparse boolean activity that can be sliced (``ArrayIntervall``).  Sample indices are
``round(seconds * 16000)``; this rounding is unpinned by the reference tree.
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from functools import cached_property
from pathlib import Path

import numpy as np

from . import core_chime6 as _c6
from .core import GSS, WPE, Beamformer  # noqa: F401  (re-exported API)

SAMPLE_RATE = 16000


class IntervalActivity:
    """Sparse boolean sample activity: sorted (start, end) intervals; ``a[i:j]`` -> dense bool array."""

    def __init__(self, intervals=(), default=False):
        self.intervals = sorted((int(s), int(e)) for s, e in intervals)
        self.default = bool(default)

    @classmethod
    def zeros(cls):
        return cls()

    @classmethod
    def ones(cls):
        return cls(default=True)

    def __getitem__(self, item):
        if not isinstance(item, slice) or item.step not in (None, 1):
            raise TypeError(item)
        start, stop = int(item.start or 0), int(item.stop)
        out = np.full(max(stop - start, 0), self.default, dtype=bool)
        for s, e in self.intervals:
            if e <= start:
                continue
            if s >= stop:
                break
            out[max(s, start) - start:min(e, stop) - start] = True
        return out


def parse_rttm(paths, sample_rate=SAMPLE_RATE):
    """{file id: {speaker: [(start, end), ...]}} in samples from one or several RTTM files."""
    if isinstance(paths, (str, Path)):
        paths = [paths]
    data = {}
    for p in paths:
        for line in Path(p).read_text().splitlines():
            f = line.split()
            if not f or f[0] != 'SPEAKER':
                continue
            file_id, onset, dur, name = f[1], float(f[3]), float(f[4]), f[7]
            start = int(round(onset * sample_rate))
            end = int(round((onset + dur) * sample_rate))
            data.setdefault(file_id, {}).setdefault(name, []).append((start, end))
    return data


def _session_keys(data):
    """The kaldi scripts append postfixes to the file id (S02_U06.ENH, S02_U06): drop them
    (core_chime6_rttm.py:46-56, rttm.py:415-423)."""
    out = {k.replace('_U06', '').replace('.ENH', ''): v for k, v in data.items()}
    assert len(out) == len(data), (tuple(out), tuple(data))
    return out


@dataclass
class Activity:
    """core_chime6_rttm.py:31-69"""
    garbage_class: bool = False
    rttm: str = None

    @cached_property
    def _data(self):
        return _session_keys(parse_rttm(self.rttm))

    def __getitem__(self, session_id):
        data = {k: IntervalActivity(v) for k, v in self._data[session_id].items()}
        if self.garbage_class is False:
            data['Noise'] = IntervalActivity.zeros()
        elif self.garbage_class is True:
            data['Noise'] = IntervalActivity.ones()
        elif self.garbage_class is not None:
            raise ValueError(self.garbage_class)
        return data


def get_chime6_files(chime6_dir, worn=False, flat=False):
    """rttm.py:21-117: {session: {array: [CH1..CH4 wavs]}} (or a flat sorted list per session)."""
    chime6_dir = Path(chime6_dir)
    files = {}
    for p in sorted(chime6_dir.glob('audio/*/*.wav')):
        session_id, rest = p.name.split('_', 1)
        is_worn = rest.startswith('P')
        if is_worn != worn:
            continue
        if worn:
            files.setdefault(session_id, {})[rest.split('.')[0]] = str(p)
        else:
            files.setdefault(session_id, {}).setdefault(rest.split('.')[0], []).append(str(p))
    if flat and not worn:
        files = {s: [f for a in sorted(v) for f in v[a]] for s, v in files.items()}
    return files


class RTTMDatabase:
    """rttm.py:285-547 without lazy_dataset: examples are plain dicts in a list."""

    def __init__(self, rttm_path, audio_paths, alias=None):
        self._rttm_path, self._audio_paths, self._alias = rttm_path, audio_paths, alias

    @cached_property
    def _rttm(self):
        return _session_keys(parse_rttm(self._rttm_path))

    @staticmethod
    def example_id(file_id, speaker_id, start, end):
        """rttm.py:427-458: 'S02_U06.-1-000000100_000000200' (the kaldi recipe needs the U06)"""
        digits = len(str(16000 * 60 * 60 * 10))
        return f'{file_id}_U06.-{speaker_id}-{str(start).zfill(digits)}_{str(end).zfill(digits)}'

    def get_dataset_for_session(self, session, *, audio_read=False, adjust_times=False, context_samples=0,
                                equal_start_context=False):
        if adjust_times:
            raise ValueError(adjust_times)                # rttm.py:513-522: undefined without transcriptions
        sessions = (session,) if isinstance(session, str) else tuple(session)
        out = []
        for session_id in sessions:
            for speaker_id, intervals in self._rttm[session_id].items():
                for start, end in intervals:
                    ex = {'example_id': self.example_id(session_id, speaker_id, start, end),
                          'start': start, 'end': end, 'num_samples': end - start,
                          'session_id': session_id, 'speaker_id': speaker_id,
                          'audio_path': self._audio_paths[session_id], 'dataset': session_id}
                    if context_samples != 0:
                        # backup_orig_start_end + AddContext (database.py:713-1053, flat indices)
                        ex['start_orig'], ex['end_orig'], ex['num_samples_orig'] = start, end, end - start
                        ex['start'] = max(start - context_samples, 0)
                        ex['end'] = end + context_samples
                        ex['num_samples'] = ex['end'] - ex['start']
                    out.append(ex)
        out.sort(key=lambda e: e['example_id'])
        if audio_read is True:
            for ex in out:
                ex['audio_data'] = load_audio_files(ex['audio_path'], ex['start'], ex['end'])
        elif audio_read is not False:
            raise TypeError(audio_read)
        return out

    get_iterator_for_session = get_dataset_for_session


def load_audio_files(paths, start, stop):
    """rttm.py:550-600 (recursive_load_audio): one (N,) array per single-channel file, stacked;
    files may end before `stop` (context beyond the recording): cut to the shortest."""
    from .audio_io import load_audio
    xs = [load_audio(p, start=start, stop=stop) for p in paths]
    n = min(x.shape[-1] for x in xs)
    return np.stack([x[..., :n] for x in xs], axis=0)


def get_database(chime6_dir, rttm, multiarray):
    """core_chime6_rttm.py:286-357"""
    chime6_dir = Path(chime6_dir)
    if multiarray is True:
        audio_paths = get_chime6_files(chime6_dir, worn=False, flat=True)
    elif multiarray == 'outer_array_mics':
        audio_paths = {s: [f for a in sorted(v) for f in (v[a][0], v[a][-1])]
                       for s, v in get_chime6_files(chime6_dir).items()}
    elif multiarray == 'first_array_mics':
        audio_paths = {s: [v[a][0] for a in sorted(v)] for s, v in get_chime6_files(chime6_dir).items()}
    else:
        raise ValueError(multiarray)
    alias = {}
    for p in sorted(chime6_dir.glob('transcriptions/*/*.json')):
        alias.setdefault(p.parts[-2], []).append(p.with_suffix('').name)
    return RTTMDatabase(rttm, audio_paths, alias=alias)


@dataclass
class Enhancer(_c6.Enhancer):
    """core_chime6_rttm.py:72-282: the database is a field, examples may carry 'audio_data'."""
    db: RTTMDatabase = None

    def get_iterator(self, session_id):
        return self.db.get_iterator_for_session(
            session_id, audio_read=False, adjust_times=False, context_samples=self.context_samples,
            equal_start_context=False)

    def _load_example(self, ex):
        array_start, array_end = ex['start'], ex['end']
        ex_array_activity = {k: arr[array_start:array_end] for k, arr in self.activity[ex['session_id']].items()}
        obs = ex['audio_data'] if 'audio_data' in ex else load_audio_files(ex['audio_path'], array_start, array_end)
        n = obs.shape[-1]
        if n < array_end - array_start:                       # context beyond the end of the recording
            ex_array_activity = {k: v[:n] for k, v in ex_array_activity.items()}
        return obs, ex_array_activity, ex['speaker_id']

    def enhance_session(self, session_ids, audio_dir, dataset_slice=False, audio_dir_exist_ok=False,
                        batch_size=8, skip_existing=False, strict=True, schedule='auto'):
        """core_chime6_rttm.py:137-185: one directory per session ('dataset' = session id)."""
        from . import sharding
        from .session import SessionScheduler, run_distributed
        audio_dir = Path(audio_dir)
        it = self.get_iterator(session_ids)
        rank, world = sharding.init_process_group()      # binds the GPU of this rank (torchrun / mpiexec / srun)
        if rank == 0:
            audio_dir.mkdir(exist_ok=audio_dir_exist_ok or skip_existing)
        sharding.barrier()
        if dataset_slice is not False:
            if dataset_slice is True:
                it = it[:2]
            elif isinstance(dataset_slice, (int, slice)):
                it = it[:dataset_slice] if isinstance(dataset_slice, int) else it[dataset_slice]
            else:
                raise ValueError(dataset_slice)
        examples = [it[i] for i in range(len(it))]
        sched = SessionScheduler(self, self._load_example,
                                 lambda ex: audio_dir / ex['dataset'] / f'{ex["example_id"]}.wav',
                                 self._finish_example, batch_size=batch_size, skip_existing=skip_existing,
                                 strict=strict)
        return run_distributed(sched, examples, schedule)


def get_enhancer(
    database_rttm,
    activity_rttm,
    chime6_dir='/net/fastdb/chime6/CHiME6',
    multiarray='outer_array_mics',
    context_samples=240000,

    wpe=True,
    wpe_tabs=10,
    wpe_delay=2,
    wpe_iterations=3,
    wpe_psd_context=0,

    activity_garbage_class=True,

    stft_size=1024,
    stft_shift=256,
    stft_fading=True,

    bss_iterations=20,
    bss_iterations_post=1,

    bf_drop_context=True,

    bf='mvdrSouden_ban',
    postfilter=None,
):
    """core_chime6_rttm.py:360-422"""
    assert wpe is True or wpe is False, wpe
    return Enhancer(
        db=get_database(chime6_dir, database_rttm, multiarray),
        multiarray=multiarray,
        reference_array=None,
        context_samples=context_samples,
        wpe_block=WPE(taps=wpe_tabs, delay=wpe_delay, iterations=wpe_iterations,
                      psd_context=wpe_psd_context) if wpe else None,
        activity=Activity(garbage_class=activity_garbage_class, rttm=activity_rttm),
        gss_block=GSS(iterations=bss_iterations, iterations_post=bss_iterations_post, verbose=False),
        bf_drop_context=bf_drop_context,
        bf_block=Beamformer(type=bf, postfilter=postfilter),
        stft_size=stft_size,
        stft_shift=stft_shift,
        stft_fading=stft_fading,
    )


def signature_defaults():
    return {k: v.default for k, v in inspect.signature(get_enhancer).parameters.items()}
