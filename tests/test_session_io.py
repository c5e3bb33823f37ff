"""CPU: wav I/O restatement (audio_io.py vs the reference's doctests), batch planning and the
session driver's control flow (resume, failure isolation) with a stand-in enhancer."""
import struct
from pathlib import Path

import numpy as np
import pytest

from pb_chime5_b200 import audio_io
from pb_chime5_b200.session import SessionScheduler, plan_batches, stack_arrays

ROOT = Path(__file__).resolve().parent.parent


def test_dump_audio_matches_reference_doctest(tmp_path):
    # pb_chime5/io/audiowrite.py:32-38: [1, 2, -4, 4] -> peak normalised 16-bit PCM
    p = tmp_path / 'a.wav'
    audio_io.dump_audio(np.array([1, 2, -4, 4], dtype=np.float32), p)
    got = audio_io.load_audio(p)
    np.testing.assert_allclose(got, [0.24996948, 0.49996948, -0.99996948, 0.99996948], atol=1e-8)
    assert audio_io.wav_info(p).bits == 16 and audio_io.wav_info(p).channels == 1
    # audiowrite.py:40-46: without normalisation float values in [-1, 1) are written as they are
    a = np.arange(10, dtype=np.float32) / 32
    audio_io.dump_audio(a, p, normalize=False)
    np.testing.assert_array_equal(audio_io.load_audio(p), a.astype(np.float64))
    with pytest.raises(TypeError):
        audio_io.dump_audio(np.array(['a']), p)


def test_load_audio_slices_and_channels(tmp_path):
    rng = np.random.default_rng(0)
    x = (rng.integers(-32768, 32767, size=(3, 1000))).astype(np.int16)
    p = tmp_path / 'm.wav'
    audio_io.dump_audio(x, p, normalize=False)
    full = audio_io.load_audio(p)
    assert full.shape == (3, 1000) and full.dtype == np.float64
    np.testing.assert_array_equal(full, x / 32768.0)
    np.testing.assert_array_equal(audio_io.load_audio(p, start=100, stop=350), full[:, 100:350])
    np.testing.assert_array_equal(audio_io.load_audio(p, start=900, stop=5000), full[:, 900:])
    np.testing.assert_array_equal(audio_io.load_audio(p, start=10, frames=5), full[:, 10:15])
    np.testing.assert_array_equal(audio_io.load_audio(p, dtype=np.int16), x)
    y, sr = audio_io.load_audio(p, return_sample_rate=True)
    assert sr == 16000
    with pytest.raises(ValueError):
        audio_io.load_audio(p, expected_sample_rate=8000)
    out = np.empty((3, 250))
    assert audio_io.load_audio(p, start=100, stop=350, out=out) is out
    # IEEE float file with an extra chunk before the data
    f32 = np.linspace(-1, 1, 64, dtype='<f4')
    q = tmp_path / 'f.wav'
    body = (b'WAVE' + b'fmt ' + struct.pack('<IHHIIHH', 16, 3, 1, 16000, 64000, 4, 32) +
            b'LIST' + struct.pack('<I', 4) + b'abcd' + b'data' + struct.pack('<I', f32.nbytes) + f32.tobytes())
    q.write_bytes(b'RIFF' + struct.pack('<I', len(body)) + body)
    np.testing.assert_array_equal(audio_io.load_audio(q), f32.astype(np.float64))
    bad = tmp_path / 'bad.wav'
    bad.write_bytes(b'NIST_1A\n   1024\n' + bytes(100))
    with pytest.raises(RuntimeError):
        audio_io.load_audio(bad)


def test_plan_batches_properties():
    rng = np.random.default_rng(1)
    lengths = [int(v) for v in rng.integers(5000, 400000, size=203)]
    keys = [('S02',) if i % 3 else ('S09',) for i in range(203)]
    batches = plan_batches(lengths, keys, batch_size=8, window=64, max_batch_samples=8 * 200000)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(203))
    for b in batches:
        assert 1 <= len(b) <= 8
        assert len({keys[i] for i in b}) == 1
        assert len(b) * max(lengths[i] for i in b) <= 8 * 200000 or len(b) == 1
        assert [lengths[i] for i in b] == sorted((lengths[i] for i in b), reverse=True)
        assert max(b) // 64 == min(b) // 64                     # never crosses a window
    assert plan_batches(lengths, keys, 8, 64, 8 * 200000) == batches
    # padding waste stays small compared with taking the examples in order
    waste = sum(len(b) * lengths[b[0]] - sum(lengths[i] for i in b) for b in batches) / sum(lengths)
    naive = plan_batches(lengths, keys, 8, 1)                  # window 1: one example per batch
    assert len(naive) == 203 and waste < 0.25
    assert plan_batches([], None) == []


def test_stack_arrays():
    a, b = np.arange(8.).reshape(4, 2), np.arange(12.).reshape(4, 3)
    assert stack_arrays([a, b], True).shape == (8, 2)
    np.testing.assert_array_equal(stack_arrays([a, b], 'outer_array_mics'), np.concatenate([a[[0, -1], :2], b[[0, -1], :2]]))
    assert stack_arrays([a, b], 'first_array_mics').shape == (2, 2)
    with pytest.raises(ValueError):
        stack_arrays([a, b], 'nope')


class _FakeEnhancer:
    def __init__(self):
        self.calls = []

    def enhance_observation_batch(self, obs_list, acts, spk, exs=None):
        self.calls.append([e['example_id'] for e in exs])
        if any(e.get('poison') for e in exs):
            raise FloatingPointError('poisoned batch')
        return [0.5 * o[0] for o in obs_list]


def _examples(tmp_path, n):
    exs = []
    for i in range(n):
        exs.append({'example_id': f'ex{i}', 'session_id': 'S02', 'num_samples': 1000 + 37 * i,
                    'start': {'observation': {'U01': 0}}, 'end': {'observation': {'U01': 1000 + 37 * i}}})
    return exs


def test_session_scheduler_resume_and_failure_isolation(tmp_path):
    exs = _examples(tmp_path, 11)
    exs[4]['poison'] = True
    exs[7]['missing'] = True

    def load(ex):
        if ex.get('missing'):
            raise FileNotFoundError(ex['example_id'])
        n = ex['num_samples']
        return np.full((2, n), 0.25) * np.sign(np.sin(np.arange(n))), {'P05': np.ones(n, bool), 'Noise': np.ones(n, bool)}, 'P05'

    path = lambda ex: tmp_path / 'out' / f"{ex['example_id']}.wav"     # noqa: E731
    enh = _FakeEnhancer()
    rep = SessionScheduler(enh, load, path, batch_size=4, window=8, skip_existing=True).run(exs)
    assert rep.done == 9 and rep.skipped == 0
    assert sorted(e for e, _ in rep.failed) == ['ex4', 'ex7']
    for i in range(11):
        assert path(exs[i]).exists() == (i not in (4, 7))
    x = audio_io.load_audio(path(exs[3]))
    assert x.shape == (1000 + 37 * 3,) and abs(np.abs(x).max() - 32767 / 32768) < 1e-9      # dump_audio normalisation
    # the poisoned batch was retried example by example
    assert any(c == ['ex4'] for c in enh.calls)
    # resume: only the two failures are tried again
    for e in exs:
        e.pop('poison', None), e.pop('missing', None)
    enh2 = _FakeEnhancer()
    rep2 = SessionScheduler(enh2, load, path, batch_size=4, window=8, skip_existing=True).run(exs)
    assert rep2.skipped == 9 and rep2.done == 2 and not rep2.failed
    assert sorted(i for c in enh2.calls for i in c) == ['ex4', 'ex7']
    # strict mode keeps the reference behaviour: the first failure raises
    exs[0]['missing'] = True
    with pytest.raises(FileNotFoundError):
        SessionScheduler(_FakeEnhancer(), load, path, batch_size=4, skip_existing=False, strict=True).run(exs)


def _fake_chime6(tmp_path, total=60000):
    """audio/dev/S02_U0{1,2}.CH{1,2}.wav + an RTTM with two speakers"""
    rng = np.random.default_rng(3)
    d = tmp_path / 'CHiME6' / 'audio' / 'dev'
    d.mkdir(parents=True)
    src = rng.standard_normal((2, total)) * (np.sin(np.arange(total) / 700.0 + np.array([[0.0], [2.0]])) > 0)
    for a in ('U01', 'U02'):
        for c in ('CH1', 'CH2'):
            x = rng.standard_normal(2) @ src + 0.05 * rng.standard_normal(total)
            audio_io.dump_audio(0.3 * x / np.abs(x).max(), d / f'S02_{a}.{c}.wav', normalize=False)
    audio_io.dump_audio(np.zeros(10), d / 'S02_P05.wav', normalize=False)     # worn microphone: ignored
    rttm = tmp_path / 'rttm'
    rttm.write_text('SPEAKER S02_U06.ENH 1 0.50 0.60 <NA> <NA> 1 <NA> <NA>\n'
                    'SPEAKER S02_U06.ENH 1 1.40 0.45 <NA> <NA> 2 <NA> <NA>\n'
                    'SPEAKER S02_U06.ENH 1 2.20 0.70 <NA> <NA> 1 <NA> <NA>\n')
    return tmp_path / 'CHiME6', rttm


def test_rttm_front_door_plumbing(tmp_path):
    from pb_chime5_b200 import core_chime6_rttm as r
    chime6_dir, rttm = _fake_chime6(tmp_path)
    assert r.parse_rttm(rttm) == {'S02_U06.ENH': {'1': [(8000, 17600), (35200, 46400)], '2': [(22400, 29600)]}}
    assert r.RTTMDatabase.example_id('S02', 1, 100, 200) == 'S02_U06.-1-000000100_000000200'     # rttm.py:437
    files = r.get_chime6_files(chime6_dir)
    assert sorted(files['S02']) == ['U01', 'U02'] and [Path(f).name for f in files['S02']['U01']] == ['S02_U01.CH1.wav', 'S02_U01.CH2.wav']
    assert len(r.get_chime6_files(chime6_dir, flat=True)['S02']) == 4
    assert list(r.get_chime6_files(chime6_dir, worn=True)['S02']) == ['P05']
    db = r.get_database(chime6_dir, rttm, 'first_array_mics')
    exs = db.get_dataset_for_session('S02', audio_read=True, context_samples=4000)
    assert [e['speaker_id'] for e in exs] == ['1', '1', '2'] and exs[0]['example_id'] == 'S02_U06.-1-000008000_000017600'
    ex = exs[1]
    assert (ex['start'], ex['end'], ex['start_orig'], ex['num_samples_orig']) == (31200, 50400, 35200, 11200)
    assert ex['audio_data'].shape == (2, 19200) and [Path(p).name for p in ex['audio_path']] == ['S02_U01.CH1.wav', 'S02_U02.CH1.wav']
    act = r.Activity(garbage_class=True, rttm=str(rttm))['S02']
    assert list(act) == ['1', '2', 'Noise']
    a = act['1'][7990:8010]
    assert a.tolist() == [False] * 10 + [True] * 10 and act['Noise'][0:5].all() and not act['2'][0:100].any()
    assert not r.Activity(garbage_class=False, rttm=str(rttm))['S02']['Noise'][0:9].any()
    assert 'Noise' not in r.Activity(garbage_class=None, rttm=str(rttm))['S02']
    sig = r.signature_defaults()
    assert list(sig)[:5] == ['database_rttm', 'activity_rttm', 'chime6_dir', 'multiarray', 'context_samples']
    assert sig['multiarray'] == 'outer_array_mics' and sig['activity_garbage_class'] is True


def test_two_rank_session_task_farm(tmp_path):
    """world_size 2 on CPU (gloo), launched the way `mpiexec -np 2` launches the reference (OMPI_*
    variables, no RANK / WORLD_SIZE): `run_distributed` binds the ranks, every rank plans the same
    length-sorted batches and pulls them from the shared work queue (dlp_mpi.split_managed semantics,
    core.py:381).  Every example is enhanced exactly once; a rank that is 20x slower per batch ends up
    with far fewer batches instead of holding the job (it can hold at most prefetch + 1 claimed
    batches at the end)."""
    import json
    import os
    import subprocess
    import sys
    script = tmp_path / 'farm.py'
    script.write_text('''
import json, sys, time
sys.path.insert(0, %r)
import numpy as np
from pb_chime5_b200 import sharding
from pb_chime5_b200.session import SessionScheduler, run_distributed

class Fake:
    def __init__(self, delay): self.delay, self.calls = delay, []
    def enhance_observation_batch(self, obs, acts, spk, exs=None):
        time.sleep(self.delay)
        self.calls.append([e['example_id'] for e in exs])
        return [o[0] for o in obs]

rank, world = sharding.init_process_group('gloo')
exs = [{'example_id': f'ex{i}', 'session_id': 'S02', 'num_samples': 800 + 13 * i,
        'start': {'observation': {'U01': 0}}, 'end': {'observation': {'U01': 800 + 13 * i}}} for i in range(120)]
load = lambda ex: (np.zeros((2, ex['num_samples'])), {'P05': np.ones(ex['num_samples'], bool)}, 'P05')
got = {}
enh = Fake(0.2 if rank == 1 else 0.01)
sched = SessionScheduler(enh, load, None, batch_size=4, window=64, prefetch=1, skip_existing=False, strict=True,
                         sink_fn=lambda ex, x: got.__setitem__(ex['example_id'], x.shape[-1]))
rep = run_distributed(sched, exs, sys.argv[1])
sharding.barrier()
print(json.dumps({'rank': rank, 'world': world, 'done': rep.done, 'batches': rep.batches, 'ids': sorted(got),
                  'lens_ok': all(got[e['example_id']] == e['num_samples'] for e in exs if e['example_id'] in got)}), flush=True)
sharding.shutdown()
''' % str(ROOT))
    for schedule, port in (('dynamic', '29541'), ('lpt', '29542')):
        env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
        env.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=port, OMPI_COMM_WORLD_SIZE='2')
        procs = [subprocess.Popen([sys.executable, str(script), schedule],
                                  env=dict(env, OMPI_COMM_WORLD_RANK=str(r), OMPI_COMM_WORLD_LOCAL_RANK=str(r)),
                                  stdout=subprocess.PIPE, text=True) for r in range(2)]
        outs = [json.loads(p.communicate(timeout=180)[0].strip().splitlines()[-1]) for p in procs]
        assert all(p.returncode == 0 for p in procs)
        outs.sort(key=lambda o: o['rank'])
        assert [o['world'] for o in outs] == [2, 2] and all(o['lens_ok'] for o in outs)
        assert sorted(outs[0]['ids'] + outs[1]['ids']) == sorted(f'ex{i}' for i in range(120))    # exactly once
        assert outs[0]['done'] + outs[1]['done'] == 120
        if schedule == 'dynamic':
            assert outs[0]['batches'] >= 3 * outs[1]['batches'] > 0, outs       # the fast rank took most of the work
        else:
            assert abs(outs[0]['done'] - outs[1]['done']) <= 2                  # static balanced shards


def test_session_scheduler_strict_failure_leaves_no_thread_behind(tmp_path):
    """strict mode: the first failure raises and the loader thread ends (it used to stay blocked on
    the full prefetch queue, holding the prefetched audio)."""
    import threading
    exs = _examples(tmp_path, 40)
    exs[0]['missing'] = True

    def load(ex):
        if ex.get('missing'):
            raise FileNotFoundError(ex['example_id'])
        n = ex['num_samples']
        return np.zeros((2, n)), {'P05': np.ones(n, bool)}, 'P05'

    before = threading.active_count()
    with pytest.raises(FileNotFoundError):
        SessionScheduler(_FakeEnhancer(), load, None, batch_size=2, prefetch=1, skip_existing=False, strict=True,
                         sink_fn=lambda ex, x: None).run(exs)
    import time
    time.sleep(0.5)
    assert threading.active_count() <= before


def test_session_batches_split_by_channel_and_class_count(tmp_path):
    """examples of one session with a different channel count (a missing array) or class count are
    enhanced in their own pass instead of tripping the batch (strict mode included)"""
    exs = _examples(tmp_path, 6)

    def load(ex):
        n = ex['num_samples']
        d = 3 if ex['example_id'] == 'ex2' else 2
        acts = {'P05': np.ones(n, bool)}
        if ex['example_id'] == 'ex4':
            acts['P06'] = np.ones(n, bool)
        return np.zeros((d, n)), acts, 'P05'

    class Enh(_FakeEnhancer):
        def enhance_observation_batch(self, obs_list, acts, spk, exs=None):
            assert len({o.shape[0] for o in obs_list}) == 1 and len({len(a) for a in acts}) == 1
            return super().enhance_observation_batch(obs_list, acts, spk, exs)

    enh = Enh()
    rep = SessionScheduler(enh, load, None, batch_size=8, skip_existing=False, strict=True,
                           sink_fn=lambda ex, x: None).run(exs)
    assert rep.done == 6 and not rep.failed
    assert sorted(map(sorted, enh.calls)) == [['ex0', 'ex1', 'ex3', 'ex5'], ['ex2'], ['ex4']]
    assert 0 < rep.padding_efficiency <= 1
