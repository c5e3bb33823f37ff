"""CPU: host-side mirror of the reference interface, the C-ABI surface and the
utterance sharding (no kernels are launched here)."""
import ctypes
import inspect
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import gss_oracle as oracle
import torch

from pb_chime5_b200 import _lib, core, sharding, synth

ROOT = Path(__file__).resolve().parent.parent


def test_abi_exports_every_declared_symbol():
    header = (ROOT / 'include' / 'gss.h').read_text()
    declared = set(re.findall(r'\b(gss_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no prototypes found in include/gss.h'
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        getattr(handle, name)            # raises AttributeError if not exported
    assert _lib.lib().gss_version() >= 100
    # the developer entry points live in libgss_dev.so only: the product library exports none of them
    dev_header = (ROOT / 'include' / 'gss_dev.h').read_text()
    dev_declared = set(re.findall(r'\b(gss_debug_[a-z0-9_]+)\s*\(', dev_header))
    assert dev_declared == set(_lib.dev_exported_symbols()), dev_declared ^ set(_lib.dev_exported_symbols())
    dev_handle = ctypes.CDLL(str(_lib.DEV_LIB_PATH))
    for name in dev_declared:
        getattr(dev_handle, name)
        assert not hasattr(handle, name), f'{name} leaked into the product library'
    for name in declared:
        getattr(dev_handle, name)        # the developer library is a superset


def test_abi_argument_errors_map_to_reference_exceptions():
    L = _lib.lib()
    n = ctypes.c_size_t(0)
    assert L.gss_workspace_bytes(_lib.OP_WPE, 1, 513, 24, 941, 5, 10, ctypes.byref(n)) == 0
    assert n.value > 513 * 240 * 240 * 16
    # null pointers / limits never reach a launch
    with pytest.raises(AssertionError):
        _lib.check(L.gss_cacgmm_c64(None, None, None, 1, 1, 1e-10, 1e-10, 1, 1, 4, 10, 3, 10, None,
                                    None, None, None, None, None, 0, None))
    one = ctypes.c_void_p(16)            # dummy non-null, never dereferenced on the host
    with pytest.raises(AssertionError, match='sure'):      # cacgmm.py:248  D < 35
        _lib.check(L.gss_cacgmm_c64(one, one, one, 1, 1, 1e-10, 1e-10, 1, 1, 35, 10, 3, 10, None,
                                    None, None, None, None, None, 0, None))
    with pytest.raises(AssertionError, match='sure'):      # cacgmm.py:247  K < 20
        _lib.check(L.gss_cacgmm_c64(one, one, one, 1, 1, 1e-10, 1e-10, 1, 1, 4, 10, 20, 10, None,
                                    None, None, None, None, None, 0, None))
    with pytest.raises(NotImplementedError):               # iterations_post == 0 (core.py:198-202)
        _lib.check(L.gss_cacgmm_c64(one, one, one, 1, 0, 1e-10, 1e-10, 1, 1, 4, 10, 3, 10, None,
                                    None, None, None, None, None, 0, None))
    with pytest.raises(AssertionError):                    # beamforming_wrapper.py:44  D < 30
        _lib.check(L.gss_beamform_c64(one, one, one, one, 0, 0, 0, 1, 1, 30, 10, None, None, None, None, one, 1 << 30, None))
    with pytest.raises(NotImplementedError):               # unknown beamformer type
        _lib.check(L.gss_beamform_c64(one, one, one, one, 17, 0, 0, 1, 1, 4, 10, None, None, None, None, one, 1 << 30, None))
    with pytest.raises(RuntimeError, match='workspace'):
        _lib.check(L.gss_beamform_c64(one, one, one, one, 0, 0, 0, 1, 1, 4, 10, None, None, None, None, one, 8, None))


def test_get_enhancer_signature_matches_reference():
    expect = dict(
        multiarray=False, reference_array=None, context_samples=240000, wpe=True, wpe_tabs=10,
        wpe_delay=2, wpe_iterations=3, wpe_psd_context=0, activity_type='annotation',
        activity_path=None, activity_garbage_class=True, stft_size=1024, stft_shift=256,
        stft_fading=True, bss_iterations=20, bss_iterations_post=1, bf_drop_context=True,
        bf='mvdrSouden_ban', postfilter=None)
    got = core.signature_defaults()
    got.pop('database_path')
    assert got == expect                                   # names, order-insensitive defaults
    assert list(inspect.signature(core.get_enhancer).parameters)[:len(expect)] == list(expect)
    e = core.get_enhancer()
    assert e.wpe_block == core.WPE(taps=10, delay=2, iterations=3, psd_context=0)
    assert e.gss_block == core.GSS(iterations=20, iterations_post=1, verbose=False)
    assert e.bf_block == core.Beamformer(type='mvdrSouden_ban', postfilter=None)
    assert e.activity.garbage_class is True and e.bf_drop_context is True
    assert core.get_enhancer(wpe=False).wpe_block is None
    with pytest.raises(AssertionError):
        core.get_enhancer(wpe=1)
    with pytest.raises(AssertionError):
        core.get_enhancer(activity_path='x')


def test_reference_signature_if_available():
    from oracle import refboot
    if not refboot.available():
        pytest.skip('reference tree not present (GPU box)')
    code = ("import warnings; warnings.filterwarnings('ignore');"
            "from oracle import refboot; refboot.boot();"
            "import inspect, pb_chime5.core as c, json;"
            "print(json.dumps({k: repr(v.default) for k, v in inspect.signature(c.get_enhancer).parameters.items() if k != 'database_path'}));"
            "print(json.dumps([f.name for f in __import__('dataclasses').fields(c.Enhancer)]))")
    out = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, check=True).stdout.splitlines()
    import json, dataclasses
    ref_defaults = json.loads(out[-2])
    mine = {k: repr(v) for k, v in core.signature_defaults().items() if k != 'database_path'}
    assert mine == ref_defaults
    assert [f.name for f in dataclasses.fields(core.Enhancer)] == json.loads(out[-1])


def test_chime6_front_door_signature_if_available():
    """pb_chime5/core_chime6.py:573-635: keyword order and defaults are API (sacred reads them)"""
    from pb_chime5_b200 import core_chime6
    assert core_chime6.WPE is core.WPE and core_chime6.GSS is core.GSS and core_chime6.Beamformer is core.Beamformer
    assert list(core_chime6.signature_defaults())[:3] == ['multiarray', 'context_samples', 'reference_array']
    assert core_chime6.signature_defaults()['database_path'].endswith('chime6.json')
    ex = {'start': 1000, 'end': 9000, 'start_orig': 3000, 'end_orig': 8000}
    assert core_chime6.start_end_context_frames(ex, 1024, 256, True) == \
        (core.samples_to_stft_frames(2000, 1024, 256, fading=True), core.samples_to_stft_frames(1000, 1024, 256, fading=True))
    from oracle import refboot
    if not refboot.available():
        return
    code = ("import warnings; warnings.filterwarnings('ignore');"
            "from oracle import refboot; refboot.boot();"
            "import inspect, pb_chime5.core_chime6 as c, json;"
            "print(json.dumps([[k, repr(v.default)] for k, v in inspect.signature(c.get_enhancer).parameters.items() if k != 'database_path']))")
    out = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, check=True).stdout.splitlines()
    import json
    mine = [[k, repr(v)] for k, v in core_chime6.signature_defaults().items() if k != 'database_path']
    assert mine == json.loads(out[-1])


def test_activity_framing_matches_oracle_and_doctest():
    vad = np.array([0, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0, 0, 0])
    assert core.activity_time_to_frequency(vad, 4, 2, True).tolist() == \
        [False, False, True, True, True, True, False, False]          # database.py:432-433
    assert core.activity_time_to_frequency([vad, vad], 4, 2, False).tolist() == \
        [[False, True, True, True, True, False]] * 2                  # database.py:447-449
    assert core.activity_time_to_frequency(np.zeros(200000), 1024, 256, False, False).shape == (778,)
    rng = np.random.default_rng(0)
    for n in (1, 700, 1024, 1025, 5000, 20001):
        sa = rng.random((3, n)) < 0.002
        for fading in (True, False):
            for pad in (True, False):
                if not pad and n + 2 * 768 * fading < 1024:
                    continue
                a = core.activity_time_to_frequency(sa, 1024, 256, fading, pad)
                b = oracle.activity_time_to_frequency(sa, 1024, 256, fading, pad)
                assert np.array_equal(a, b), (n, fading, pad)


def test_context_frames():
    ex = {'start': {'original': 0}, 'start_orig': {'original': 240000},
          'end': {'original': 720000}, 'end_orig': {'original': 480000}}
    assert core.start_end_context_frames(ex, 1024, 256, True) == (941, 941)
    assert core.samples_to_stft_frames(240000, 1024, 256, fading=True) == 941
    assert core.samples_to_stft_frames(0, 1024, 256, fading=True) == 3
    ex['start_orig']['original'] = -1
    with pytest.raises(AssertionError):
        core.start_end_context_frames(ex, 1024, 256, True)


def test_block_argument_errors_without_gpu():
    bf = core.Beamformer(type='lcmv', postfilter=None)
    with pytest.raises(NotImplementedError):
        bf(np.zeros((4, 10, 3), complex), np.zeros((10, 3)), np.zeros((10, 3)))
    bf = core.Beamformer(type='sum', postfilter='wiener')
    with pytest.raises(NotImplementedError):
        bf(np.zeros((4, 10, 3), complex), np.zeros((10, 3)), np.zeros((10, 3)))
    w = core.WPE(10, 2, 3, 0)
    with pytest.raises(NotImplementedError):
        w(np.zeros((4, 10), complex))
    with pytest.raises(AssertionError):
        w(np.zeros((4, 10, 3), complex), stack=True)


def test_shard_indices():
    assert sharding.shard_indices(10, 1, 4) == [1, 5, 9]                       # kaldi_run.py:73-76
    allidx = sorted(sum((sharding.shard_indices(23, r, 4) for r in range(4)), []))
    assert allidx == list(range(23))
    lengths = [900, 100, 100, 100, 500, 400, 50, 50]
    parts = [sharding.shard_indices(8, r, 2, lengths) for r in range(2)]
    assert sorted(parts[0] + parts[1]) == list(range(8))
    loads = [sum(lengths[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 100
    assert sharding.shard_indices(0, 0, 2) == []
    assert sharding.shard_indices(1, 1, 2) == []


def test_two_rank_gloo_sharding(tmp_path):
    """world_size 2 on CPU (gloo): broadcast of the work list + strided shards +
    max-over-ranks timing reduction."""
    script = tmp_path / 'w.py'
    script.write_text('''
import os, sys, json
sys.path.insert(0, %r)
from pb_chime5_b200 import sharding
rank, world = sharding.init_process_group('gloo')
work = sharding.broadcast_work_list([f"utt{i}" for i in range(7)] if rank == 0 else None)
mine = [work[i] for i in sharding.shard_indices(len(work), rank, world)]
t = sharding.max_over_ranks(1.0 + rank)
n = sharding.sum_over_ranks(len(mine))
sharding.barrier()
print(json.dumps({"rank": rank, "mine": mine, "t": t, "n": n}), flush=True)
''' % str(ROOT))
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29531', WORLD_SIZE='2')
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, text=True) for r in range(2)]
    import json
    outs = [json.loads(p.communicate(timeout=120)[0].strip().splitlines()[-1]) for p in procs]
    assert all(p.returncode == 0 for p in procs)
    outs.sort(key=lambda o: o['rank'])
    assert outs[0]['mine'] == ['utt0', 'utt2', 'utt4', 'utt6']
    assert outs[1]['mine'] == ['utt1', 'utt3', 'utt5']
    assert outs[0]['t'] == outs[1]['t'] == 2.0
    assert outs[0]['n'] == 7


def test_synth_is_seeded_and_shaped():
    a1, m1 = synth.make_utterance(3, D=4, T=50, F=5, K=3)
    a2, m2 = synth.make_utterance(3, D=4, T=50, F=5, K=3)
    assert a1.dtype == np.complex64 and a1.shape == (4, 50, 5) and m1.shape == (3, 50)
    assert np.array_equal(a1, a2) and np.array_equal(m1, m2)
    assert m1[-1].all()                              # 'Noise' garbage class always active
    assert (m1[:-1].sum(axis=1) >= 8).all()          # >= 2*D active frames per speaker


def test_int8_digit_split_scheme_bounds():
    """The arithmetic of csrc/wpe_gram_i8.cu restated in numpy (no GPU): 40-bit fixed point per
    power-whitened row, five balanced base-256 digits, digit pairs of order p + q <= 4, exact integer
    sums.  Checks the digit identity, the INT32 accumulator range and the error bound the design
    quotes (<= 1e-9 sqrt(R_ii R_jj) even for white, heavy-tailed power envelopes)."""
    rng = np.random.default_rng(0)
    D, T, L, delay, NS = 6, 500, 8, 2, 5
    Y = (rng.standard_normal((D, T)) + 1j * rng.standard_normal((D, T))) * np.exp(2.0 * rng.standard_normal(T))
    Y = Y.astype(np.complex64).astype(np.complex128)
    lam = np.mean(np.abs(Y) ** 2, axis=0)
    mu = np.sqrt(1.0 / np.maximum(lam, 1e-10 * lam.max()))
    rows = [Y]
    for k in range(L):
        r = np.zeros((D, T), complex)
        r[:, delay + k:] = Y[:, :T - delay - k]
        rows.append(r)
    U = np.concatenate(rows, 0) * mu
    Rex = U @ U.conj().T
    mrow = np.maximum(np.abs(U.real).max(1), np.abs(U.imag).max(1))
    sc = 2.0 ** (37 - np.floor(np.log2(mrow)))

    def planes(x):
        X = np.rint(x * sc[:, None]).astype(np.int64)
        assert np.abs(X).max() < 2 ** 38
        Xb = X + 0x8080808080                                  # the bias folded into the FMA's magic constant
        dig = [(((Xb >> (8 * s)) & 0xFF) ^ 0x80).astype(np.uint8).view(np.int8).astype(np.int64) for s in range(NS)]
        assert all((-128 <= d).all() and (d <= 127).all() for d in dig)
        assert (sum(d << (8 * s) for s, d in enumerate(dig)) == X).all()
        return dig[::-1]                                       # plane p = digit of weight 256^(4 - p)

    Ar, Ai = planes(U.real), planes(U.imag)

    def prod(A, B):
        acc = [np.zeros((A[0].shape[0],) * 2, np.int64) for _ in range(NS)]
        for p in range(NS):
            for q in range(NS - p):
                acc[p + q] += A[p] @ B[q].T
        assert max(np.abs(a).max() for a in acc) < 2 ** 31     # INT32 accumulators of the tensor core
        v = acc[0].astype(np.float64)
        for o in range(1, NS):
            v = v * 256.0 + acc[o]                             # float64 Horner of the epilogue
        return v * 2.0 ** 32 / np.outer(sc, sc)

    R = (prod(Ar, Ar) + prod(Ai, Ai)) + 1j * (prod(Ai, Ar) - prod(Ar, Ai))
    d = np.sqrt(np.diag(Rex).real)
    assert (np.abs(R - Rex) / np.outer(d, d)).max() < 1e-9
    assert np.abs(np.diag(R).imag).max() == 0.0


def test_cfg4_sized_work_list_shards_evenly():
    """BASELINE configs[3]: 20 000 dev-shaped utterances over 8 ranks.  Every utterance exactly once;
    with known lengths the greedy longest-first deal balances the audio seconds to < 0.1 %, the strided
    deal (kaldi_run.py:73-76) the utterance counts to +-1; session batches cover a shard exactly once."""
    from pb_chime5_b200.session import plan_batches
    rng = np.random.default_rng(4)
    n, world = 20000, 8
    lengths = (np.clip(rng.lognormal(np.log(2.0), 0.8, size=n), 0.3, 20.0) * 16000 + 2 * 240000).astype(int).tolist()
    shards = [sharding.shard_indices(n, r, world, lengths) for r in range(world)]
    assert sorted(i for s in shards for i in s) == list(range(n))
    load = [sum(lengths[i] for i in s) for s in shards]
    assert (max(load) - min(load)) / max(load) < 1e-3
    strided = [sharding.shard_indices(n, r, world) for r in range(world)]
    assert sorted(i for s in strided for i in s) == list(range(n))
    assert max(len(s) for s in strided) - min(len(s) for s in strided) <= 1
    mine = shards[3]
    batches = plan_batches([lengths[i] for i in mine], None, batch_size=8, window=64, max_batch_samples=8 * 60 * 16000)
    assert sorted(j for b in batches for j in b) == list(range(len(mine)))
    padded = sum(len(b) * lengths[mine[b[0]]] for b in batches)
    assert sum(lengths[i] for i in mine) / padded > 0.97        # length bucketing keeps the padding below 3 %


def test_check_info_severity_and_stage_words():
    """status words: one row per stage; failure codes raise the reference's exception types, the
    WPE 'singular' word only warns (nara_wpe falls back to lstsq silently); host lists or tensors."""
    import warnings
    from pb_chime5_b200 import ops
    ok = [[0, 0], [0, 0], [0, 0]]
    ops.check_info(ok, ('wpe', 'cacgmm', 'beamform'))
    with pytest.raises(ValueError, match='beamform.*utterance 1, frequency 7'):
        ops.check_info([[0, 0], [0, 0], [0, _lib.INFO_NOT_POSDEF | (7 << 8)]], ('wpe', 'cacgmm', 'beamform'))
    with pytest.raises(AssertionError, match='non-finite SNR'):
        ops.check_info([0, _lib.INFO_NONFINITE], 'beamform')
    with pytest.raises(RuntimeError, match='did not converge'):
        ops.check_info(torch.tensor([_lib.INFO_NO_CONVERGE | (3 << 8)]), 'cacgmm')
    # a singular WPE word in stage 0 does not hide the failure of stage 2, and alone it only warns
    with pytest.raises(ValueError):
        ops.check_info([[_lib.INFO_SINGULAR | (400 << 8)], [0], [_lib.INFO_NOT_POSDEF | (2 << 8)]], ('wpe', 'cacgmm', 'beamform'))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        ops.check_info([[_lib.INFO_SINGULAR | (5 << 8), 0], [0, 0]], ('wpe', 'cacgmm'))
    assert len(w) == 1 and 'singular normal equations' in str(w[0].message) and 'frequency 5' in str(w[0].message)
    ops.check_info(None, 'x')


def test_parse_beamformer_dsl_strings():
    """the string grammar of get_bf_vector (beamformer_wrapper.py:142-224) and its error behaviour"""
    from pb_chime5_b200.extraction import parse_beamformer
    assert parse_beamformer('mvdr_souden') == dict(rank1=None, core='mvdr_souden', ban=False, channel=0)
    assert parse_beamformer('rank1_gev+mvdr_souden+ban') == dict(rank1='rank1_gev', core='mvdr_souden', ban=True, channel=0)
    assert parse_beamformer('rank1_pca+wmwf')['rank1'] == 'rank1_pca'
    assert parse_beamformer('scaled_gev_atf+mvdr+ban') == dict(rank1=None, core='scaled_gev_atf+mvdr', ban=True, channel=0)
    assert parse_beamformer('ch13+ban') == dict(rank1=None, core='ch', ban=True, channel=13)
    for bad in ('music', 'rank1_gev+pca', 'mvdr', 'ban'):
        with pytest.raises(ValueError):
            parse_beamformer(bad)
    with pytest.raises(AssertionError):
        parse_beamformer('lcmv_souden')
    with pytest.raises(AssertionError):
        parse_beamformer(3)
    from pb_chime5_b200 import core
    assert core.Beamformer('wmwf+ban', None)._dsl_type() == (0x100 | 2 | (1 << 6), 0)
    assert core.Beamformer('mvdrSouden_ban', None)._dsl() is None
    with pytest.raises(NotImplementedError):
        core.Beamformer('nonsense', None)._dsl()


def test_launcher_environment_detection(monkeypatch):
    """rank / world / local rank from torchrun, mpiexec (Open MPI, MPICH) and srun variables"""
    for v in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'OMPI_COMM_WORLD_RANK', 'OMPI_COMM_WORLD_SIZE',
              'OMPI_COMM_WORLD_LOCAL_RANK', 'PMI_RANK', 'PMI_SIZE', 'MPI_LOCALRANKID', 'SLURM_PROCID',
              'SLURM_NTASKS', 'SLURM_LOCALID', 'PMIX_RANK', 'OMPI_UNIVERSE_SIZE'):
        monkeypatch.delenv(v, raising=False)
    assert sharding.rank_world() == (0, 1) and sharding.local_rank() == 0
    monkeypatch.setenv('OMPI_COMM_WORLD_RANK', '3'); monkeypatch.setenv('OMPI_COMM_WORLD_SIZE', '8')
    monkeypatch.setenv('OMPI_COMM_WORLD_LOCAL_RANK', '3')
    assert sharding.rank_world() == (3, 8) and sharding.local_rank() == 3
    monkeypatch.setenv('RANK', '1'); monkeypatch.setenv('WORLD_SIZE', '2'); monkeypatch.setenv('LOCAL_RANK', '1')
    assert sharding.rank_world() == (1, 2) and sharding.local_rank() == 1          # torchrun wins
    for v in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'OMPI_COMM_WORLD_RANK', 'OMPI_COMM_WORLD_SIZE', 'OMPI_COMM_WORLD_LOCAL_RANK'):
        monkeypatch.delenv(v)
    monkeypatch.setenv('SLURM_PROCID', '5'); monkeypatch.setenv('SLURM_NTASKS', '6'); monkeypatch.setenv('SLURM_LOCALID', '1')
    assert sharding.rank_world() == (5, 6) and sharding.local_rank() == 1
    q = sharding.WorkQueue(4)
    assert list(q) == [0, 1, 2, 3] and q.next() is None and q.taken == 4
