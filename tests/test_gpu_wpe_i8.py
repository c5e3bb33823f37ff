"""GPU: the INT8 tensor-core WPE correlation build (csrc/wpe_gram_i8.cu, tcgen05 kind::i8,
exact digit-split integer arithmetic) against a numpy float64 Gram matrix, against the float64
(DMMA) build, and end to end through gss_wpe_c64 against the oracle -- including the float64
re-do of ill-conditioned bins."""
import numpy as np
import pytest
import torch

from oracle import gss_oracle as oracle
from pb_chime5_b200 import _lib, ops, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_cuda(cuda):
    torch.cuda.set_device(cuda)


def new_stats():
    return torch.zeros(4, dtype=torch.int32, device='cuda')


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make_input(B, F, D, T, seed, spread=0.3):
    """complex64 frames with a random-walk power envelope (speech-like dynamics) + the WPE weights"""
    rng = np.random.default_rng(seed)
    Y = rng.standard_normal((B, F, D, T)) + 1j * rng.standard_normal((B, F, D, T))
    env = np.exp(np.cumsum(rng.standard_normal((B, F, 1, T)) * spread, axis=-1))
    gains = np.exp(rng.standard_normal((1, 1, D, 1)))            # unequal channel gains
    Y = (Y * env * gains).astype(np.complex64)
    lam = np.mean(np.abs(Y.astype(np.complex128)) ** 2, axis=2)
    lam = np.maximum(lam, 1e-10 * lam.max(axis=-1, keepdims=True))
    return Y, 1.0 / lam


def ref_gram(Y, inv, taps, delay, Tv=None):
    """float64 (LD + D, LD): rows [0, LD) = R = Yt L^-1 Yt^H, rows [LD, LD + D) = P^H (wpe.cu contract)"""
    D, T = Y.shape
    Tv = T if Tv is None else Tv
    Yd = Y.astype(np.complex128).copy()
    Yd[:, Tv:] = 0
    inv = inv.copy()
    inv[Tv:] = 0
    rows = []
    for k in range(taps):
        s = delay + k
        r = np.zeros((D, T), complex)
        r[:, s:] = Yd[:, :T - s]
        rows.append(r)
    A = np.concatenate(rows + [Yd], 0)
    return (A * inv) @ A[:taps * D].conj().T


def run_gram(Yt, invt, mode, taps, delay, frames=None):
    B, F, D, T = Yt.shape
    LD = taps * D
    out = torch.full((B, F, LD + D, LD), float('nan'), dtype=torch.complex128, device=Yt.device)
    ws = ops.workspace(_lib.workspace_bytes(_lib.OP_WPE, B, F, D, T, 0, taps), Yt.device)
    fr = None if frames is None else torch.tensor(frames, dtype=torch.int32, device=Yt.device)
    dev_lib = _lib.dev_lib()                                  # developer API: libgss_dev.so (include/gss_dev.h)
    _lib.check(dev_lib.gss_debug_wpe_gram(ops._ptr(Yt), ops._ptr(invt), ops._ptr(out), mode, 0, B, F, D, T,
                                          taps, delay, ops._ptr(fr), ops._ptr(ws), ws.numel(), ops._stream()), dev_lib)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def scaled_err(got, ref, LD, Y, inv):
    """max |got - ref| / sqrt(R_ii R_jj) over the lower trapezoid (P^H rows: sqrt(E_d R_jj),
    E_d = sum_t inv |y_d|^2, the diagonal the unshifted rows would have)"""
    dg = np.sqrt(np.abs(np.diag(ref[:LD]).real))
    ed = np.sqrt(np.sum(np.abs(Y.astype(np.complex128)) ** 2 * inv, axis=-1))
    tril = np.tril(np.ones((LD, LD), bool))
    e1 = (np.abs(got[:LD] - ref[:LD]) / np.outer(dg, dg))[tril].max()
    e2 = (np.abs(got[LD:] - ref[LD:]) / np.outer(ed, dg)).max()
    return float(max(e1, e2))


@pytest.mark.parametrize('D,taps,T,F', [(24, 10, 941, 3), (8, 10, 191, 4), (24, 2, 100, 2), (6, 10, 64, 2),
                                        (5, 12, 333, 2), (24, 20, 700, 1), (8, 6, 17, 2), (12, 4, 33, 1)])
def test_gram_i8_matches_float64(D, taps, T, F):
    dev = torch.device('cuda')
    delay, B = 2, 2
    Y, inv = make_input(B, F, D, T, seed=D * 100 + taps)
    Yt, invt = torch.from_numpy(Y).to(dev), torch.from_numpy(inv).to(dev)
    LD = taps * D
    g8 = run_gram(Yt, invt, 1, taps, delay)
    g64 = run_gram(Yt, invt, 0, taps, delay)
    tril = np.tril(np.ones((LD, LD), bool))
    for b in range(B):
        for f in range(F):
            ref = ref_gram(Y[b, f], inv[b, f], taps, delay)
            assert np.isfinite(g8[b, f][:LD][tril]).all() and np.isfinite(g8[b, f][LD:]).all()
            assert scaled_err(g64[b, f], ref, LD, Y[b, f], inv[b, f]) < 1e-13
            assert scaled_err(g8[b, f], ref, LD, Y[b, f], inv[b, f]) < 2e-9
            # exact Hermitian structure of the integer path: real diagonal
            assert np.abs(np.diag(g8[b, f][:LD]).imag).max() == 0.0


def test_gram_i8_ragged_and_silent():
    """per-utterance frame counts; an all-zero channel; a bin without valid frames"""
    dev = torch.device('cuda')
    D, taps, T, F, delay = 8, 8, 200, 2, 3
    Y, inv = make_input(3, F, D, T, seed=5)
    Y[0, :, 2] = 0                                            # dead channel
    frames = [200, 77, 0]
    Yt, invt = torch.from_numpy(Y).to(dev), torch.from_numpy(inv).to(dev)
    LD = taps * D
    g8 = run_gram(Yt, invt, 1, taps, delay, frames)
    tril = np.tril(np.ones((LD, LD), bool))
    for b, tv in enumerate(frames):
        for f in range(F):
            ref = ref_gram(Y[b, f], inv[b, f], taps, delay, tv)
            scale = max(np.abs(ref).max(), 1e-300)
            assert np.abs(g8[b, f][:LD] - ref[:LD])[tril].max() <= 2e-9 * scale
            assert np.abs(g8[b, f][LD:] - ref[LD:]).max() <= 2e-9 * scale
    low = np.concatenate([tril, np.ones((D, LD), bool)], 0)   # the trapezoid the contract defines
    assert np.abs(g8[2][:, low]).max() == 0.0
    dead = np.zeros((LD + D, LD), bool)
    dead[:, [k * D + 2 for k in range(taps)]] = True
    dead[[k * D + 2 for k in range(taps)] + [LD + 2], :] = True
    assert np.abs(g8[0][:, dead & low]).max() == 0.0


def test_wpe_i8_equals_float64_path_and_oracle():
    """well-conditioned benchmark-like input: no bin is re-done, outputs agree with the float64
    build to rounding and with the oracle to the parity bar"""
    dev = torch.device('cuda')
    Obs, _ = synth.make_utterance(11, D=24, T=941, F=3, K=5)      # T >> taps * D: well conditioned
    Obs[:, 3:, :] += 0.4 * Obs[:, :-3, :]
    Y = ops.pack_dtf_to_fdt(torch.from_numpy(Obs).to(dev)[None])
    x64 = ops.wpe(Y, 10, 2, 3, gram_mode='f64').cpu().numpy()
    st = new_stats()
    x8 = ops.wpe(Y, 10, 2, 3, stats=st).cpu().numpy()          # default: INT8 + float64 re-do
    assert st.tolist()[:3] == [3, 0, 0], st.tolist()           # 3 bins, none flagged, no re-do build
    assert rel_err(x8, x64) < 2e-7, rel_err(x8, x64)       # complex64 outputs: identical up to the last bit
    ref = oracle.wpe_dtf(Obs.astype(np.complex128), 10, 2, 3)
    got = ops.unpack_fdt_to_dtf(torch.from_numpy(x8).to(dev))[0].cpu().numpy()
    assert rel_err(got, ref) < 1e-5


def test_wpe_i8_illconditioned_bins_are_redone_in_float64():
    """reverberant, low-noise audio (cond 1e6+): the a-posteriori pivot test flags the bins and the
    result is the float64 build's, bit for bit; with the test disabled the INT8-only result differs"""
    dev = torch.device('cuda')
    obs, _ = synth.make_reverberant_audio(3, D=8, N=32000, K=3)
    Y = ops.stft(torch.from_numpy(obs).to(dev)[None])
    Ysub = Y[:, [5, 40, 129, 300]].contiguous()
    x64 = ops.wpe(Ysub, 10, 2, 3, gram_mode='f64').cpu().numpy()
    st = new_stats()
    x8 = ops.wpe(Ysub, 10, 2, 3, gram_mode='i8+redo', stats=st).cpu().numpy()
    bins, on_list, redo_builds = st.tolist()[:3]
    assert bins == 4 and on_list > 0
    # a bin flagged in iteration i is re-done in float64 in iterations i, i+1, ... (the flag is carried)
    assert on_list <= redo_builds <= 3 * on_list
    if redo_builds == 3 * 4:                                  # every bin flagged in the first iteration
        assert np.array_equal(x8, x64)
    else:
        assert rel_err(x8, x64) < 1e-5
    # threshold 0: nothing is re-done, the INT8 Gram matrix alone still dereverberates sensibly
    st = new_stats()
    xi = ops.wpe(Ysub, 10, 2, 1, gram_mode='i8+redo', i8_tau=0.0, stats=st).cpu().numpy()
    assert st.tolist()[1:3] == [0, 0]
    x1 = ops.wpe(Ysub, 10, 2, 1, gram_mode='f64').cpu().numpy()
    assert np.isfinite(xi).all() and rel_err(xi, x1) < 1e-3


def test_wpe_flags_are_carried_through_the_iterations():
    """a mixed batch: well-conditioned bins stay on the INT8 path, ill-conditioned ones move to the
    float64 list in the iteration that flags them and stay there; every bin of the result equals
    either path's result for that bin (float64 bit for bit where the bin was listed from the start)"""
    dev = torch.device('cuda')
    obs, _ = synth.make_reverberant_audio(3, D=8, N=96000, K=3)
    Yr = ops.stft(torch.from_numpy(obs).to(dev)[None])[:, [5, 40]].contiguous()          # ill conditioned
    T = Yr.shape[3]
    Ow, _ = synth.make_utterance(12, D=8, T=T, F=2, K=3)                                  # well conditioned
    Yw = ops.pack_dtf_to_fdt(torch.from_numpy(Ow).to(dev)[None])
    Y = torch.cat([Yr, Yw], dim=1).contiguous()                                          # 4 bins
    st = new_stats()
    x = ops.wpe(Y, 10, 2, 3, stats=st)
    x64 = ops.wpe(Y, 10, 2, 3, gram_mode='f64')
    xi8 = ops.wpe(Y, 10, 2, 3, gram_mode='i8')
    bins, on_list, redo_builds = st.tolist()[:3]
    assert bins == 4 and 1 <= on_list <= 2 and redo_builds <= 3 * on_list
    assert torch.equal(x[:, 2:], xi8[:, 2:])                  # never flagged: the INT8 path's result
    assert rel_err(x[:, :2].cpu().numpy(), x64[:, :2].cpu().numpy()) < 1e-5
    assert rel_err(x.cpu().numpy(), x64.cpu().numpy()) < 1e-5


def test_cfg3_like_ragged_batch_through_int8_path(monkeypatch):
    """BASELINE configs[2] scaled down: 24 channels, dev-shaped (ragged) utterance lengths, full
    WPE (INT8 correlation build) + GSS + GEV+BAN: every utterance of the padded batch equals its
    single-utterance run bit for bit, and the single runs meet the parity bar against the oracle.
    The correlation-build policy is pinned: the adaptive default of the WPE block may move later
    batches to the float64 build (results then differ in the last bits of complex64)."""
    from pb_chime5_b200 import core
    monkeypatch.setenv('GSS_WPE_GRAM', 'i8+redo')
    dev = torch.device('cuda')
    lens = [941, 520, 333]
    Tmax, D, F, K = 941, 24, 3, 5
    enh = core.get_enhancer(wpe_tabs=10, wpe_iterations=3, bss_iterations=10, bf='gev_ban')
    Ypad = torch.zeros((len(lens), F, D, Tmax), dtype=torch.complex64, device=dev)
    Apad = torch.zeros((len(lens), K, Tmax), dtype=torch.uint8, device=dev)
    singles, inputs = [], []
    iv = lambda v: torch.tensor([v], dtype=torch.int32, device=dev)   # noqa: E731
    for b, T in enumerate(lens):
        obs, act = synth.make_utterance(900 + b, D=D, T=T, F=F, K=K)
        Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)[None])
        A = torch.from_numpy(act.astype(np.uint8))[None].to(dev)
        Ypad[b, :, :, :T] = Y[0]
        Ypad[b, :, :, T:] = 3.0
        Apad[b, :, :T] = A[0]
        singles.append(enh.enhance_stft_batch(Y, A, iv(0), iv(3), iv(3), return_masks=True))
        inputs.append((obs, act))
    ti = torch.zeros(len(lens), dtype=torch.int32, device=dev)
    c3 = torch.full((len(lens),), 3, dtype=torch.int32, device=dev)
    X, post = enh.enhance_stft_batch(Ypad, Apad, ti, c3, c3, return_masks=True, frames=lens)
    for b, T in enumerate(lens):
        Xs, ps = singles[b]
        assert torch.equal(post[b, :, :, :T], ps[0]) and torch.equal(X[b, :, :T], Xs[0])
        if T < Tmax:
            assert float(post[b, :, :, T:].abs().max()) == 0 and float(X[b, :, T:].abs().max()) == 0
    # oracle parity of the longest utterance (WPE + guided EM masks; GEV output by magnitude)
    obs, act = inputs[0]
    ref = oracle.enhance_stft(obs.astype(np.complex128), act, 0,
                              wpe=dict(taps=10, delay=2, iterations=3, psd_context=0), gss_iterations=10,
                              bf='gev_ban', start_context_frames=3, end_context_frames=3)
    m = ops.unpack_fkt_to_ktf(singles[0][1])[0].cpu().numpy()
    m[:, :3] = 0
    m[:, -3:] = 0
    assert np.abs(m - ref['masks']).max() < 1e-4
    Xd = ops.unpack_ft_to_tf(singles[0][0])[0].cpu().numpy()
    assert rel_err(np.abs(Xd), np.abs(ref['X_hat'])) < 1e-4


def test_wpe_block_moves_to_float64_on_reverberant_data():
    """core.WPE: when more than half of the bins of a call end on the float64 list, the following
    calls skip the INT8 attempt (float64 build for every bin) and equal the pinned float64 result bit
    for bit; on well-conditioned data the block stays on the INT8 path."""
    from pb_chime5_b200 import core
    dev = torch.device('cuda')
    obs, _ = synth.make_reverberant_audio(3, D=8, N=32000, K=3)
    Y = ops.stft(torch.from_numpy(obs).to(dev)[None])[:, [5, 40, 129, 300]].contiguous()
    blk = core.WPE(taps=10, delay=2, iterations=3, psd_context=0)
    x64 = ops.wpe(Y, 10, 2, 3, gram_mode='f64')
    first = blk._run(Y)
    torch.cuda.synchronize()                       # statistics of the first call have landed
    second = blk._run(Y)
    assert blk.last_float64_fraction is not None and blk.last_float64_fraction > 0.5
    assert blk.__dict__['_gram_state']['mode'] == 'f64'
    assert torch.equal(second, x64)
    assert rel_err(first.cpu().numpy(), x64.cpu().numpy()) < 1e-5
    # well conditioned: stays on the tensor-core path
    Ow, _ = synth.make_utterance(12, D=8, T=600, F=2, K=3)
    Yw = ops.pack_dtf_to_fdt(torch.from_numpy(Ow).to(dev)[None])
    blk2 = core.WPE(taps=10, delay=2, iterations=3, psd_context=0)
    a = blk2._run(Yw)
    torch.cuda.synchronize()
    b = blk2._run(Yw)
    assert blk2.last_float64_fraction == 0.0 and blk2.__dict__['_gram_state']['mode'] is None
    assert torch.equal(a, b) and torch.equal(a, ops.wpe(Yw, 10, 2, 3, gram_mode='i8'))
