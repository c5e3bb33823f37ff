"""GPU: experimental INT8 tensor-core M-step covariance (csrc/mstep_i8.cu, gss_debug_mstep_i8)
against a numpy float64 evaluation of  Phi_k = sum_t w_kt y_t y_t^H  (cACG._fit,
complex_angular_central_gaussian.py:293-300) with EM-like weights (gamma / q: many decades)."""
import numpy as np
import pytest
import torch

from pb_chime5_b200 import _lib, ops

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_cuda(cuda):
    torch.cuda.set_device(cuda)


def run_mstep(Y, w, frames=None):
    dev = torch.device('cuda')
    B, F, D, T = Y.shape
    K = w.shape[2]
    Yt, wt = torch.from_numpy(Y).to(dev), torch.from_numpy(w).to(dev)
    out = torch.full((B, F, K, D, D), float('nan'), dtype=torch.complex128, device=dev)
    ws = ops.workspace(B * F * (-(-T // 32) * 320 * D + 4 * (D + K)) + 4096, dev)
    fr = None if frames is None else torch.tensor(frames, dtype=torch.int32, device=dev)
    _lib.check(_lib.dev_lib().gss_debug_mstep_i8(ops._ptr(Yt), ops._ptr(wt), ops._ptr(out), B, F, D, T, K, ops._ptr(fr),
                                             ops._ptr(ws), ws.numel(), ops._stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def em_like(B, F, D, T, K, seed):
    """unit-norm frames (normalize_observation, cACG.py:34-55) and weights gamma / q as the EM forms
    them: posteriors in [1e-10, 1], quadratic forms spread over ~3 decades"""
    rng = np.random.default_rng(seed)
    Y = (rng.standard_normal((B, F, D, T)) + 1j * rng.standard_normal((B, F, D, T))) * np.exp(rng.standard_normal((B, F, D, 1)))
    Y = (Y / np.linalg.norm(Y, axis=2, keepdims=True)).astype(np.complex64)
    gamma = rng.random((B, F, K, T)) ** 4
    gamma[:, :, :, ::7] = 1e-10                                 # clipped posteriors
    q = np.exp(1.2 * rng.standard_normal((B, F, K, T)))
    return Y, gamma / q


@pytest.mark.parametrize('D,K,T,F', [(24, 5, 941, 2), (24, 3, 100, 2), (8, 4, 333, 3), (4, 2, 31, 2), (16, 5, 600, 1)])
def test_mstep_i8_matches_float64(D, K, T, F):
    B = 2
    Y, w = em_like(B, F, D, T, K, seed=D + K)
    got = run_mstep(Y, w)
    Yd = Y.astype(np.complex128)
    ref = np.einsum('bfkt,bfdt,bfet->bfkde', w, Yd, Yd.conj())
    dg = np.sqrt(np.abs(np.einsum('bfkdd->bfkd', ref)))
    err = np.abs(got - ref) / (dg[..., :, None] * dg[..., None, :])
    norm = np.abs(got - ref).max(axis=(-1, -2)) / np.abs(ref).max(axis=(-1, -2))
    assert np.isfinite(got).all() and norm.max() < 1e-9 and err.max() < 2e-8, (norm.max(), err.max())
    # real diagonal by construction (ir - ri of identical integer sums)
    assert np.abs(np.einsum('bfkdd->bfkd', got).imag).max() == 0.0


def test_mstep_i8_ragged_and_zero_weights():
    Y, w = em_like(3, 2, 8, 200, 3, seed=1)
    w[0, :, 1] = 0.0                                            # a class without any weight
    frames = [200, 90, 0]
    got = run_mstep(Y, w, frames)
    Yd = Y.astype(np.complex128)
    for b, tv in enumerate(frames):
        ref = np.einsum('fkt,fdt,fet->fkde', w[b, :, :, :tv], Yd[b, :, :, :tv], Yd[b, :, :, :tv].conj())
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(got[b] - ref).max() <= 1e-9 * scale
    assert np.abs(got[0, :, 1]).max() == 0.0 and np.abs(got[2]).max() == 0.0


def test_mstep_i8_unsupported_shapes():
    Y, w = em_like(1, 1, 6, 50, 3, seed=2)
    with pytest.raises(NotImplementedError):
        run_mstep(Y, w)
