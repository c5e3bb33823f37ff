"""Developer probe for the INT8 tensor-core WPE correlation build (csrc/wpe_gram_i8.cu).

Compares gss_debug_wpe_gram mode 1 (INT8, tcgen05) with mode 0 (float64 DMMA) and with a numpy
float64 Gram matrix on a few bins, for the descriptor variants, and times both modes.
    python tools/wpe_i8_probe.py [D taps T F]
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import _lib, ops  # noqa: E402


def make(B, F, D, T, seed=0, spread=1.0):
    rng = np.random.default_rng(seed)
    Y = rng.standard_normal((B, F, D, T)) + 1j * rng.standard_normal((B, F, D, T))
    env = np.exp(spread * np.cumsum(rng.standard_normal((B, F, 1, T)) * 0.3, axis=-1))
    Y = (Y * env).astype(np.complex64)
    lam = np.mean(np.abs(Y.astype(np.complex128)) ** 2, axis=2)
    lam = np.maximum(lam, 1e-10 * lam.max(axis=-1, keepdims=True))
    return Y, 1.0 / lam


def ref_gram(Y, inv, taps, delay):
    """float64 lower trapezoid (LD + D, LD) for one bin: rows [0, LD) R, rows [LD, LD + D) P^H."""
    D, T = Y.shape
    Yd = Y.astype(np.complex128)
    rows = []
    for k in range(taps):
        s = delay + k
        r = np.zeros((D, T), complex)
        r[:, s:] = Yd[:, :T - s]
        rows.append(r)
    A = np.concatenate(rows + [Yd], 0)
    return (A * inv) @ A[:taps * D].conj().T


def gram(Yt, invt, mode, variant, taps, delay):
    B, F, D, T = Yt.shape
    LD = taps * D
    out = torch.zeros((B, F, LD + D, LD), dtype=torch.complex128, device=Yt.device)
    n = _lib.workspace_bytes(_lib.OP_WPE, B, F, D, T, 0, taps)
    ws = ops.workspace(n, Yt.device)
    _lib.check(_lib.dev_lib().gss_debug_wpe_gram(ops._ptr(Yt), ops._ptr(invt), ops._ptr(out), mode, variant,
                                             B, F, D, T, taps, delay, None, ops._ptr(ws), ws.numel(), ops._stream()))
    torch.cuda.synchronize()
    return out


def main():
    D, taps, T, F = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (24, 10, 941, 40)
    delay, B = 2, 1
    dev = torch.device('cuda:0')
    Y, inv = make(B, F, D, T)
    Yt, invt = torch.from_numpy(Y).to(dev), torch.from_numpy(inv).to(dev)
    LD = taps * D
    tril = np.tril(np.ones((LD, LD), bool))
    refs = {f: ref_gram(Y[0, f], inv[0, f], taps, delay) for f in (0, F // 2, F - 1)}

    def err(out):
        worst = 0.0
        for f, R in refs.items():
            dg = np.sqrt(np.abs(np.diag(R[:LD]).real))
            dy = np.sqrt(np.abs(np.sum(np.abs(Y[0, f].astype(np.complex128)) ** 2 * inv[0, f], axis=-1)))
            o = out[0, f].cpu().numpy()
            e1 = np.abs(o[:LD] - R[:LD]) / np.outer(dg, dg)
            e2 = np.abs(o[LD:] - R[LD:]) / np.outer(dy, dg)
            worst = max(worst, e1[tril].max(), e2.max())
        return worst

    r0 = gram(Yt, invt, 0, 0, taps, delay)
    print(f'D={D} taps={taps} T={T} F={F}: float64 DMMA vs numpy  scaled err {err(r0):.3e}', flush=True)
    variants = (int(sys.argv[5]),) if len(sys.argv) >= 6 else (0,)
    for variant in variants:
        try:
            r1 = gram(Yt, invt, 1, variant, taps, delay)
            print(f'  INT8 variant {variant}: scaled err vs numpy {err(r1):.3e}', flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f'  INT8 variant {variant}: FAILED {ex!r}', flush=True)
            return
    for mode in (0, 1):
        for _ in range(2):
            gram(Yt, invt, mode, 0, taps, delay)
        t0 = time.perf_counter()
        for _ in range(5):
            gram(Yt, invt, mode, 0, taps, delay)
        dt = (time.perf_counter() - t0) / 5
        print(f'  mode {mode}: {dt * 1e3:.3f} ms for {F} bins -> {dt * 1e3 * 513 / F:.2f} ms per 513-bin utterance-iteration', flush=True)


if __name__ == '__main__':
    main()
