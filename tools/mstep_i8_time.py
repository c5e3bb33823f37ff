"""Time the experimental INT8 M-step covariance (gss_debug_mstep_i8) against the float64 CUDA-core
kernel behind gss_weighted_cov_c64 at the cfg2 shape.  Developer tool."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import _lib, ops  # noqa: E402

B, F, D, T, K = 4, 513, 24, 941, 5
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
Y = torch.randn((B, F, D, T, 2), device=dev, generator=g).view(B, F, D, T, 2)
Y = torch.view_as_complex(Y.float().contiguous())
w = torch.rand((B, F, K, T), device=dev, generator=g, dtype=torch.float64)
out = torch.empty((B, F, K, D, D), dtype=torch.complex128, device=dev)
ws = ops.workspace(B * F * (-(-T // 32) * 320 * D + 4 * (D + K)) + 4096, dev)


def i8():
    _lib.check(_lib.dev_lib().gss_debug_mstep_i8(ops._ptr(Y), ops._ptr(w), ops._ptr(out), B, F, D, T, K, None,
                                             ops._ptr(ws), ws.numel(), ops._stream()))


wf = w.float()


def f64():
    return ops.weighted_cov(Y, wf)


for name, fn in (('INT8 tensor-core M-step (scale + Y planes + GEMM)', i8), ('float64 CUDA-core weighted_cov', f64)):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f'{name}: {e0.elapsed_time(e1) / 5 / B:.3f} ms per utterance and call (B={B})')
