"""Achieved HBM bandwidth of the standalone weighted-covariance kernel (gss_weighted_cov_c64)."""
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import json
import torch
from pb_chime5_b200 import ops
dev = torch.device('cuda:0')
peak = 6538.3
for (B, D, K) in ((64, 4, 3), (64, 4, 2), (16, 8, 4), (8, 24, 5), (8, 24, 2)):
    F, T = 513, 941
    Y = torch.randn(B, F, D, T, dtype=torch.complex64, device=dev)
    w = torch.rand(B, F, K, T, device=dev)
    for _ in range(3):
        P = ops.weighted_cov(Y, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        P = ops.weighted_cov(Y, w)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nbytes = B * F * (D * T * 8 + K * T * 4 + K * D * D * 8)
    flops = B * F * K * T * 8 * D * D
    print(json.dumps(dict(B=B, D=D, K=K, ms=round(ms, 3), GBs=round(nbytes / ms / 1e6, 1), frac_hbm=round(nbytes / ms / 1e6 / peak, 3),
                          TFLOPs=round(flops / ms / 1e9, 2))))
