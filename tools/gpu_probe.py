"""Ad-hoc GPU probe: timing of the main kernels at config-2 shape."""
import time
import numpy as np
import torch
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import ops, synth

dev = torch.device('cuda:0')
print(torch.cuda.get_device_name(0))
B = 2
obs, act = synth.make_batch(1000, B, D=24, T=941, F=513, K=5)
Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev))
A = torch.from_numpy(act).to(dev)

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

for it in (20, 100):
    ms, post = timeit(lambda: ops.cacgmm(Y, A, it), 2)
    print(f'cacgmm B={B} it={it}: {ms:.2f} ms  ({ms/B:.2f} ms/utt)')
ms, X = timeit(lambda: ops.wpe(Y, 10, 2, 3), 2)
print(f'wpe B={B}: {ms:.2f} ms ({ms/B:.2f} ms/utt)')
ti = torch.zeros(B, dtype=torch.int32, device=dev)
ms, Xh = timeit(lambda: ops.beamform_from_posterior(Y, post, ti, None, None), 3)
print(f'beamform B={B}: {ms:.2f} ms')
ms, Xh = timeit(lambda: ops.beamform_from_posterior(Y, post, ti, None, None, bf='gev_ban'), 3)
print(f'beamform gev B={B}: {ms:.2f} ms')
print('post sum', float(post.sum()), 'finite', bool(torch.isfinite(post).all()), bool(torch.isfinite(torch.view_as_real(X)).all()))

# cfg5 stress shape, one utterance: 60 s, T=3753, K=6, WPE taps=20, 200 EM iterations
del Y, post, X, Xh
torch.cuda.empty_cache()
obs, act = synth.make_batch(5000, 1, D=24, T=3753, F=513, K=6)
Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)); A = torch.from_numpy(act).to(dev)
ms, X = timeit(lambda: ops.wpe(Y, 20, 2, 3), 1)
print(f'cfg5 wpe taps=20: {ms:.1f} ms/utt')
ms, post = timeit(lambda: ops.cacgmm(X, A, 200), 1)
print(f'cfg5 cacgmm 200 it: {ms:.1f} ms/utt, finite {bool(torch.isfinite(post).all())}')
