"""cfg1 shape (1 array x 4 mics, K=3, WPE off / on, 20 EM iterations): per-block timings."""
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from pb_chime5_b200 import ops, synth
dev = torch.device("cuda:0")
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r
for B in (1, 32):
    obs, act = synth.make_batch(100, B, D=4, T=941, F=513, K=3)
    Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)); A = torch.from_numpy(act).to(dev)
    ms, post = timeit(lambda: ops.cacgmm(Y, A, 20))
    print(f'cfg1 B={B}: cacgmm 20 it {ms:.3f} ms ({ms/B:.3f} ms/utt)')
    ms, X = timeit(lambda: ops.wpe(Y, 10, 2, 3))
    print(f'cfg1 B={B}: wpe {ms:.3f} ms ({ms/B:.3f} ms/utt)')
    ti = torch.zeros(B, dtype=torch.int32, device=dev)
    ms, Xh = timeit(lambda: ops.beamform_from_posterior(Y, post, ti, None, None))
    print(f'cfg1 B={B}: beamform {ms:.3f} ms ({ms/B:.3f} ms/utt)')
