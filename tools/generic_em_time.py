"""Time the runtime-shape EM kernel (csrc/cacgmm_generic.cu) against the fused one."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from pb_chime5_b200 import ops, synth

dev = torch.device('cuda:0')


def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for D, K in ((24, 5), (24, 8), (24, 12), (32, 5)):
    obs, act = synth.make_batch(1000, 1, D=D, T=941, F=513, K=K)
    Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)); A = torch.from_numpy(act).to(dev)
    ms = timeit(lambda: ops.cacgmm(Y, A, 20))
    line = f'D={D} K={K}: 20 EM iterations {ms:.1f} ms per utterance (complex64 input)'
    if (D, K) == (24, 5):
        ms64 = timeit(lambda: ops.cacgmm(Y.to(torch.complex128), A, 20))
        line += f'; runtime-shape kernel on complex128 input {ms64:.1f} ms'
    print(line)
