"""Run single kernels at cfg2 shape for ncu (B=1)."""
import sys
import torch
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import ops, synth
which = sys.argv[1] if len(sys.argv) > 1 else 'em'
dev = torch.device('cuda:0')
obs, act = synth.make_batch(1000, 1, D=24, T=941, F=513, K=5)
Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev))
A = torch.from_numpy(act).to(dev)
if which == 'em':
    for _ in range(2):
        post = ops.cacgmm(Y, A, 20)
elif which == 'wpe':
    for _ in range(2):
        X = ops.wpe(Y, 10, 2, 1)
elif which == 'wpe_f64':                                      # float64 (DMMA, three-product) correlation build
    for _ in range(2):
        X = ops.wpe(Y, 10, 2, 1, gram_mode='f64')
torch.cuda.synchronize()
print('done')
