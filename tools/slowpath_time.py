"""EM timing when one class is rank deficient (exact Jacobi path every pass for that class)."""
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from pb_chime5_b200 import ops, synth
dev = torch.device("cuda:0")
B = 2
obs, act = synth.make_batch(1000, B, D=24, T=941, F=513, K=5)
act[:, 1] = False
act[:, 1, 100:110] = True          # 10 active frames < D = 24
Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)); A = torch.from_numpy(act).to(dev)
ops.cacgmm(Y, A, 5); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); p = ops.cacgmm(Y, A, 100); e1.record(); torch.cuda.synchronize()
print("EM with a rank-deficient class: ms/utt %.2f finite %s" % (e0.elapsed_time(e1) / B, bool(torch.isfinite(p).all())))
