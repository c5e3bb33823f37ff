import torch
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import ops, synth
dev = torch.device("cuda:0")
B = 4
obs, act = synth.make_batch(1000, B, D=24, T=941, F=513, K=5)
Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)); A = torch.from_numpy(act).to(dev)
ops.cacgmm(Y, A, 10); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); p = ops.cacgmm(Y, A, 100); e1.record(); torch.cuda.synchronize()
print("EM ms/utt %.2f  checksum %.6f" % (e0.elapsed_time(e1) / B, float(p.double().sum())))
