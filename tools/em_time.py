import torch
import sys, pathlib; sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from pb_chime5_b200 import ops, synth
dev = torch.device("cuda:0")
B = 4
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5          # em_time.py [K [T]]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 941
obs, act = synth.make_batch(1000, B, D=24, T=T, F=513, K=K)
Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev)); A = torch.from_numpy(act).to(dev)
ops.cacgmm(Y, A, 10); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); p = ops.cacgmm(Y, A, 100); e1.record(); torch.cuda.synchronize()
print("K=%d T=%d " % (K, T) + "EM ms/utt %.2f  checksum %.6f" % (e0.elapsed_time(e1) / B, float(p.double().sum())))
