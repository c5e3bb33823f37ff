"""WPE with the float64 (DMMA) correlation build: ms per utterance at the cfg2 shape (T=941) and at a
cfg3-like segment length (T=2500), plus the reverberant side line of bench.py."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import bench
from pb_chime5_b200 import core, ops, synth
dev = torch.device('cuda:0')
for T in (941, 2500):
    obs, _ = synth.make_batch(3, 2, D=24, T=T, F=513, K=5)
    Y = ops.pack_dtf_to_fdt(torch.from_numpy(obs).to(dev))
    for mode in ('f64', 'i8+redo'):
        ops.wpe(Y, 10, 2, 3, gram_mode=mode); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3): ops.wpe(Y, 10, 2, 3, gram_mode=mode)
        e1.record(); torch.cuda.synchronize()
        print('T=%d %-8s %.2f ms per utterance (3 iterations)' % (T, mode, e0.elapsed_time(e1) / 3 / 2))
c = dict(D=24, taps=10, delay=2, wpe_iterations=3)
r = bench.measure_wpe_reverberant(torch, core, ops, synth, c)
print(json.dumps({k: r[k] for k in ('ms_per_utterance_f64', 'ms_per_utterance_i8_redo')}))
