"""Does the float32 -> float64 widening (F2F.F64.F32) cost FP64-pipe time?  Times the M-phase pattern
(8 widenings per 40 DFMA, 16 warps per SM) with the hardware conversion and with integer bit
manipulation, next to the plain 3-operand DFMA stream (developer library, csrc/probe.cu modes 3/6/7)."""
import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from pb_chime5_b200 import _lib
dl = _lib.dev_lib()
scratch = torch.empty(int(dl.gss_debug_fp64_peak_scratch_bytes()), dtype=torch.uint8, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, mode in (('dfma3 16 warps', 3), ('F2F + dfma', 6), ('bits + dfma', 7)):
    fl = ctypes.c_double(0.0)
    _lib.check(dl.gss_debug_fp64_peak(mode, 2000, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(fl), st), dl)
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(dl.gss_debug_fp64_peak(mode, 20000, ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(fl), st), dl)
        e1.record(); torch.cuda.synchronize()
        best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    print('%-16s %.2f TFLOP/s (DFMA flops only)' % (name, best))
