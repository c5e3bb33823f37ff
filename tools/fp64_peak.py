"""Print the FP64 peaks of this GPU through the developer library (csrc/probe.cu)."""
import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from pb_chime5_b200 import _lib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import bench
print(bench.measure_fp64_peak(torch, _lib))
