"""DRAM traffic of the fused EM kernel from an ncu --set full report -> profiles/em_kernel_traffic.json
(the `roofline.traffic` source of bench.py), stamped with the git hash and the launch it came from.

    python tools/em_traffic.py gpurun_out/<report>.ncu-rep <utterances in the profiled launch> <EM iterations>
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, B, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
unit_scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
best = None
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    if 'cacgmm_em_kernel' not in d.get('Kernel Name', ''):
        continue
    rd = float(d['dram__bytes_read.sum']) * unit_scale[u['dram__bytes_read.sum']]
    wr = float(d['dram__bytes_write.sum']) * unit_scale[u['dram__bytes_write.sum']]
    best = dict(kernel=d['Kernel Name'], dram_bytes_read=rd, dram_bytes_write=wr,
                gpu_time_ms=float(d['gpu__time_duration.sum']) * {'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(u['gpu__time_duration.sum'], 1))
assert best, 'no cacgmm_em_kernel launch in the report'
head = subprocess.run(['git', '-C', str(ROOT), 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
best.update(utterances_in_launch=B, em_iterations=iters, git=head, report=Path(rep).name,
            dram_bytes_per_utterance=(best['dram_bytes_read'] + best['dram_bytes_write']) / B,
            note='physical DRAM bytes of ONE launch of the fused EM kernel (cfg2 shape); independent of the iteration '
                 'count as long as Y[f] stays L2 resident between the iterations')
(ROOT / 'profiles' / 'em_kernel_traffic.json').write_text(json.dumps(best, indent=1) + '\n')
print(json.dumps(best, indent=1))
