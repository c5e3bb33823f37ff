"""Instruction histograms of the built objects (no GPU needed): proves which kernels carry
tcgen05 / TMA / FP64-MMA instructions.  Writes profiles/<tag>_sass_<object>.txt.

    python tools/sass_hist.py r2
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / 'pb_chime5_b200' / 'csrc' / 'build'
KEYS = ['UTCIMMA', 'UTCHMMA', 'UTCCP', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'SYNCS', 'DMMA', 'DFMA', 'DMUL',
        'DADD', 'F2F', 'LDS', 'STS', 'LDG', 'STG', 'LDGSTS', 'LDL', 'STL', 'BAR', 'MUFU', 'SHFL']


def main(tag):
    head = subprocess.run(['git', '-C', str(ROOT), 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    for obj in sorted(BUILD.glob('*.o')):
        sass = subprocess.run(['cuobjdump', '-sass', str(obj)], capture_output=True, text=True).stdout
        per_fn, fn = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.search(r'Function : (\S+)', line)
            if m:
                fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
                per_fn[fn] = collections.Counter()
                continue
            m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
            if m and fn is not None:
                per_fn[fn][m.group(1)] += 1
        out = [f'# {obj.name}: SASS mnemonic counts per kernel (static), git {head}, sm_100a', '']
        for fn, c in per_fn.items():
            tot = sum(c.values())
            if tot < 50:
                continue
            keys = ', '.join(f'{k} {c[k]}' for k in KEYS if c[k])
            out.append(f'{fn}\n    {tot} instructions: {keys}')
        (ROOT / 'profiles' / f'{tag}_sass_{obj.stem}.txt').write_text('\n'.join(out) + '\n')
        print(obj.name, len(per_fn), 'functions')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'r2')
