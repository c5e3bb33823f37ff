"""Summarise an .ncu-rep: headline metrics + barrier-delimited SASS segments."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print('==', d.get('Kernel Name', ('', ''))[0][:90])
    for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
              'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
              'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
              'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
              'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
              'sass__inst_executed_register_spilling', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
              'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_srcunit_tex_op_read.sum']:
        if k in d: print(f'  {k:75s} {d[k][0]:>16s} {d[k][1]}')
    st = {k: float(v[0]) for k, v in d.items() if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio')}
    print('  stalls/issue:', ', '.join(f"{k.split('stalled_')[1].split('_per')[0]}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
if len(sys.argv) > 2:
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; data = rows[2:]
    si = hdr.index('# Samples'); so = hdr.index('Source'); ie = hdr.index('Instructions Executed')
    s0 = hdr.index('stall_barrier'); names = hdr[s0:s0 + 17]
    data = [r for r in data if len(r) > si and not r[0].startswith('Kernel Name') and r[0] != 'Address']
    tot = sum(int(r[si] or 0) for r in data)
    seg_start = 0; segs = []
    for i, r in enumerate(data):
        if 'BAR.SYNC' in r[so] or 'BAR.RED' in r[so]:
            segs.append((seg_start, i)); seg_start = i + 1
    segs.append((seg_start, len(data) - 1))
    for a, b in segs:
        chunk = data[a:b + 1]
        s = sum(int(r[si] or 0) for r in chunk)
        if s < tot * 0.004: continue
        ninst = sum(int(r[ie] or 0) for r in chunk)
        ops = {}
        for r in chunk:
            parts = r[so].split()
            op = parts[1] if parts[0].startswith('@') else parts[0]
            op = op.split('.')[0]
            ops[op] = ops.get(op, 0) + int(r[ie] or 0)
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
        stl = {n: sum(int(r[s0 + j] or 0) for r in chunk) for j, n in enumerate(names)}
        tops = sorted(stl.items(), key=lambda kv: -kv[1])[:4]
        print(f'[{a:5d}-{b:5d}] samples {100*s/tot:5.1f}%  winst {ninst/1e9:6.2f}G  ops {[(k, round(v/1e9,2)) for k,v in top]}  stalls {[(k[6:], round(100*v/s)) for k,v in tops]}')
