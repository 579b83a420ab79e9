#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics per launch + the top stall sites of launch 0.
usage: python profiles/ncu_summary.py <report.ncu-rep> [kernel-regex]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__cycles_elapsed.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print('kernels:', [r[ix['Kernel Name']][:60] for r in rows[2:]])
for k in KEYS:
    if k in ix:
        print(f'{k} [{units[ix[k]]}]:', [r[ix[k]] for r in rows[2:]])
for h in hdr:
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio'):
        vals = [float(r[ix[h]] or 0) for r in rows[2:]]
        if max(vals) > 0.15:
            print(h.replace('smsp__average_warps_issue_stalled_', 'stall/').replace('_per_issue_active.ratio', ''), [round(v, 2) for v in vals])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-count', '1'] + (['--kernel-name', 'regex:' + sys.argv[2]] if len(sys.argv) > 2 else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
seen, uniq = set(), []
for r in data:
    if r[ix['Address']] not in seen:
        seen.add(r[ix['Address']]); uniq.append(r)
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in uniq)
print('total samples', tot, 'warp instructions', sum(f(r, 'Instructions Executed') for r in uniq))
print('--- top stall sites (samples, executed, top stalls)')
for r in sorted(uniq, key=lambda r: -f(r, '# Samples'))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    st = {k: f(r, k) for k in hdr if k.startswith('stall_') and 'Not' not in k}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(r[ix['Source']].strip()[:64].ljust(64), int(f(r, '# Samples')), int(f(r, 'Instructions Executed')), top)
