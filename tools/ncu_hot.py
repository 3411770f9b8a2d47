#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall-reason totals and the hottest SASS instructions of the first profiled launch.
usage: python tools/ncu_hot.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
        'sm__inst_executed_pipe_tensor', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_pipe_xu.sum',
        'sm__warps_active.avg.per_cycle_active', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_uniform.sum']
for a, b, c in zip(h, v, u):
    if any(a.endswith(w) for w in want): print(f'{a} = {b} {c}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
hi = heads[0]; end = heads[1] - 1 if len(heads) > 1 else len(rows)
h = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
isrc = h.index('Source'); isamp = h.index('# Samples'); iinst = h.index('Instructions Executed')
stalls = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
def I(x):
    try: return int(x)
    except Exception: return 0
tot = sum(I(r[isamp]) for r in data)
print('total samples', tot, 'SASS instructions', len(data))
agg = {}
for r in data:
    for i in stalls: agg[h[i]] = agg.get(h[i], 0) + I(r[i])
for k, val in sorted(agg.items(), key=lambda x: -x[1])[:10]: print(f'  {k:26s}{val:8d} {100 * val / max(tot, 1):5.1f}%')
mn = {}
for r in data:
    op = r[isrc].split()[0] if r[isrc].split() else ''
    if op.startswith('@'): op = r[isrc].split()[1]
    mn[op.split('.')[0]] = mn.get(op.split('.')[0], 0) + I(r[iinst])
print('executed warp instructions by mnemonic:', ', '.join(f'{k} {v}' for k, v in sorted(mn.items(), key=lambda x: -x[1])[:18]))
for r in sorted(data, key=lambda r: -I(r[isamp]))[:topn]:
    st = sorted([(I(r[i]), h[i]) for i in stalls], reverse=True)[:2]
    print(r[isamp].rjust(6), r[iinst].rjust(8), r[isrc][:70].ljust(70), st)
