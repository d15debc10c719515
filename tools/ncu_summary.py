#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics, stall mix and executed-opcode histogram.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [rows_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
rows_per_launch = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
print('kernel:', r[hdr.index('Kernel Name')][:90])
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
for w in want:
    if w in hdr:
        print(f'  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}')
print('stalls per issue:')
for i, h in enumerate(hdr):
    if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try:
            v = float(r[i])
        except ValueError:
            continue
        if v > 0.05:
            print(f'  {h.split("issue_stalled_")[1].split("_per_issue")[0]:22s} {v:.2f}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, samp, tot, n = collections.Counter(), collections.Counter(), 0, 0
for rr in rows[2:]:
    if rr and rr[0] == 'Kernel Name':
        break
    if len(rr) < len(hdr):
        continue
    try:
        e, s = int(rr[iE]), int(rr[iSm])
    except ValueError:
        continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', rr[iS].strip())
    op = m.group(2) if m else rr[iS][:10]
    base = op.split('.')[0]
    if base == 'MUFU':
        base = op
    elif base in ('LDS', 'STS', 'LDL', 'STL', 'LDG', 'STG'):
        base += '.128' if '.128' in op else ('.64' if '.64' in op else '')
    ops[base] += e
    samp[base] += s
    tot += e
    n += 1
print(f'static instructions {n}, executed warp-instructions {tot}' +
      (f', per row-thread {tot * 32 / rows_per_launch:.0f}' if rows_per_launch else ''))
ts = sum(samp.values()) or 1
for k, v in ops.most_common(22):
    print(f'  {k:12s} {100 * v / tot:5.1f}% of instr   {100 * samp[k] / ts:5.1f}% of stall samples' +
          (f'   {v * 32 / rows_per_launch:7.1f}/row' if rows_per_launch else ''))
