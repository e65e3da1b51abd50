#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, warp-stall mix and the hottest SASS lines (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg']
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')])
    for k in keys:
        if k in hdr:
            print(f"  {k} [{units[hdr.index(k)]}] = {r[hdr.index(k)]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print('total samples', tot, 'sass instrs', len(data), 'warp-instr executed', sum(int(r[ix['Instructions Executed']]) for r in data))
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f"  {s:26s} {v:7d} {100 * v / tot:5.1f}%")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top]:
    st = max(stalls, key=lambda s: int(r[ix[s]]))
    print(r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(8), st[6:].ljust(16), r[ix['Source']].strip()[:100])
