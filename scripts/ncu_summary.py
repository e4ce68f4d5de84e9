#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics per launch + warp-stall samples by SASS instruction.
usage: ncu_summary.py <file.ncu-rep> [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('---- launch')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:72s} {r[i][:70]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if not hi:
    sys.exit(0)
hdr = rows[hi[0]]
si, so, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
end = hi[1] - 1 if len(hi) > 1 else len(rows)
data = [r for r in rows[hi[0] + 1:end] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data) or 1
print(f"---- first launch: {len(data)} SASS instructions, {tot} samples, {sum(int(r[ie]) for r in data)} warp-instructions executed")
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stalls}
print("  stall reasons: " + ", ".join(f"{k[6:]} {v / tot * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for n, r in sorted(enumerate(data), key=lambda x: -int(x[1][si]))[:topn]:
    top = max(stalls, key=lambda i: int(r[i] or 0))
    print(f"  #{n:4d} {int(r[si]) / tot * 100:5.1f}%  exec {r[ie]:>8s}  {hdr[top][6:]:16s} {r[so].strip()[:100]}")
