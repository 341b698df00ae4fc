#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw + source pages) without a GPU: per-kernel key metrics, opcode mix
per kernel and the most-sampled instructions.   usage: ncu_summary.py REPORT [units_per_launch]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit_row = rows[0], rows[1]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_average_branch_targets_threads_uniform.pct', 'lts__t_bytes.sum', 'sm__cycles_elapsed.avg']
for vals in rows[2:]:
    print("==", vals[hdr.index('Kernel Name')][:70])
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"   {h:75s} {unit_row[i]:12s} {vals[i]}")
    for i, h in enumerate(hdr):
        if 'stalled' in h and h.endswith('per_warp_active.pct') and float(vals[i] or 0) > 4:
            print(f"   {h:75s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = set()
for a, b in zip(starts[:-1], starts[1:]):
    name = rows[a][1][:60]
    if name in seen:
        continue
    seen.add(name)
    h = rows[a + 1]
    isrc, iinst, ismp, ithr = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
    data = [r for r in rows[a + 2:b] if len(r) > iinst and r[iinst].isdigit()]
    tot = sum(int(r[iinst]) for r in data)
    print(f"\n== source: {name}: {tot/1e6:.1f}M warp instructions, {len(data)} SASS")
    ops = collections.Counter()
    for r in data:
        t = r[isrc].split()
        op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '')).split('.')[0]
        ops[op] += int(r[iinst])
    print("   opcode mix:", ", ".join(f"{k} {v/1e6:.1f}M" + (f" ({v/units:.1f}/u)" if units else "") for k, v in ops.most_common(14)))
    for r in sorted(data, key=lambda r: -int(r[ismp] or 0))[:14]:
        print(f"   smp={r[ismp]:>6} exec={int(r[iinst])/1e6:7.2f}M thr={r[ithr]:>4} {r[isrc][:80]}")
