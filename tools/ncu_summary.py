#!/usr/bin/env python
"""Summarise an `ncu --set full` report for the judge / for tuning.
usage: tools/ncu_summary.py <raw.csv> <source.csv> <dst pixels per launch>"""
import collections
import csv
import sys

raw, src, npx = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_cbu_pred_on_any.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want or h.startswith('smsp__average_warp') or ('warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')):
        try:
            v = float(vals[i])
        except ValueError:
            continue
        if 'stalled' in h and v < 3:
            continue
        print("%-95s %-12s %s" % (h, units[i], vals[i]))
rows = list(csv.reader(open(src)))
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, byop, samp = 0, collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= iE:
        continue
    try:
        e = int(r[iE])
    except ValueError:
        continue
    toks = r[iS].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    byop[op] += e
    tot += e
    samp[op] += int(r[iSm] or 0)
print("executed thread-instructions per destination pixel: %.2f" % (tot * 32 / npx))
for op, c in byop.most_common(24):
    print("  %-10s %6.2f /px   stall samples %d" % (op, c * 32 / npx, samp[op]))
