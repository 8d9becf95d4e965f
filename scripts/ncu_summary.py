"""Prints the metrics that matter from an `ncu --page raw --csv` dump: python scripts/ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct',
        'sm__cycles_elapsed.max', 'smsp__average_warp', 'smsp__average_warps_issue_stalled', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__inst_executed_pipe_', 'smsp__inst_executed_pipe_']
for w in want:
    for c in [h for h in hdr if h.startswith(w)]:
        vals = [r[idx[c]] for r in rows[2:]]
        if all(v in ('0', '', 'n/a') for v in vals):
            continue
        print('%-95s %-12s %s' % (c[:95], units[idx[c]], vals))
