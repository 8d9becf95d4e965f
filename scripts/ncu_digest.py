"""Digest of an `ncu --page raw --csv` dump of scripts/profile_calls.py: one row per profiled kernel with the
counters the roofline argument uses, plus profiles/ncu_traffic.json (DRAM bytes per launch of the dominant kernel of
each entry-point family, read by bench.py for roofline.traffic).
  python scripts/ncu_digest.py gpurun_out/<tag>/full_raw.csv gpurun_out/<tag>/ncu_full.log profiles/<name>"""
import csv
import json
import re
import sys

raw, log, out = sys.argv[1:4]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'time_us'), ('dram__bytes_read.sum', 'dram_read_MB'),
        ('dram__bytes_write.sum', 'dram_write_MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_pct'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex_pct'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wavefronts'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block')]
cols = [(c, n) for c, n in cols if c in idx]
# the entry-point call each kernel belongs to: profile_calls.py prints them in launch order; a call may launch
# several kernels, so kernels are matched to calls by name below
calls = [l.split('profiled: ')[1].strip() for l in open(log) if l.startswith('profiled: ')]


def to_mb(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}[unit]


with open(out + '.csv', 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow([n for _c, n in cols])
    for r in rows[2:]:
        line = []
        for c, n in cols:
            v = r[idx[c]]
            if n.endswith('_MB'):
                v = '%.1f' % to_mb(v, units[idx[c]])
            elif n == 'kernel':
                v = re.sub(r'\(.*', '', v.replace('void ', ''))
            line.append(v)
        w.writerow(line)
print('wrote', out + '.csv', len(rows) - 2, 'kernels;', len(calls), 'calls profiled')
