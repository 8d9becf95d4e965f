"""Digest of an `ncu --page raw --csv` dump of scripts/profile_calls.py: one row per profiled kernel with the
counters the roofline argument uses, plus profiles/ncu_traffic.json (DRAM bytes per launch of the dominant kernel of
each entry-point family, read by bench.py for roofline.traffic).
  python scripts/ncu_digest.py gpurun_out/<tag>/full_raw.csv gpurun_out/<tag>/ncu_full.log profiles/<name>"""
import csv
import json
import re
import sys

raw, log, out = sys.argv[1:4]
stamp = sys.argv[4] if len(sys.argv) > 4 else None        # e.g. the git SHA the capture was taken at
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
calls = []
for l in open(log):
    if l.startswith('profiled: '):
        parts = [x.strip() for x in l.split('profiled: ')[1].split('|')]
        nk = int(parts[1].split()[1]) if len(parts) > 1 else None
        cps = int(parts[2].split()[1]) if len(parts) > 2 else 0
        calls.append((parts[0], nk, cps))


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

# ---- DRAM traffic per entry-point call, averaged per family over the calls of one config-2 step -------------
if calls and all(nk is not None for _k, nk, _c in calls) and sum(nk for _k, nk, _c in calls) == len(rows) - 2:
    def family(key):
        name = key.split()[0]
        if name in ('xm_conv', 'xm_wgrad') and ' cin3 ' in key + ' ':
            return name + ':image_layer'
        return name
    fams, r = {}, 2
    for key, nk, cps in calls:
        dram = 0.0
        names = []
        for rr in rows[r:r + nk]:
            dram += to_mb(rr[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) * 1e6
            dram += to_mb(rr[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']]) * 1e6
            names.append(re.sub(r'\(.*', '', rr[idx['Kernel Name']].replace('void ', '')))
        r += nk
        f = fams.setdefault(family(key), {'bytes': 0.0, 'calls': 0, 'per_call': {}})
        f['bytes'] += dram * cps
        f['calls'] += cps
        f['per_call'][key] = {'dram_bytes': dram, 'calls_per_step': cps, 'kernels': names}
    traffic = {k: {'dram_bytes_per_launch': v['bytes'] / max(v['calls'], 1), 'calls_per_step': v['calls'],
                   'per_call': v['per_call']} for k, v in fams.items()}
    if stamp:
        traffic['_meta'] = {'captured_at': stamp, 'source': raw}
    with open(out + '_traffic.json', 'w') as f:
        json.dump(traffic, f, indent=1)
    print('wrote', out + '_traffic.json')
else:
    print('no per-call kernel counts in the log: traffic json not written')
