"""Per-call device time of one config-2 launch program, grouped by (entry point, layer geometry, variant):
  python scripts/step_breakdown.py [--inner-steps 5] [--tasks 32]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

ap = argparse.ArgumentParser()
ap.add_argument('--tasks', type=int, default=32)
ap.add_argument('--inner-steps', type=int, default=5)
a = ap.parse_args()
spec = pspec.miniimagenet_spec(5)
e = eng.MamlEngine(spec, a.tasks, 5, a.inner_steps, 0.5, device='cuda')
X, Y = make_tasks(a.tasks, 5, 5, (3, 84, 84), seed=0)
e.x.copy_(X); e.y.copy_(Y); e.theta.copy_(pspec.init_flat_params(spec))
times = bench.kernel_breakdown(e, reps=3)
work = bench.program_work(e)
groups = collections.OrderedDict()
for idx, (fn, args, name) in enumerate(e.prog.calls):
    fam = bench.family(name, args)
    ms = times[fam]['per_call'][idx]
    key = name
    if hasattr(args, 'g'):
        g = args.g
        key += ' cin%d %dx%d' % (g.cin, g.hin, g.win)
        if name == 'xm_conv':
            key += ' dgrad' if args.mode == 1 else ' fwd'
            key += ' x2' if args.src2 else ''
        if name == 'xm_wgrad':
            key += ' x2' if args.x2 else ''
    elif name == 'xm_head':
        key += ' dual' if args.dual else ''
    d = groups.setdefault(key, [0.0, 0, 0.0, 0.0])
    d[0] += ms; d[1] += 1
    w = work.get(fam, {}).get('per_call', {}).get(idx)
    if w:
        if name in ('xm_conv', 'xm_wgrad'):
            d[2] += w
        else:
            d[3] += w
total = sum(d[0] for d in groups.values())
print('total %.3f ms' % total)
for k, (ms, n, fl, by) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    extra = ''
    if fl:
        extra = '%7.1f TFLOP/s' % (fl / ms / 1e9)
    if by:
        extra = '%7.1f GB/s (algorithmic)' % (by / ms / 1e6)
    print('%-40s %3d calls %8.3f ms  %6.3f ms/call %5.1f%%  %s' % (k, n, ms, ms / n, 100 * ms / total, extra))
