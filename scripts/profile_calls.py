"""One launch of every distinct (entry point, layer geometry, variant) of the config-2 launch program, inside a
cudaProfilerStart/Stop range, for an `ncu --set full --profile-from-start off` capture:

  ncu --set full --clock-control none --profile-from-start off -o /tmp/full python scripts/profile_calls.py
  ncu -i /tmp/full.ncu-rep --page raw --csv > gpurun_out/full_raw.csv

The whole program is replayed once first so that every buffer a selected call reads holds real data."""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from exploring_meta_b200 import _lib, engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

ap = argparse.ArgumentParser()
ap.add_argument('--tasks', type=int, default=32)
ap.add_argument('--only', default='', help='comma-separated substrings of the call keys to keep')
a = ap.parse_args()
spec = pspec.miniimagenet_spec(5)
e = eng.MamlEngine(spec, a.tasks, 5, 1, 0.5, device='cuda')
X, Y = make_tasks(a.tasks, 5, 5, (3, 84, 84), seed=0)
e.x.copy_(X); e.y.copy_(Y); e.theta.copy_(pspec.init_flat_params(spec))
stream = torch.cuda.current_stream().cuda_stream
e.prog.replay(stream)
torch.cuda.synchronize()


def key(name, args):
    k = name
    if hasattr(args, 'g'):
        g = args.g
        k += ' cin%d %dx%d' % (g.cin, g.hin, g.win)
        if name == 'xm_conv':
            k += (' dgrad' if args.mode == 1 else ' fwd') + (' x2' if args.src2 else '')
        if name == 'xm_wgrad':
            k += ' x2' if args.x2 else ''
    elif name == 'xm_head':
        k += ' dual' if args.dual else ''
    return k


chosen, seen = [], set()
for fn, args, name in e.prog.calls:
    if isinstance(args, tuple):
        continue
    k = key(name, args)
    if k in seen or (a.only and not any(s in k for s in a.only.split(','))):
        continue
    seen.add(k)
    chosen.append((fn, args, k))
# how often each distinct call occurs in the full config-2 program (5 inner steps): weights for per-family averages
e5 = eng.MamlEngine(spec, a.tasks, 5, 5, 0.5, device='cuda')
count5 = {}
for _fn, args, name in e5.prog.calls:
    if not isinstance(args, tuple):
        count5[key(name, args)] = count5.get(key(name, args), 0) + 1
del e5
lib = _lib.load()
kernels = []
torch.cuda.profiler.start()
for fn, args, k in chosen:
    n0 = lib.xm_launch_count()
    _lib.check(fn(ctypes.byref(args), stream), k)
    kernels.append(lib.xm_launch_count() - n0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
for (_fn, _a, k), n in zip(chosen, kernels):
    print('profiled: %s | kernels %d | calls_per_step %d' % (k, n, count5.get(k, 0)))
