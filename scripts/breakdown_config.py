"""Per-call device time of the launch program of another BASELINE config:  python scripts/breakdown_config.py 4"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

cfg = int(sys.argv[1])
if cfg == 1:
    spec, tasks, ways, shots = pspec.omniglot_spec(5), 32, 5, 1
    e = eng.MamlEngine(spec, tasks, shots, 1, 0.5, device='cuda')
elif cfg == 4:
    spec, tasks, ways, shots = pspec.omniglot_spec(20), 256, 20, 5
    e = eng.MamlEngine(spec, tasks, shots, 1, 0.5, device='cuda')
else:
    spec, tasks, ways, shots = pspec.anil_body_spec('min', 5), 32, 5, 5
    e = eng.AnilEngine(spec, tasks, shots, 1, 0.5, device='cuda')
    e.head.normal_(0, 0.05)
X, Y = make_tasks(tasks, ways, shots, (spec.in_c, spec.in_h, spec.in_w), seed=0)
e.x.copy_(X); e.y.copy_(Y); e.theta.copy_(pspec.init_flat_params(spec))
times = bench.kernel_breakdown(e, reps=2)
groups = collections.OrderedDict()
for idx, (fn, args, name) in enumerate(e.prog.calls):
    fam = bench.family(name, args)
    ms = times[fam]['per_call'][idx]
    key = name
    if hasattr(args, 'g'):
        g = args.g
        key += ' cin%d->%d %dx%d s%d' % (g.cin, g.cout, g.hin, g.win, g.stride)
        if name == 'xm_conv':
            key += (' dgrad' if args.mode == 1 else ' fwd') + (' x2' if args.src2 else '')
        if name == 'xm_wgrad':
            key += ' x2' if args.x2 else ''
    d = groups.setdefault(key, [0.0, 0])
    d[0] += ms; d[1] += 1
total = sum(d[0] for d in groups.values())
print('config %d total %.3f ms' % (cfg, total))
for k, (ms, n) in sorted(groups.items(), key=lambda kv: -kv[1][0])[:16]:
    print('%-44s %3d calls %9.3f ms %8.3f ms/call %5.1f%%' % (k, n, ms, ms / n, 100 * ms / total))
