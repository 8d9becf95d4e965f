"""Task 4 of the seed-0 batch: per-task first-order gradient from engines of different batch sizes, and run-to-run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

spec = pspec.miniimagenet_spec(5)
theta = pspec.init_flat_params(spec, seed=42).cuda()
X, Y = make_tasks(32, 5, 5, (3, 84, 84), seed=0)
X, Y = X.cuda(), Y.cuda()
T, lr, tk = 5, 0.001, int(sys.argv[1]) if len(sys.argv) > 1 else 4
offs, P = spec.param_offsets()
names = []
for l in range(4):
    names += ['bn%d.g' % l, 'bn%d.b' % l, 'conv%d.w' % l, 'conv%d.b' % l]
names += ['lin.w', 'lin.b']


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def row(tasks, lo, which, reps=1):
    e = eng.MamlEngine(spec, tasks, 5, T, lr, mode='first', device='cuda')
    outs = []
    for _ in range(reps):
        e.run(X[lo:lo + tasks], Y[lo:lo + tasks], theta); torch.cuda.synchronize()
        outs.append((e.bar[0][which].clone(), e.loss[which].item(), e.correct[which].item()))
    return outs


r32 = row(32, 0, tk, reps=2)
r4 = row(4, (tk // 4) * 4, tk % 4, reps=2)
r1 = row(1, tk, 0, reps=2)
print('loss/correct 32:', r32[0][1:], ' 4:', r4[0][1:], ' 1:', r1[0][1:])
print('run-to-run  32: %.2e   4: %.2e   1: %.2e' % (rel(r32[0][0], r32[1][0]), rel(r4[0][0], r4[1][0]), rel(r1[0][0], r1[1][0])))
print('32 vs 4: %.2e   32 vs 1: %.2e   4 vs 1: %.2e' % (rel(r32[0][0], r4[0][0]), rel(r32[0][0], r1[0][0]), rel(r4[0][0], r1[0][0])))
for i, n in enumerate(names):
    a, b = offs[i], (offs[i + 1] if i + 1 < len(offs) else P)
    print('%-8s 32v4 %.2e  32v1 %.2e  4v1 %.2e' % (n, rel(r32[0][0][a:b], r4[0][0][a:b]), rel(r32[0][0][a:b], r1[0][0][a:b]),
                                                   rel(r4[0][0][a:b], r1[0][0][a:b])))
