"""Decision-flip statistics at the config-2 shape (MiniImagenetCNN, 5-way 5-shot, T = 5) and the calm inner lr:
for each data seed one task is run by (a) the fp64 oracle, (b) the fp32 oracle = the reference's own arithmetic,
(c) the CUDA path (twice, to expose run-to-run differences from the atomics' order).  Prints the meta-gradient
rel-L2 of (b) and (c) against (a): values ~1e-6 are rounding, values >= 1e-4 are a flipped ReLU / max-pool
decision somewhere in the 5 x 4 x 1.4 M pre-activations.
  python scripts/diag_flip_rate.py [--seeds 16] [--lr 0.001] [--tasks 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks
from oracle import maml_oracle as mo

ap = argparse.ArgumentParser()
ap.add_argument('--seeds', type=int, default=16)
ap.add_argument('--seed0', type=int, default=100)
ap.add_argument('--lr', type=float, default=0.001)
ap.add_argument('--tasks', type=int, default=1)
ap.add_argument('--steps', type=int, default=5)
a = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)
spec = pspec.miniimagenet_spec(5)
ospec = mo.miniimagenet_spec(5)
params = mo.init_params(ospec, seed=42)
mask = ~mo.conv_bias_mask(ospec)
e = eng.MamlEngine(spec, a.tasks, 5, a.steps, a.lr, mode='second', device='cuda')
rows = []
for seed in range(a.seed0, a.seed0 + a.seeds):
    X, Y = make_tasks(a.tasks, 5, 5, (3, 84, 84), seed=seed)
    r64 = mo.meta_iteration([p.double() for p in params], X.double(), Y, ospec, a.steps, a.lr)
    r32 = mo.meta_iteration(params, X, Y, ospec, a.steps, a.lr)
    g64 = mo.flatten(r64['grad'])[mask]
    e_ref = mo.rel_l2(mo.flatten(r32['grad'])[mask], g64)
    ours = []
    for rep in range(2):
        e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda())
        torch.cuda.synchronize()
        ours.append(mo.rel_l2(e.grad.cpu()[mask], g64))
    dl = float((e.loss.cpu().double() - r64['loss']).abs().max())
    rows.append((seed, e_ref, ours[0], ours[1]))
    print('seed %4d  e_ref %.2e  e_new %.2e / %.2e  |dloss| %.1e  correct %s/%s' % (
        seed, e_ref, ours[0], ours[1], dl, e.correct.cpu().tolist(), r64['correct'].tolist()), flush=True)
thr = 1e-4
print('flipped (> %.0e): reference fp32 %d / %d, CUDA path %d / %d (first run), %d / %d (second run)' % (
    thr, sum(r[1] > thr for r in rows), len(rows), sum(r[2] > thr for r in rows), len(rows),
    sum(r[3] > thr for r in rows), len(rows)))
