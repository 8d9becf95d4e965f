"""Resident-input throughput of the other BASELINE.json configs on one GPU (parity cases of bench.py's headline
config; reported under profiles/, not a bench line):  python scripts/bench_configs.py [--steps 10]
  cfg 1  MAML Omniglot 5-way 1-shot, 64-filter stride-2 CNN, meta-batch 32, 1 inner step, second order
  cfg 3  ANIL Mini-ImageNet 5-way 5-shot, 64-filter body forward once + head-only adaptation, meta-batch 32
  cfg 4  MAML Omniglot 20-way 5-shot, meta-batch 256 (the per-GPU shard at 1 GPU), 1 inner step, second order"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from exploring_meta_b200 import spec as pspec
from exploring_meta_b200.synthetic import make_tasks
from exploring_meta_b200.trainer import AnilTrainer, MamlTrainer

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=10)
a = ap.parse_args()


def timed(tr, X, Y, steps):
    e = tr.engine
    e.x.copy_(X); e.y.copy_(Y)
    for _ in range(3):
        tr.meta_step(e.x, e.y)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        tr.meta_step(e.x, e.y)
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


out = []
s1 = pspec.omniglot_spec(5)
tr = MamlTrainer(s1, 32, 1, 1, 0.5, device='cuda')
tr.theta.copy_(pspec.init_flat_params(s1))
X, Y = make_tasks(32, 5, 1, (1, 28, 28), seed=0)
ms = timed(tr, X.cuda(), Y.cuda(), a.steps)
out.append({'config': 'cfg1 MAML Omniglot 5w1s B=32 T=1', 'ms_per_step': ms, 'tasks_per_s': 32e3 / ms,
            'launches': tr.engine.launches_per_run, 'image_block': tr.engine.img})
del tr

s3 = pspec.anil_body_spec('min', 5)
tr = AnilTrainer(s3, 32, 5, 1, 0.5, device='cuda')
body = pspec.init_flat_params(s3)
head = torch.nn.Linear(tr.engine.D, 5)
torch.nn.init.xavier_uniform_(head.weight)
tr.theta_all.copy_(torch.cat([body, head.weight.detach().reshape(-1), torch.zeros(5)]))
X, Y = make_tasks(32, 5, 5, (3, 84, 84), seed=0)
ms = timed(tr, X.cuda(), Y.cuda(), a.steps)
out.append({'config': 'cfg3 ANIL Mini-ImageNet 5w5s B=32 (64-filter body)', 'ms_per_step': ms, 'tasks_per_s': 32e3 / ms,
            'launches': tr.engine.launches_per_run, 'image_block': tr.engine.img})
del tr

s4 = pspec.omniglot_spec(20)
tr = MamlTrainer(s4, 256, 5, 1, 0.5, device='cuda')
tr.theta.copy_(pspec.init_flat_params(s4))
X, Y = make_tasks(256, 20, 5, (1, 28, 28), seed=0)
ms = timed(tr, X.cuda(), Y.cuda(), a.steps)
out.append({'config': 'cfg4 MAML Omniglot 20w5s B=256 T=1', 'ms_per_step': ms, 'tasks_per_s': 256e3 / ms,
            'launches': tr.engine.launches_per_run, 'image_block': tr.engine.img})
for o in out:
    print(json.dumps(o))
