"""Runs one eager (un-captured) replay of the config-2 launch program, for ncu:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python scripts/profile_step.py --inner-steps 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from exploring_meta_b200 import _lib, engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

ap = argparse.ArgumentParser()
ap.add_argument('--tasks', type=int, default=32)
ap.add_argument('--inner-steps', type=int, default=1)
ap.add_argument('--reps', type=int, default=1)
ap.add_argument('--fast-tf32', action='store_true')
a = ap.parse_args()
_lib.load().xm_set_precision(0 if a.fast_tf32 else 1)
spec = pspec.miniimagenet_spec(5)
e = eng.MamlEngine(spec, a.tasks, 5, a.inner_steps, 0.5, device='cuda')
X, Y = make_tasks(a.tasks, 5, 5, (3, 84, 84), seed=0)
e.x.copy_(X); e.y.copy_(Y); e.theta.copy_(pspec.init_flat_params(spec))
for _ in range(a.reps):
    e.prog.replay(torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print('done', float(e.loss.mean()))
