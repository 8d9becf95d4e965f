"""One meta-optimisation of the config-5 path (40 tasks x 2 x 2000 transitions) for an ncu launch list / full capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from exploring_meta_b200.rl_engine import TrpoEngine
from exploring_meta_b200.synthetic import make_replays
from oracle import rl_oracle as ro

RL = bench.RL
e = TrpoEngine(RL['tasks'], RL['episodes'] * RL['horizon'], 2, 2, (100, 100), 'tanh', RL['inner_lr'], RL['gamma'], RL['tau'],
               RL['value_reg'], device='cuda')
e.load_replays(make_replays(RL['tasks'], RL['episodes'], RL['horizon'], seed=0))
theta = torch.cat([p.reshape(-1) for p in ro.init_policy(seed=42)]).cuda()
old = e.adapt(theta).clone()
e.set_old_policies(old)
new, diag = e.meta_optimize(theta, RL['max_kl'], RL['ls_max_steps'], RL['backtrack_factor'], RL['outer_lr'])
torch.cuda.synchronize()
print('done', diag['ls_step'])
