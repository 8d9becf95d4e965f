#!/usr/bin/env python
"""Generates tests/golden/rl/rl_trpo_small.npz from the REFERENCE'S OWN FILES (build container only):

    python tests/golden/make_golden_rl.py            # needs /root/reference

One meta-optimisation of MAML-TRPO (rl/maml_trpo.py:101-140 minus the environment rollouts) on synthetic
Particles2D-style replays, in float64, through the reference's unmodified core_functions/rl.py
(trpo_update, meta_optimize_trpo) and core_functions/policies.py (DiagNormalPolicy) on top of the cherry /
learn2learn restatements.  Before anything is written oracle/rl_oracle.py must reproduce the run."""
import os
import sys
from copy import deepcopy

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from exploring_meta_b200.synthetic import make_replays     # noqa: E402
from oracle import rl_oracle as ro                          # noqa: E402
from oracle import rl_ref_loader                            # noqa: E402

CFG = {'inner_lr': 0.05, 'tau': 1.0, 'gamma': 0.99, 'value_reg': 2, 'max_kl': 0.01, 'ls_max_steps': 15,
       'backtrack_factor': 0.5, 'outer_lr': 0.3}
TASKS, EPISODES, HORIZON, SEED = 3, 4, 25, 0


def main():
    ns = rl_ref_loader.load()
    rl, pol, ch = ns.rl, ns.policies, ns.cherry
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(42)
    policy = pol.DiagNormalPolicy(2, 2, activation='tanh').double()
    theta0 = [p.detach().clone() for p in policy.parameters()]
    baseline = ch.LinearValue(2, CFG['value_reg']).double()
    data = make_replays(TASKS, EPISODES, HORIZON, seed=SEED, dtype=torch.float64)
    params = {'inner_lr': CFG['inner_lr'], 'gamma': CFG['gamma'], 'tau': CFG['tau'], 'max_kl': CFG['max_kl'],
              'ls_max_steps': CFG['ls_max_steps'], 'backtrack_factor': CFG['backtrack_factor'], 'outer_lr': CFG['outer_lr']}
    iter_replays, iter_policies = [], []
    for sup, qry in data:
        mk = lambda r: ch.Replay(r['states'], r['actions'], r['rewards'], r['dones'], r['next_states'])   # noqa: E731
        sup_r, qry_r = mk(sup), mk(qry)
        learner = deepcopy(policy)                                   # rl/maml_trpo.py:107
        learner = rl.trpo_update(sup_r, learner, baseline, CFG['inner_lr'], CFG['gamma'], CFG['tau'], first_order=True)
        iter_replays.append([sup_r, qry_r])
        iter_policies.append(learner)
    old_params = [[p.detach().clone() for p in lp.parameters()] for lp in iter_policies]
    old_loss, old_kl = rl.meta_surrogate_loss(iter_replays, iter_policies, policy, baseline, params, False)
    rl.meta_optimize_trpo(params, policy, baseline, iter_replays, iter_policies)
    theta1 = [p.detach().clone() for p in policy.parameters()]

    # ---- the self-contained restatement must reproduce the reference-file run ------------------------------------
    o_old = [ro.trpo_update([p.clone().requires_grad_() for p in theta0], sup, CFG['inner_lr'], CFG['tau'], CFG['gamma'],
                            CFG['value_reg'], first_order=True) for sup, _q in data]
    for a, b in zip(o_old, old_params):
        for x, y in zip(a, b):
            assert torch.allclose(x.detach(), y, rtol=1e-10, atol=1e-12)
    new, diag = ro.meta_optimize_trpo(theta0, [[s, q] for s, q in data], [[x.detach() for x in a] for a in o_old], CFG)
    assert abs(diag['old_loss'] - float(old_loss)) < 1e-10 and abs(diag['old_kl'] - float(old_kl)) < 1e-12
    for x, y in zip(new, theta1):
        assert torch.allclose(x, y, rtol=1e-8, atol=1e-10), (x - y).abs().max()
    print('restatement == reference files: old_loss %.12f  old_kl %.3e  line-search step %d' % (
        diag['old_loss'], diag['old_kl'], diag['ls_step']))
    flat = lambda ps: torch.cat([p.reshape(-1) for p in ps]).numpy()                                   # noqa: E731
    np.savez_compressed(os.path.join(HERE, 'rl', 'rl_trpo_small.npz'), theta0=flat(theta0), theta1=flat(theta1),
                        old_params=np.stack([flat(p) for p in old_params]), old_loss=float(old_loss),
                        old_kl=float(old_kl), grad=diag['grad'].numpy(), step=diag['step'].numpy(),
                        ls_step=diag['ls_step'], tasks=TASKS, episodes=EPISODES, horizon=HORIZON, seed=SEED)
    print('wrote rl_trpo_small.npz')


if __name__ == '__main__':
    main()
