#!/usr/bin/env python
"""Generates tests/golden/rl/rl_ppo_small.npz from the REFERENCE'S OWN FILES (build container only):

    python tests/golden/make_golden_ppo.py            # needs /root/reference

The train half of one MAML-PPO / ANIL-PPO outer iteration (rl/maml_ppo.py:95-130, rl/anil_ppo.py:98-131) minus the
environment: per task ``policy.clone()`` (learn2learn MAML restatement), the reference's unmodified
``core_functions/rl.py::fast_adapt_ppo`` with a stub task whose ``run()`` returns the seeded synthetic replays, the mean
validation loss, ``backward()``.  float64.  Before anything is written oracle/rl_oracle.py::fast_adapt_ppo must
reproduce the run (validation losses, adapted parameters, meta-gradient)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from exploring_meta_b200.synthetic import make_replays     # noqa: E402
from oracle import l2l_shim                                 # noqa: E402
from oracle import rl_oracle as ro                          # noqa: E402
from oracle import rl_ref_loader                            # noqa: E402

CFG = {'inner_lr': 0.05, 'tau': 1.0, 'gamma': 0.99, 'value_reg': 2, 'ppo_epochs': 3, 'ppo_clip_ratio': 0.1,
       'adapt_steps': 1, 'adapt_batch_size': 4, 'max_path_length': 25}
TASKS, EPISODES, HORIZON, SEED = 3, 4, 25, 2


class StubTask:
    """``task.run(learner, episodes=...)``: the support replays in order (one per adaptation step), then the query replay."""
    def __init__(self, ch, sups, qry):
        mk = lambda r: ch.Replay(r['states'], r['actions'], r['rewards'], r['dones'], r['next_states'])   # noqa: E731
        self.queue = [mk(r) for r in sups] + [mk(qry)]

    def run(self, learner, episodes=None, render=False):
        return self.queue.pop(0)


def main():
    ns = rl_ref_loader.load()
    rl, pol, ch = ns.rl, ns.policies, ns.cherry
    rl.set_device(torch.device('cpu'))
    torch.set_default_dtype(torch.float64)
    data = make_replays(TASKS, EPISODES, HORIZON, seed=SEED, dtype=torch.float64)
    data2 = make_replays(TASKS, EPISODES, HORIZON, seed=SEED + 100, dtype=torch.float64)     # second-step support replays
    out = {}
    for anil, steps in ((False, 1), (True, 1), (False, 2), (True, 2)):
        cfg = dict(CFG, adapt_steps=steps)
        sups = [[sup] if steps == 1 else [sup, data2[t][0]] for t, (sup, _q) in enumerate(data)]
        torch.manual_seed(42)
        policy = pol.DiagNormalPolicyANIL(2, 2, 100) if anil else pol.DiagNormalPolicy(2, 2, activation='tanh')
        policy = policy.double()
        policy.sigma.data += torch.tensor([0.1, -0.2])
        theta0 = [p.detach().clone() for p in policy.parameters()]
        maml = l2l_shim.MAML(policy, lr=CFG['inner_lr'])
        baseline = ch.LinearValue(2, CFG['value_reg']).double()
        losses, adapted = [], []
        total = 0.0
        for t, (_sup, qry) in enumerate(data):
            learner = maml.clone()
            loss, _rew, _suc = rl.fast_adapt_ppo(StubTask(ch, sups[t], qry), learner, baseline, cfg, anil=anil)
            losses.append(float(loss))
            adapted.append(torch.cat([p.detach().reshape(-1) for p in learner.module.parameters()]))
            total = total + loss
        (total / TASKS).backward()
        grad = torch.cat([p.grad.reshape(-1) for p in policy.parameters()])
        # ---- the restatement must reproduce the reference-file run ------------------------------------------
        ps = [p.clone().requires_grad_() for p in theta0]
        tot = 0.0
        for t, (_sup, qry) in enumerate(data):
            v, new = ro.fast_adapt_ppo(ps, sups[t], qry, CFG, anil=anil)
            assert abs(float(v) - losses[t]) < 1e-12, (float(v), losses[t])
            assert torch.allclose(torch.cat([x.detach().reshape(-1) for x in new]), adapted[t], rtol=1e-10, atol=1e-12)
            tot = tot + v
        g_or = torch.cat([g.reshape(-1) for g in torch.autograd.grad(tot / TASKS, ps)])
        assert torch.allclose(g_or, grad, rtol=1e-9, atol=1e-12), (g_or - grad).abs().max()
        key = ('anil' if anil else 'maml') + ('2' if steps == 2 else '')
        out[key + '_theta0'] = torch.cat([p.reshape(-1) for p in theta0]).numpy()
        out[key + '_adapted'] = torch.stack(adapted).numpy()
        out[key + '_valid_loss'] = np.array(losses)
        out[key + '_grad'] = grad.numpy()
        print('%s: restatement == reference files; valid losses %s, |grad| %.3e' % (key, np.round(losses, 6), float(grad.norm())))
    np.savez_compressed(os.path.join(HERE, 'rl', 'rl_ppo_small.npz'), tasks=TASKS, episodes=EPISODES, horizon=HORIZON,
                        seed=SEED, seed2=SEED + 100, **out)
    print('wrote rl_ppo_small.npz')


if __name__ == '__main__':
    main()
