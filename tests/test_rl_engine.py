"""Config 5 (MAML-TRPO policy MLP): the product's launch sequences (exploring_meta_b200/rl_engine.py) against the
oracle (oracle/rl_oracle.py, float64) and the golden fixture generated from the reference's own rl.py / policies.py
(tests/golden/rl/rl_trpo_small.npz).  ``kdev`` runs them on the CPU emulator of the C ABI (``-m "not gpu"``: checks
the forward-over-reverse algebra, the Gauss-Newton Fisher product, CG and the line search) and on the CUDA kernels
(``-m gpu``: the parity test proper).

Tolerances (float32 path against the float64 oracle): adapted parameters 1e-5, meta-gradient 1e-4 rel-L2, the
parameters after one meta-optimisation 1e-4 rel-L2 of the UPDATE (theta1 - theta0), same accepted line-search step."""
import os

import numpy as np
import pytest
import torch

from exploring_meta_b200.rl_engine import TrpoEngine
from exploring_meta_b200.synthetic import make_replays
from oracle import cherry_shim as ch
from oracle import rl_oracle as ro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rl', 'rl_trpo_small.npz')
CFG = {'inner_lr': 0.05, 'tau': 1.0, 'gamma': 0.99, 'value_reg': 2, 'max_kl': 0.01, 'ls_max_steps': 15,
       'backtrack_factor': 0.5, 'outer_lr': 0.3}


def flat(ps):
    return torch.cat([p.detach().reshape(-1) for p in ps])


def rel(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _engine(tasks, episodes, horizon, seed, dev, activation='tanh'):
    data32 = make_replays(tasks, episodes, horizon, seed=seed)
    data64 = make_replays(tasks, episodes, horizon, seed=seed, dtype=torch.float64)
    e = TrpoEngine(tasks, episodes * horizon, 2, 2, (100, 100), activation, CFG['inner_lr'], CFG['gamma'], CFG['tau'],
                   CFG['value_reg'], device=dev)
    e.load_replays(data32)
    return e, data64


def test_advantages_match_cherry_restatement(kdev):
    e, data = _engine(3, 4, 25, 0, kdev)
    for t, (sup, qry) in enumerate(data):
        for k, rep in enumerate((sup, qry)):
            adv = ch.normalize(ro.compute_advantages(rep, CFG['tau'], CFG['gamma'], CFG['value_reg'])).reshape(-1)
            scale = -1.0 / e.n if k == 0 else -1.0 / (e.n * e.tasks)
            assert rel(e.coef[k, t], scale * adv) < 2e-5


def test_trpo_update_matches_oracle(kdev):
    e, data = _engine(3, 4, 25, 0, kdev)
    theta = ro.init_policy(dtype=torch.float64, seed=42)
    th32 = flat(theta).float().to(kdev)
    out = e.adapt(th32)
    for t, (sup, _q) in enumerate(data):
        ref = ro.trpo_update([p.clone().requires_grad_() for p in theta], sup, CFG['inner_lr'], CFG['tau'], CFG['gamma'],
                             CFG['value_reg'], first_order=True)
        assert rel(out[t], flat(ref)) < 1e-5
        # the adaptation step itself (theta' - theta = -lr * gradient): 1e-4 of its norm + the fp32 rounding of theta'
        d, dref = out[t].cpu().double() - flat(theta), flat(ref) - flat(theta)
        assert float((d - dref).norm()) <= 1e-4 * float(dref.norm()) + 2e-7 * float(flat(theta).norm())


@pytest.mark.parametrize('activation', ['tanh', 'relu'])
def test_meta_gradient_and_fisher_product_match_oracle(kdev, activation):
    import math
    act = torch.tanh if activation == 'tanh' else torch.relu
    e, data = _engine(2, 3, 20, 5, kdev, activation)
    theta = [p.requires_grad_() for p in ro.init_policy(dtype=torch.float64, seed=3)]
    theta[0].data += torch.tensor([0.2, -0.3], dtype=torch.float64)          # sigma != 0: exercises the log-std terms
    saved = ro.policy_mean
    ro.policy_mean = lambda params, states, activation=act: saved(params, states, activation)
    try:
        old = [[x.detach() for x in ro.trpo_update(theta, sup, CFG['inner_lr'], CFG['tau'], CFG['gamma'], CFG['value_reg'],
                                                   first_order=True)] for sup, _q in data]
        loss, kl = ro.meta_surrogate_loss(theta, [[s, q] for s, q in data], old, CFG)
        g_ref = flat(torch.autograd.grad(loss, theta, retain_graph=True))
        fvp = ch.hessian_vector_product(kl, theta, damping=1e-5)
        torch.manual_seed(1)
        v = torch.randn(g_ref.numel(), dtype=torch.float64)
        hv_ref = fvp(v)
        # a perturbed point for the line-search quantities (KL > 0 there)
        cand = [(p.detach() - 0.05 * torch.randn_like(p) * p.detach().abs().mean()).requires_grad_() for p in theta]
        loss_c, kl_c = ro.meta_surrogate_loss(cand, [[s, q] for s, q in data], old, CFG)
    finally:
        ro.policy_mean = saved
    th32 = flat(theta).float().to(kdev)
    e.set_old_policies(torch.stack([flat(o) for o in old]).float().to(kdev))
    l, k, g = e.gradient(th32)
    assert abs(float(l) - float(loss)) < 1e-5 * max(1.0, abs(float(loss))) and abs(float(k)) < 1e-9
    assert rel(g, g_ref) < 1e-4
    hv = e.fvp(th32, v.float().to(kdev))
    assert rel(hv, hv_ref) < 1e-4
    lc, kc = e.loss_and_kl(flat(cand).float().to(kdev))
    assert abs(float(lc) - float(loss_c)) < 2e-5 * max(1.0, abs(float(loss_c)))
    assert abs(float(kc) - float(kl_c)) < 1e-4 * abs(float(kl_c)) + 1e-9


def test_meta_optimize_trpo_matches_reference_fixture(kdev):
    g = np.load(GOLD)
    e, _data = _engine(int(g['tasks']), int(g['episodes']), int(g['horizon']), int(g['seed']), kdev)
    theta0 = torch.from_numpy(g['theta0']).float().to(kdev)
    old = e.adapt(theta0).clone()                                       # fast_adapt_trpo(first_order=True): rl/maml_trpo.py:107-111
    assert rel(old, torch.from_numpy(g['old_params'])) < 1e-5
    e.set_old_policies(old)
    new, diag = e.meta_optimize(theta0, CFG['max_kl'], CFG['ls_max_steps'], CFG['backtrack_factor'], CFG['outer_lr'])
    assert abs(diag['old_loss'] - float(g['old_loss'])) < 1e-5 and abs(diag['old_kl']) < 1e-9
    assert rel(diag['grad'], torch.from_numpy(g['grad'])) < 1e-4
    assert rel(diag['step'], torch.from_numpy(g['step'])) < 2e-3        # 10 CG iterations on a damped 1e-5 system
    assert diag['ls_step'] == int(g['ls_step'])
    upd_ref = torch.from_numpy(g['theta1'] - g['theta0'])
    assert rel(new.cpu().double() - torch.from_numpy(g['theta0']), upd_ref) < 2e-3
    assert rel(new, torch.from_numpy(g['theta1'])) < 1e-4


def test_reference_api_surface(kdev):
    """The functions a user of core_functions/rl.py calls, with the reference's signatures: trpo_update on a deepcopy
    of the policy per task (rl/maml_trpo.py:107-111), then meta_optimize_trpo mutating the policy in place."""
    from copy import deepcopy
    from exploring_meta_b200.core_functions import rl as xrl
    from exploring_meta_b200.core_functions.policies import DiagNormalPolicy, LinearValue
    g = np.load(GOLD)
    torch.manual_seed(42)
    policy = DiagNormalPolicy(2, 2, activation='tanh').to(kdev)
    assert [k for k, _ in policy.named_parameters()] == ['sigma', 'mean.0.weight', 'mean.0.bias', 'mean.2.weight',
                                                         'mean.2.bias', 'mean.4.weight', 'mean.4.bias']
    policy.load_flat_parameters(torch.from_numpy(g['theta0']).float().to(kdev))      # the fixture's float64-initialised policy
    baseline = LinearValue(2, CFG['value_reg'])
    params = dict(CFG)
    data = make_replays(int(g['tasks']), int(g['episodes']), int(g['horizon']), seed=int(g['seed']))
    iter_replays, iter_policies = [], []
    for sup, qry in data:
        sup_r, qry_r = ch.Replay(**{k: sup[k] for k in ('states', 'actions', 'rewards', 'dones', 'next_states')}), qry
        learner = xrl.trpo_update(sup_r, deepcopy(policy), baseline, CFG['inner_lr'], CFG['gamma'], CFG['tau'],
                                  first_order=True)
        iter_replays.append([sup_r, qry_r])
        iter_policies.append(learner)
    assert rel(torch.stack([p.flat_parameters() for p in iter_policies]), torch.from_numpy(g['old_params'])) < 1e-5
    loss, kl = xrl.meta_surrogate_loss(iter_replays, iter_policies, policy, baseline, params, False)
    assert abs(float(loss) - float(g['old_loss'])) < 1e-5 and abs(float(kl)) < 1e-9
    adv = xrl.compute_advantages(baseline, CFG['tau'], CFG['gamma'], data[0][0]['rewards'].to(kdev),
                                 data[0][0]['dones'].to(kdev), data[0][0]['states'].to(kdev),
                                 data[0][0]['next_states'].to(kdev))
    d64 = make_replays(int(g['tasks']), int(g['episodes']), int(g['horizon']), seed=int(g['seed']), dtype=torch.float64)
    assert rel(adv, ro.compute_advantages(d64[0][0], CFG['tau'], CFG['gamma'], CFG['value_reg'])) < 2e-5
    diag = xrl.meta_optimize_trpo(params, policy, baseline, iter_replays, iter_policies)
    assert diag['ls_step'] == int(g['ls_step'])
    assert rel(policy.flat_parameters(), torch.from_numpy(g['theta1'])) < 1e-4


GOLD_PPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rl', 'rl_ppo_small.npz')
PPO_CFG = {'inner_lr': 0.05, 'tau': 1.0, 'gamma': 0.99, 'value_reg': 2, 'ppo_epochs': 3, 'ppo_clip_ratio': 0.1}


@pytest.mark.parametrize('anil', [False, True])
def test_ppo_inner_loop_and_meta_gradient_match_reference_fixture(kdev, anil):
    """MAML-PPO / ANIL-PPO: the reference's own fast_adapt_ppo + backward on fixed replays
    (tests/golden/make_golden_ppo.py) against the task-batched kernels: adapted parameters after 3 clipped-PPO inner
    steps, validation loss, and the second-order gradient of the mean validation loss w.r.t. the initial parameters."""
    g = np.load(GOLD_PPO)
    key = 'anil' if anil else 'maml'
    tasks, n = int(g['tasks']), int(g['episodes']) * int(g['horizon'])
    e = TrpoEngine(tasks, n, 2, 2, (100, 100), 'tanh', PPO_CFG['inner_lr'], PPO_CFG['gamma'], PPO_CFG['tau'],
                   PPO_CFG['value_reg'], device=kdev)
    e.load_replays(make_replays(tasks, int(g['episodes']), int(g['horizon']), seed=int(g['seed'])))
    theta0 = torch.from_numpy(g[key + '_theta0']).float().to(kdev)
    valid, grad, adapted = e.ppo_meta_gradient(theta0, PPO_CFG['ppo_epochs'], PPO_CFG['ppo_clip_ratio'], anil=anil)
    ref_ad = torch.from_numpy(g[key + '_adapted'])
    assert rel(adapted, ref_ad) < 1e-5
    d, dref = adapted.cpu().double() - torch.from_numpy(g[key + '_theta0']), ref_ad - torch.from_numpy(g[key + '_theta0'])
    assert float((d - dref).norm()) <= 2e-4 * float(dref.norm()) + 2e-7 * float(ref_ad.norm())
    if anil:        # the body did not move
        body = slice(2, 2 + 200 + 100 + 10000 + 100)
        assert torch.equal(adapted[:, body].cpu(), theta0[body].cpu().expand(tasks, -1))
    assert float(valid.abs().max()) < 1e-6                              # normalised advantages: the ratio-1 loss is ~0
    assert rel(grad, torch.from_numpy(g[key + '_grad'])) < 2e-4


@pytest.mark.parametrize('anil', [False, True])
@pytest.mark.parametrize('steps', [1, 2])
def test_fast_adapt_ppo_reference_call_pattern(kdev, anil, steps):
    """rl/maml_ppo.py:103-129 / rl/anil_ppo.py:106-130 with the product's modules: policy.clone(), fast_adapt_ppo on a
    stub task returning the fixture's replays (one or two adaptation steps), mean loss, backward() -> master .grad ==
    the reference's."""
    from exploring_meta_b200.core_functions import rl as xrl
    from exploring_meta_b200.core_functions.maml import MAML
    from exploring_meta_b200.core_functions.policies import DiagNormalPolicy, DiagNormalPolicyANIL, LinearValue
    g = np.load(GOLD_PPO)
    key = ('anil' if anil else 'maml') + ('2' if steps == 2 else '')
    tasks = int(g['tasks'])
    policy = (DiagNormalPolicyANIL(2, 2, 100) if anil else DiagNormalPolicy(2, 2, activation='tanh')).to(kdev)
    policy.load_flat_parameters(torch.from_numpy(g[key + '_theta0']).float().to(kdev))
    maml = MAML(policy, lr=PPO_CFG['inner_lr'])
    baseline = LinearValue(2, PPO_CFG['value_reg'])
    params = dict(PPO_CFG, adapt_steps=steps, adapt_batch_size=int(g['episodes']))
    data = make_replays(tasks, int(g['episodes']), int(g['horizon']), seed=int(g['seed']))
    data2 = make_replays(tasks, int(g['episodes']), int(g['horizon']), seed=int(g['seed2']))

    class StubTask:
        def __init__(self, queue):
            self.queue = list(queue)

        def run(self, learner, episodes=None, render=False):
            return self.queue.pop(0)

    total = 0.0
    for t, (sup, qry) in enumerate(data):
        learner = maml.clone()
        queue = [sup, qry] if steps == 1 else [sup, data2[t][0], qry]
        loss, _rew, _suc = xrl.fast_adapt_ppo(StubTask(queue), learner, baseline, params, anil=anil)
        assert rel(learner.module.flat_parameters(), torch.from_numpy(g[key + '_adapted'][t])) < 1e-5
        total = total + loss
    (total / tasks).backward()
    grad = torch.cat([p.grad.reshape(-1) for p in policy.parameters()])
    assert rel(grad, torch.from_numpy(g[key + '_grad'])) < 2e-4


GOLD_VPG = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rl', 'rl_vpg_small.npz')
VPG_CFG = {'inner_lr': 0.05, 'tau': 1.0, 'gamma': 0.99, 'value_reg': 2}


@pytest.mark.parametrize('anil', [False, True])
@pytest.mark.parametrize('first_order', [False, True])
def test_vpg_adaptation_and_meta_gradient_match_reference_fixture(kdev, anil, first_order):
    """MAML-VPG / ANIL-VPG: the reference's own fast_adapt_vpg + backward on fixed replays
    (tests/golden/make_golden_vpg.py) against the task-batched kernels: adapted parameters after the a2c step on the RAW
    advantages, validation losses, and the (second- or first-order) gradient of the mean validation loss."""
    g = np.load(GOLD_VPG)
    key = ('anil' if anil else 'maml') + ('_fo' if first_order else '')
    tasks, n = int(g['tasks']), int(g['episodes']) * int(g['horizon'])
    e = TrpoEngine(tasks, n, 2, 2, (100, 100), 'tanh', VPG_CFG['inner_lr'], VPG_CFG['gamma'], VPG_CFG['tau'],
                   VPG_CFG['value_reg'], device=kdev)
    e.load_replays(make_replays(tasks, int(g['episodes']), int(g['horizon']), seed=int(g['seed'])), normalize=False)
    t0 = torch.from_numpy(g[key + '_theta0'])
    theta0 = t0.float().to(kdev)
    valid, grad, adapted = e.vpg_meta_gradient(theta0, anil=anil, first_order=first_order)
    ref_ad = torch.from_numpy(g[key + '_adapted'])
    assert rel(adapted, ref_ad) < 1e-5
    d, dref = adapted.cpu().double() - t0, ref_ad - t0
    assert float((d - dref).norm()) <= 2e-4 * float(dref.norm()) + 2e-7 * float(ref_ad.norm())
    if anil:        # the body did not move
        body = slice(2, 2 + 200 + 100 + 10000 + 100)
        assert torch.equal(adapted[:, body].cpu(), theta0[body].cpu().expand(tasks, -1))
    ref_valid = torch.from_numpy(g[key + '_valid_loss']) / tasks          # the engine's losses carry the 1 / tasks of the mean
    assert float((valid.cpu().double() - ref_valid).abs().max()) < 2e-5 * float(ref_valid.abs().max()) + 1e-7
    assert rel(grad, torch.from_numpy(g[key + '_grad'])) < 2e-4


@pytest.mark.parametrize('anil', [False, True])
@pytest.mark.parametrize('steps', [1, 2])
def test_fast_adapt_vpg_reference_call_pattern(kdev, anil, steps):
    """rl/maml_vpg.py / rl/anil_vpg.py with the product's modules: policy.clone(), fast_adapt_vpg on a stub task
    returning the fixture's replays (one or two adaptation steps), mean loss, backward() -> master .grad == the
    reference's."""
    from exploring_meta_b200.core_functions import rl as xrl
    from exploring_meta_b200.core_functions.maml import MAML
    from exploring_meta_b200.core_functions.policies import DiagNormalPolicy, DiagNormalPolicyANIL, LinearValue
    g = np.load(GOLD_VPG)
    key = ('anil' if anil else 'maml') + ('2' if steps == 2 else '')
    tasks = int(g['tasks'])
    policy = (DiagNormalPolicyANIL(2, 2, 100) if anil else DiagNormalPolicy(2, 2, activation='tanh')).to(kdev)
    policy.load_flat_parameters(torch.from_numpy(g[key + '_theta0']).float().to(kdev))
    maml = MAML(policy, lr=VPG_CFG['inner_lr'])
    baseline = LinearValue(2, VPG_CFG['value_reg'])
    params = dict(VPG_CFG, adapt_steps=steps, adapt_batch_size=int(g['episodes']))
    data = make_replays(tasks, int(g['episodes']), int(g['horizon']), seed=int(g['seed']))
    data2 = make_replays(tasks, int(g['episodes']), int(g['horizon']), seed=int(g['seed2']))

    class StubTask:
        def __init__(self, queue):
            self.queue = list(queue)

        def run(self, learner, episodes=None, render=False):
            return self.queue.pop(0)

    total = 0.0
    for t, (sup, qry) in enumerate(data):
        learner = maml.clone()
        queue = [sup, qry] if steps == 1 else [sup, data2[t][0], qry]
        loss, _rew, _suc = xrl.fast_adapt_vpg(StubTask(queue), learner, baseline, params, anil=anil)
        assert rel(learner.module.flat_parameters(), torch.from_numpy(g[key + '_adapted'][t])) < 1e-5
        assert abs(float(loss) - float(g[key + '_valid_loss'][t])) < 2e-5 * abs(float(g[key + '_valid_loss'][t])) + 1e-7
        total = total + loss
    (total / tasks).backward()
    grad = torch.cat([p.grad.reshape(-1) for p in policy.parameters()])
    assert rel(grad, torch.from_numpy(g[key + '_grad'])) < 2e-4
    if steps == 1 and not anil:
        # value-only entry point: vpg_a2c_loss of the un-adapted policy on a support replay == the oracle
        sup64 = make_replays(tasks, int(g['episodes']), int(g['horizon']), seed=int(g['seed']), dtype=torch.float64)[0][0]
        th64 = [p.requires_grad_() for p in _unflat64(g[key + '_theta0'])]
        lp = ro.log_prob(th64, sup64['states'], sup64['actions'])
        ref = ch.a2c_policy_loss(lp, ro.compute_advantages(sup64, VPG_CFG['tau'], VPG_CFG['gamma'], VPG_CFG['value_reg']))
        got = xrl.vpg_a2c_loss(data[0][0], policy, baseline, VPG_CFG['gamma'], VPG_CFG['tau'])
        assert abs(float(got) - float(ref)) < 2e-5 * abs(float(ref)) + 1e-7


def _unflat64(vec):
    like = ro.init_policy(dtype=torch.float64)
    out, o = [], 0
    for p in like:
        out.append(torch.as_tensor(vec[o:o + p.numel()]).view_as(p).clone())
        o += p.numel()
    return out
