"""Self-contained CPU restatement of the reference's MAML-TRPO policy path (config 5; TEST INFRASTRUCTURE).

Follows, function by function:
  core_functions/policies.py:30-56   DiagNormalPolicy (2 x 100 hidden units, `sigma` = log-std parameter, log_prob =
                                     per-dimension Normal log-density averaged over the action dimensions)
  core_functions/rl.py:95-110        compute_advantages (discounted returns, LinearValue fit, bootstraps, GAE)
  core_functions/rl.py:346-358       trpo_a2c_loss  (-mean(log_prob * normalised advantages))
  core_functions/rl.py:361-374       trpo_update    (inner step theta' = theta - lr * grad, first or second order)
  core_functions/rl.py:441-473       meta_surrogate_loss (re-adaptation on the stored support replays, KL(new || old),
                                     importance-weighted surrogate on the query replay; means over tasks)
  core_functions/rl.py:264-316       fast_adapt_ppo (PPO inner loop of MAML / ANIL - PPO, policies.py:70-126 for the ANIL policy)
  core_functions/rl.py:409-438       meta_optimize_trpo (gradient, Fisher-vector product of the KL, conjugate gradient,
                                     step scaling by max_kl, backtracking line search)
plus the cherry functions restated in oracle/cherry_shim.py.  Parameters are a flat list in
``DiagNormalPolicy.parameters()`` order -- a module's own parameters precede its children's: sigma [out], then
W1 [100, in], b1, W2 [100, 100], b2, W3 [out, 100], b3.

PARITY UNPINNED by the reference (no tests or fixtures there).  tests/golden/make_golden_rl.py checks this file
against the reference's own rl.py / policies.py (run in place through oracle/rl_ref_loader.py) in float64 and commits
the vectors; no CUDA path for config 5 exists yet -- this is the checker it will be built against.
"""
import math

import torch
from torch.distributions import Normal
from torch.distributions.kl import kl_divergence

from . import cherry_shim as ch

EPSILON = 1e-6


def init_policy(in_size=2, out_size=2, hidden=(100, 100), seed=42, dtype=torch.float32):
    """Parameters initialised like DiagNormalPolicy under torch.manual_seed(seed) (policies.py:17-21,41-47)."""
    torch.manual_seed(seed)
    sizes = [in_size] + list(hidden) + [out_size]
    params = [torch.full((out_size,), math.log(1.0), dtype=dtype)]
    for i, o in zip(sizes[:-1], sizes[1:]):
        lin = torch.nn.Linear(i, o)
        torch.nn.init.xavier_uniform_(lin.weight)
        lin.bias.data.zero_()
        params += [lin.weight.detach().to(dtype), lin.bias.detach().to(dtype)]
    return params


def policy_mean(params, states, activation=torch.tanh):
    h = states
    n_lin = (len(params) - 1) // 2
    for l in range(n_lin):
        h = h @ params[1 + 2 * l].t() + params[2 + 2 * l]
        if l < n_lin - 1:
            h = activation(h)
    return h


def density(params, states):
    scale = torch.exp(torch.clamp(params[0], min=math.log(EPSILON)))
    return Normal(loc=policy_mean(params, states), scale=scale)


def log_prob(params, states, actions):
    return density(params, states).log_prob(actions).mean(dim=1, keepdim=True)


def compute_advantages(rep, tau, gamma, value_reg):
    """rl.py:95-110 with a freshly fitted LinearValue(state_size, reg=value_reg) (update_vf=True)."""
    s, r, d, ns = rep['states'], rep['rewards'], rep['dones'], rep['next_states']
    returns = ch.discount(gamma, r, d)
    baseline = ch.LinearValue(s.size(1), value_reg).to(s.dtype)
    baseline.fit(s, returns)
    with torch.no_grad():
        values, next_values = baseline(s), baseline(ns)
    bootstraps = values * (1.0 - d) + next_values * d
    return ch.generalized_advantage(gamma, tau, r, d, bootstraps, torch.zeros(1, dtype=s.dtype))


def a2c_loss(params, rep, tau, gamma, value_reg):
    lp = log_prob(params, rep['states'], rep['actions'])
    adv = ch.normalize(compute_advantages(rep, tau, gamma, value_reg)).detach()
    return ch.a2c_policy_loss(lp, adv)


def trpo_update(params, rep, inner_lr, tau, gamma, value_reg, first_order=False):
    loss = a2c_loss(params, rep, tau, gamma, value_reg)
    grads = torch.autograd.grad(loss, params, retain_graph=not first_order, create_graph=not first_order)
    return [p - inner_lr * g for p, g in zip(params, grads)]


def policy_mean_anil(params, states, activation=torch.tanh, body_no_grad=False):
    """DiagNormalPolicyANIL.forward_pass (policies.py:98-104): body = all layers but the last, head = the last Linear;
    with ``features_no_grad`` the body runs under no_grad."""
    n_lin = (len(params) - 1) // 2
    def body(h):
        for l in range(n_lin - 1):
            h = activation(h @ params[1 + 2 * l].t() + params[2 + 2 * l])
        return h
    if body_no_grad:
        with torch.no_grad():
            feats = body(states)
    else:
        feats = body(states)
    return feats @ params[2 * n_lin - 1].t() + params[2 * n_lin]


def fast_adapt_ppo(params, support, query, cfg, anil=False, activation=torch.tanh):
    """core_functions/rl.py:264-316 on fixed replays (the environment rollouts replaced by ``support`` -- one replay or a
    list, one per adaptation step -- and ``query``): per step ``ppo_epochs`` inner steps of learn2learn ``adapt`` (second
    order) on the clipped PPO objective with the old log-probabilities fixed at the learner the step started from
    (no_grad), then the PPO objective of the adapted learner on the query replay against its own detached
    log-probabilities.  ANIL: the body runs under no_grad during the inner loop (``turn_off_body_grads``) and
    ``allow_unused`` leaves its parameters un-adapted.  Returns (validation loss with graph, adapted parameter list)."""
    def lp(ps, rep, body_no_grad=False):
        mean = policy_mean_anil(ps, rep['states'], activation, body_no_grad) if anil else policy_mean(ps, rep['states'], activation)
        scale = torch.exp(torch.clamp(ps[0], min=math.log(EPSILON)))
        return Normal(loc=mean, scale=scale).log_prob(rep['actions']).mean(dim=1, keepdim=True)
    new = list(params)
    for rep in (support if isinstance(support, (list, tuple)) else [support]):
        adv = ch.normalize(compute_advantages(rep, cfg['tau'], cfg['gamma'], cfg['value_reg'])).detach()
        with torch.no_grad():
            old_lp = lp(new, rep, anil)
        for _epoch in range(cfg['ppo_epochs']):
            loss = ch.ppo_policy_loss(lp(new, rep, anil), old_lp, adv, clip=cfg['ppo_clip_ratio'])
            grads = torch.autograd.grad(loss, new, retain_graph=True, create_graph=True, allow_unused=anil)
            new = [p if g is None else p - cfg['inner_lr'] * g for p, g in zip(new, grads)]
    adv_q = ch.normalize(compute_advantages(query, cfg['tau'], cfg['gamma'], cfg['value_reg'])).detach()
    with torch.no_grad():
        old_q = lp(new, query)
    valid = ch.ppo_policy_loss(lp(new, query), old_q, adv_q, clip=cfg['ppo_clip_ratio'])
    return valid, new


def fast_adapt_vpg(params, support, query, cfg, anil=False, first_order=False, activation=torch.tanh):
    """core_functions/rl.py:208-254 on fixed replays: ``vpg_a2c_loss`` is ``a2c.policy_loss`` of the learner's
    log-probabilities against the GAE advantages of a freshly fitted baseline -- NOT normalised (unlike
    ``trpo_a2c_loss``, rl.py:355) -- ``learner.adapt(loss, first_order, allow_unused=anil)`` is one MAML step per support
    replay (``support``: one replay or a list, one per adaptation step), and the validation loss is the same loss of the
    adapted learner on the query replay.  ANIL: the body runs under no_grad during the inner losses
    (``turn_off_body_grads``) and its parameters stay un-adapted; the query loss sees the whole network.
    Returns (validation loss with graph, adapted parameter list)."""
    def lp(ps, rep, body_no_grad=False):
        mean = policy_mean_anil(ps, rep['states'], activation, body_no_grad) if anil else policy_mean(ps, rep['states'], activation)
        scale = torch.exp(torch.clamp(ps[0], min=math.log(EPSILON)))
        return Normal(loc=mean, scale=scale).log_prob(rep['actions']).mean(dim=1, keepdim=True)
    new = list(params)
    for rep in (support if isinstance(support, (list, tuple)) else [support]):
        adv = compute_advantages(rep, cfg['tau'], cfg['gamma'], cfg['value_reg'])
        loss = ch.a2c_policy_loss(lp(new, rep, anil), adv)
        grads = torch.autograd.grad(loss, new, retain_graph=not first_order, create_graph=not first_order, allow_unused=anil)
        new = [p if g is None else p - cfg['inner_lr'] * g for p, g in zip(new, grads)]
    adv_q = compute_advantages(query, cfg['tau'], cfg['gamma'], cfg['value_reg'])
    return ch.a2c_policy_loss(lp(new, query), adv_q), new


def meta_surrogate_loss(params, iter_replays, iter_old_params, cfg):
    mean_loss, mean_kl = 0.0, 0.0
    for replays, old_params in zip(iter_replays, iter_old_params):
        new = params
        for rep in replays[:-1]:
            new = trpo_update(new, rep, cfg['inner_lr'], cfg['tau'], cfg['gamma'], cfg['value_reg'], first_order=False)
        valid = replays[-1]
        old_d, new_d = density(old_params, valid['states']), density(new, valid['states'])
        mean_kl = mean_kl + kl_divergence(new_d, old_d).mean()
        adv = ch.normalize(compute_advantages(valid, cfg['tau'], cfg['gamma'], cfg['value_reg'])).detach()
        old_lp = old_d.log_prob(valid['actions']).mean(dim=1, keepdim=True).detach()
        new_lp = new_d.log_prob(valid['actions']).mean(dim=1, keepdim=True)
        mean_loss = mean_loss + ch.trpo_policy_loss(new_lp, old_lp, adv)
    return mean_loss / len(iter_replays), mean_kl / len(iter_replays)


def meta_optimize_trpo(params, iter_replays, iter_old_params, cfg):
    """rl.py:409-438.  Returns (new parameters, diagnostics)."""
    params = [p.detach().clone().requires_grad_() for p in params]
    old_loss, old_kl = meta_surrogate_loss(params, iter_replays, iter_old_params, cfg)
    grad = torch.autograd.grad(old_loss, params, retain_graph=True)
    grad = torch.nn.utils.parameters_to_vector([g.detach() for g in grad])
    Fvp = ch.hessian_vector_product(old_kl, params)
    step = ch.conjugate_gradient(Fvp, grad)
    shs = 0.5 * torch.dot(step, Fvp(step))
    step = step / torch.sqrt(shs / cfg['max_kl'])
    old_loss = old_loss.detach()
    diag = {'old_loss': float(old_loss), 'old_kl': float(old_kl.detach()), 'grad': grad.clone(), 'step': step.clone(), 'ls_step': -1}
    sizes = [p.numel() for p in params]
    chunks = [c.view_as(p) for c, p in zip(torch.split(step, sizes), params)]
    out = [p.detach().clone() for p in params]
    for ls in range(cfg['ls_max_steps']):
        stepsize = cfg['backtrack_factor'] ** ls * cfg['outer_lr']
        cand = [(p.detach() - stepsize * u).requires_grad_() for p, u in zip(params, chunks)]
        new_loss, kl = meta_surrogate_loss(cand, iter_replays, iter_old_params, cfg)
        if new_loss < old_loss and kl < cfg['max_kl']:
            out = [c.detach() for c in cand]
            diag['ls_step'] = ls
            break
    return out, diag
