"""Mirror of the reference's ``core_functions/rl.py`` for the MAML-TRPO path (config 5): same function names,
argument order and side effects, running on libxmeta's kernels through ``rl_engine.TrpoEngine``.

  compute_advantages   rl.py:95-110     trpo_a2c_loss        rl.py:346-358     trpo_update   rl.py:361-374
  fast_adapt_trpo      rl.py:377-406    meta_surrogate_loss  rl.py:441-473     meta_optimize_trpo   rl.py:409-438
  fast_adapt_ppo       rl.py:264-316 (MAML-PPO and, with ``anil=True`` + ``DiagNormalPolicyANIL``, ANIL-PPO)

``episodes`` / replays are anything with cherry's ExperienceReplay accessors (``state() action() reward() done()
next_state()``) or dicts with the plural keys.  Losses are returned as detached device scalars: the second-order
information the reference carries in autograd graphs is recomputed inside ``meta_optimize_trpo`` from the stored
replays (exactly what the reference does: it re-adapts from ``iter_replays``, rl.py:449-454)."""
import torch

from ..rl_engine import KEYS, LOG_EPS, TrpoEngine

_ACCESSOR = {'states': 'state', 'actions': 'action', 'rewards': 'reward', 'dones': 'done', 'next_states': 'next_state'}
_engines = {}


def get_episode_values(episodes):
    """rl.py:49-56 (without the device move: tensors are copied into the engine's buffers)."""
    if isinstance(episodes, dict):
        return tuple(episodes[k] for k in KEYS)
    return tuple(getattr(episodes, _ACCESSOR[k])() for k in KEYS)


def _as_dict(episodes):
    return dict(zip(KEYS, get_episode_values(episodes)))


def _engine(policy, baseline, tasks, n, inner_lr, gamma, tau, device):
    key = (tasks, n, policy.input_size, policy.output_size, tuple(policy.hiddens), policy.activation, str(device))
    e = _engines.get(key)
    if e is None:
        e = TrpoEngine(tasks, n, policy.input_size, policy.output_size, tuple(policy.hiddens), policy.activation,
                       inner_lr, gamma, tau, baseline.reg, device=device)
        _engines[key] = e
    e.lr, e.gamma, e.tau, e.reg = float(inner_lr), float(gamma), float(tau), float(baseline.reg)
    return e


def _device_of(policy):
    return next(policy.parameters()).device


def compute_advantages(baseline, tau, gamma, rewards, dones, states, next_states, update_vf=True):
    """Un-normalised GAE advantages [N, 1] of one replay with the LinearValue baseline fitted to its returns
    (``update_vf=False`` -- re-using an earlier fit -- is not what the TRPO path does and is not supported)."""
    if not update_vf:
        raise NotImplementedError('compute_advantages(update_vf=False)')
    import ctypes
    from .. import _lib
    from ..engine import _p
    dev = states.device
    n, sd = states.shape[0], states.shape[1]
    f = lambda t: t.reshape(n, -1).float().contiguous()                                    # noqa: E731
    s, ns, r, d = f(states), f(next_states), f(rewards).reshape(n), f(dones).reshape(n)
    coef, adv = torch.empty(n, device=dev), torch.empty(n, device=dev)
    a = _lib.XmRlAdvArgs()
    a.replays, a.n, a.state_dim = 1, n, sd
    a.gamma, a.tau, a.reg, a.coef_scale = gamma, tau, baseline.reg, 1.0
    a.states, a.next_states, a.rewards, a.dones = _p(s), _p(ns), _p(r), _p(d)
    a.coef, a.advantages = _p(coef), _p(adv)
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == 'cuda' else 0
    _lib.check(_lib.load().xm_rl_advantages(ctypes.byref(a), stream), 'xm_rl_advantages')
    return adv.view(n, 1)


def trpo_a2c_loss(episodes, learner, baseline, gamma, tau, update_vf=True):
    """-mean(log_prob * normalised advantage) of ``learner`` on ``episodes`` (value only)."""
    rep = _as_dict(episodes)
    dev = _device_of(learner)
    e = _engine(learner, baseline, 1, rep['states'].shape[0], 0.0, gamma, tau, dev)
    e.load_replays([[rep, rep]])
    theta = learner.flat_parameters().to(dev)
    return e.a2c_loss(theta)[0]


def trpo_update(episodes, learner, baseline, inner_lr, gamma, tau, anil=False, first_order=False):
    """One inner step theta' = theta - inner_lr * grad(trpo_a2c_loss) written into ``learner`` (``maml_update`` mutates
    and returns the module it is given, rl.py:374)."""
    if anil:
        raise NotImplementedError('ANIL-TRPO: the reference\'s own path (rl/anil_trpo.py -> rl.py:382, `.module` on a bare '
                                  'policy) raises before any arithmetic -- there is no behaviour to mirror (DESIGN 10)')
    rep = _as_dict(episodes)
    dev = _device_of(learner)
    e = _engine(learner, baseline, 1, rep['states'].shape[0], inner_lr, gamma, tau, dev)
    e.load_replays([[rep, rep]])
    out = e.adapt(learner.flat_parameters().to(dev))
    return learner.load_flat_parameters(out[0])


def fast_adapt_trpo(task, learner, baseline, params, anil=False, first_order=False, render=False):
    """rl.py:377-406: ``task.run(policy, episodes=...)`` collects the replays (environment side, caller-provided)."""
    task_replay = []
    for _step in range(params['adapt_steps']):
        support_episodes = task.run(learner, episodes=params['adapt_batch_size'], render=render)
        task_replay.append(support_episodes)
        learner = trpo_update(support_episodes, learner, baseline, params['inner_lr'], params['gamma'], params['tau'],
                              anil=anil, first_order=first_order)
    query_episodes = task.run(learner, episodes=params['adapt_batch_size'])
    task_replay.append(query_episodes)
    valid_loss = trpo_a2c_loss(query_episodes, learner, baseline, params['gamma'], params['tau'], update_vf=False)
    query_rew = get_episode_values(query_episodes)[2].sum().item() / params['adapt_batch_size']
    return learner, valid_loss, task_replay, query_rew, 0.0


def _prepare(iter_replays, iter_policies, policy, baseline, params):
    if any(len(r) != 2 for r in iter_replays):
        raise NotImplementedError('one adaptation step (one support replay + the query replay per task)')
    dev = _device_of(policy)
    reps = [[_as_dict(s), _as_dict(q)] for s, q in iter_replays]
    n = reps[0][0]['states'].shape[0]
    e = _engine(policy, baseline, len(reps), n, params['inner_lr'], params['gamma'], params['tau'], dev)
    e.load_replays(reps)
    e.set_old_policies(torch.stack([p.flat_parameters() for p in iter_policies]).to(dev))
    return e, policy.flat_parameters().to(dev)


def meta_surrogate_loss(iter_replays, iter_policies, policy, baseline, params, anil):
    """(mean surrogate loss, mean KL(new || old)) over the tasks at the current ``policy`` parameters."""
    if anil:
        raise NotImplementedError('ANIL-TRPO: the reference\'s own path (rl/anil_trpo.py -> rl.py:382, `.module` on a bare '
                                  'policy) raises before any arithmetic -- there is no behaviour to mirror (DESIGN 10)')
    e, theta = _prepare(iter_replays, iter_policies, policy, baseline, params)
    return e.loss_and_kl(theta)


def meta_optimize_trpo(params, policy, baseline, iter_replays, iter_policies, anil=False):
    """One TRPO meta-step: CG direction from the second-order meta-gradient and the Fisher-vector product of the KL,
    step scaling by ``max_kl``, backtracking line search; updates ``policy`` in place like the reference."""
    if anil:
        raise NotImplementedError('ANIL-TRPO: the reference\'s own path (rl/anil_trpo.py -> rl.py:382, `.module` on a bare '
                                  'policy) raises before any arithmetic -- there is no behaviour to mirror (DESIGN 10)')
    e, theta = _prepare(iter_replays, iter_policies, policy, baseline, params)
    new, diag = e.meta_optimize(theta, params['max_kl'], params['ls_max_steps'], params['backtrack_factor'],
                                params['outer_lr'])
    policy.load_flat_parameters(new)
    return diag



class _PpoValidLoss(torch.autograd.Function):
    """The validation loss of ``fast_adapt_ppo`` as a function of the learner's (cloned, pre-adaptation) parameters:
    the value and its second-order gradient were computed by the kernels; ``backward`` hands the gradient to autograd,
    which carries it through the ``clone()`` edges into the master policy's ``.grad`` (rl/maml_ppo.py:129)."""

    @staticmethod
    def forward(ctx, value, grad_flat, *params):
        ctx.grad_flat, ctx.shapes = grad_flat, [p.shape for p in params]
        return value.clone()

    @staticmethod
    def backward(ctx, gout):
        outs, o = [], 0
        for shp in ctx.shapes:
            n = 1
            for d in shp:
                n *= d
            outs.append((gout * ctx.grad_flat[o:o + n]).view(shp))
            o += n
        return (None, None) + tuple(outs)


def fast_adapt_ppo(task, learner, baseline, params, anil=False, render=False):
    """rl.py:264-316.  ``learner`` = ``policy.clone()`` of a ``MAML``-wrapped ``DiagNormalPolicy`` /
    ``DiagNormalPolicyANIL``; ``task.run(learner, episodes=...)`` collects the replays (environment side,
    caller-provided), one support replay per adaptation step, each followed by ``ppo_epochs`` clipped-PPO steps against
    the log-probabilities of the policy the step started from.  Returns ``(valid_loss, query_rew,
    query_success_rate)``; ``valid_loss.backward()`` accumulates the second-order meta-gradient into the master policy
    like the reference's autograd graph does (the cotangent is carried back through the steps, last to first)."""
    policy = learner.module
    before = list(policy.parameters())                       # the clone's differentiable copies of the master
    dev = before[0].device
    theta = torch.cat([p.detach().reshape(-1).float() for p in before])
    epochs, clip = params['ppo_epochs'], params['ppo_clip_ratio']
    if anil:
        policy.turn_off_body_grads()
    supports, chains, e = [], [], None
    for _step in range(params.get('adapt_steps', 1)):
        support = _as_dict(task.run(learner, episodes=params['adapt_batch_size'], render=render))
        e = _engine(policy, baseline, 1, support['states'].shape[0], params['inner_lr'], params['gamma'], params['tau'], dev)
        e.load_replays([[support, support]])
        theta = e.ppo_adapt(theta, epochs, clip, anil)[0].clone()
        supports.append(support)
        chains.append(e.ppo_thetas.clone())                   # theta_0 .. theta_E of this step
        # re-bind the learner's parameters to the adapted values (what learner.adapt leaves behind) for the next rollouts
        o = 0
        for module in policy.modules():
            for name, p in list(module._parameters.items()):
                if p is not None:
                    module._parameters[name] = theta[o:o + p.numel()].view_as(p)
                    o += p.numel()
    if anil:
        policy.turn_on_body_grads()
    query_episodes = task.run(learner, episodes=params['adapt_batch_size'])
    query = _as_dict(query_episodes)
    e.load_replays([[supports[-1], query]])
    valid, grad = e.ppo_outer(epochs, clip, anil)             # back through the last step (its chain is still loaded)
    if len(supports) > 1:
        cur = e.ppo_task_grads
        nxt = e.pertask if cur.data_ptr() == e.bar.data_ptr() else e.bar
        for s in reversed(range(len(supports) - 1)):
            e.load_replays([[supports[s], query]])
            e.ppo_thetas.copy_(chains[s])
            e.ppo_set_old()
            cur = e.ppo_backprop(cur, nxt, epochs, clip, anil)
            nxt = e.pertask if cur.data_ptr() == e.bar.data_ptr() else e.bar
        grad = cur[0].clone()
    valid_loss = _PpoValidLoss.apply(valid[0], grad, *before)
    query_rew = query['rewards'].sum().item() / params['adapt_batch_size']
    return valid_loss, query_rew, 0.0


def vpg_a2c_loss(episodes, learner, baseline, gamma, tau, dice=False):
    """rl.py:208-226, value only: ``a2c.policy_loss`` of ``learner`` on ``episodes`` against the raw (un-normalised) GAE
    advantages of the baseline re-fitted on these episodes."""
    if dice:
        raise NotImplementedError('the DiCE objective is not on the built path')
    rep = _as_dict(episodes)
    dev = _device_of(learner)
    e = _engine(learner, baseline, 1, rep['states'].shape[0], 0.0, gamma, tau, dev)
    e.load_replays([[rep, rep]], normalize=False)
    return e.a2c_loss(learner.flat_parameters().to(dev))[0]


def fast_adapt_vpg(task, learner, baseline, params, anil=False, first_order=False, render=False):
    """rl.py:229-254.  ``learner`` = ``policy.clone()`` of a ``MAML``-wrapped ``DiagNormalPolicy`` /
    ``DiagNormalPolicyANIL``; ``task.run(learner, episodes=...)`` collects the replays (environment side,
    caller-provided), one support replay per adaptation step.  Returns ``(valid_loss, query_rew,
    query_success_rate)``; ``valid_loss.backward()`` accumulates the (second-order unless ``first_order``) meta-gradient
    into the master policy like the reference's autograd graph: bar_S = d valid / d theta_S on the query replay, then
    bar_s = bar_{s+1} - lr * M H(theta_s; support_s) (M bar_{s+1}) back through the steps."""
    policy = learner.module
    before = list(policy.parameters())                       # the clone's differentiable copies of the master
    dev = before[0].device
    thetas = [torch.cat([p.detach().reshape(-1).float() for p in before])]
    lr = getattr(learner, 'lr', params.get('inner_lr'))
    if anil:
        policy.turn_off_body_grads()
    supports, e = [], None
    for _step in range(params.get('adapt_steps', 1)):
        support = _as_dict(task.run(learner, episodes=params['adapt_batch_size'], render=render))
        supports.append(support)
        e = _engine(policy, baseline, 1, support['states'].shape[0], lr, params['gamma'], params['tau'], dev)
        e.load_replays([[support, support]], normalize=False)
        adapted = e.adapt(thetas[-1], head_only=1 if anil else 0)[0].clone()
        thetas.append(adapted)
        # re-bind the learner's parameters to the adapted values (what learner.adapt leaves behind) for the next rollouts
        o = 0
        for module in policy.modules():
            for name, p in list(module._parameters.items()):
                if p is not None:
                    module._parameters[name] = adapted[o:o + p.numel()].view_as(p)
                    o += p.numel()
    if anil:
        policy.turn_on_body_grads()
    query = _as_dict(task.run(learner, episodes=params['adapt_batch_size']))
    e.load_replays([[supports[-1], query]], normalize=False)
    P = e.P
    valid = e.a2c_loss(thetas[-1], stride=0, k=1)[0].clone()
    cur, nxt = e.a2c_grad(thetas[-1], 0, 1, e.bar), e.pertask
    if not first_order:
        for s in reversed(range(len(supports))):
            if s != len(supports) - 1:
                e.load_replays([[supports[s], query]], normalize=False)
            e.hvp(thetas[s], cur, P, nxt, head_only=3 if anil else 0)
            cur, nxt = nxt, cur
    valid_loss = _PpoValidLoss.apply(valid, cur[0].clone(), *before)
    query_rew = query['rewards'].sum().item() / params['adapt_batch_size']
    return valid_loss, query_rew, 0.0
