"""Task-batched MAML-TRPO meta-optimisation (config 5): host-side launch sequences over ``xm_rl_advantages`` /
``xm_rl_sweep`` (csrc/rl.cu).

Replaces, for a whole meta-batch of tasks at once, the reference's per-task Python loops of
``core_functions/rl.py``: ``trpo_update`` (:361-374), ``meta_surrogate_loss`` (:441-473) and ``meta_optimize_trpo``
(:409-438).  Every task carries its own adapted policy theta'_t = theta - lr * grad L_t(theta).

Second order without an autograd tape (same forward-over-reverse scheme as the vision engine):
  * meta-gradient:  sum_t (I - lr H_t) bar'_t, with bar'_t = d surrogate_t / d theta'_t and H_t v = the tangent of
    the support-gradient sweep in direction v  (``XM_RL_HVP``);
  * Fisher-vector product of the mean KL (``trpo.hessian_vector_product`` at rl.py:417): it is evaluated at the point
    where the re-adapted policy equals the stored one, so the KL gradient w.r.t. the policy outputs vanishes and the
    Hessian is exactly J^T F J (tests/test_rl_oracle.py proves the identity against double backward): per CG iteration
    theta'dot_t = (I - lr H_t) v, one ``XM_RL_FISHER`` sweep, and (I - lr H_t) applied once more -- no third derivative.
Advantages (returns, LinearValue fit, GAE, normalisation) depend on the replays only: computed once per
meta-optimisation instead of once per loss evaluation (the reference recomputes them 1 + line-search-steps times per
task with Python per-timestep loops).  Conjugate gradient and the line search are host-driven, as in the reference.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import (XM_ACT_RELU, XM_ACT_TANH, XM_RL_A2C, XM_RL_FISHER, XM_RL_FORWARD, XM_RL_GRAD, XM_RL_HVP,
                   XM_RL_SURROGATE, XmRlAdvArgs, XmRlSweepArgs)
from . import engine as _engine
from .engine import _p

LOG_EPS = math.log(1e-6)           # policies.py:14,51
KEYS = ('states', 'actions', 'rewards', 'dones', 'next_states')


def policy_num_params(in_dim, out_dim, hiddens):
    h1, h2 = hiddens
    return out_dim + h1 * in_dim + h1 + h2 * h1 + h2 + out_dim * h2 + out_dim


class TrpoEngine:
    """Static buffers + launch sequences for ``tasks`` tasks with one support and one query replay of ``n``
    transitions each (``adapt_steps = 1``, as in the reference's configuration)."""

    def __init__(self, tasks, n, state_dim=2, action_dim=2, hiddens=(100, 100), activation='tanh', inner_lr=0.1,
                 gamma=0.99, tau=1.0, value_reg=1e-5, device='cuda'):
        self.device = torch.device(device)
        _engine._require_cuda(self.device)
        self.lib = _lib.load()
        assert len(hiddens) == 2, 'two hidden layers (the reference default [100, 100])'
        self.tasks, self.n, self.sd, self.ad = int(tasks), int(n), int(state_dim), int(action_dim)
        self.hiddens = tuple(int(h) for h in hiddens)
        self.act = {'tanh': XM_ACT_TANH, 'relu': XM_ACT_RELU}[activation]
        self.lr, self.gamma, self.tau, self.reg = float(inner_lr), float(gamma), float(tau), float(value_reg)
        self.P = policy_num_params(self.sd, self.ad, self.hiddens)
        B, n, P = self.tasks, self.n, self.P
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)      # noqa: E731
        # replay k: 0 = support, 1 = query
        self.states, self.next_states = f32(2, B, n, self.sd), f32(2, B, n, self.sd)
        self.actions = f32(2, B, n, self.ad)
        self.rewards, self.dones = f32(2, B, n), f32(2, B, n)
        self.coef = f32(2, B, n)                   # per-sample loss weights from the normalised advantages
        self.mu_old, self.logstd_old = f32(B, n, self.ad), f32(B, self.ad)
        self.theta_prime, self.bar, self.tdot, self.pertask = f32(B, P), f32(B, P), f32(B, P), f32(B, P)
        self.task_loss, self.task_kl = f32(B), f32(B)
        probe = self._sweep_args(XM_RL_A2C, XM_RL_GRAD, 0)
        nbytes = int(self.lib.xm_rl_sweep_scratch_bytes(ctypes.byref(probe)))
        if nbytes < 0:
            raise _lib.XmetaError('xm_rl_sweep: unsupported policy shape')
        self.partial = torch.zeros(nbytes // 8 + 1, dtype=torch.float64, device=self.device)
        self.total_tasks = self.tasks

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == 'cuda' else 0

    # ---- inputs ---------------------------------------------------------------------------------------------------------
    def load_replays(self, replays, normalize=True):
        """replays: list (tasks) of [support, query] dicts with keys states / actions / rewards / dones / next_states
        (``exploring_meta_b200.synthetic.make_replays``) -- or objects with cherry's ExperienceReplay accessors.
        ``normalize=False``: loss weights from the raw GAE advantages (``vpg_a2c_loss``, rl.py:208-226)."""
        def field(rep, key):
            if isinstance(rep, dict):
                return rep[key]
            return getattr(rep, {'states': 'state', 'actions': 'action', 'rewards': 'reward', 'dones': 'done',
                                 'next_states': 'next_state'}[key])()
        for k in range(2):
            for key, dst in zip(KEYS, (self.states, self.actions, self.rewards, self.dones, self.next_states)):
                stacked = torch.stack([field(task[k], key).reshape(dst.shape[2:]).float() for task in replays])
                dst[k].copy_(stacked, non_blocking=True)
        self.prepare(normalize)

    def prepare(self, normalize=True):
        """Advantages of every replay -> per-sample loss weights (once per meta-optimisation):
        support: -adv / n  (a2c.policy_loss, rl.py:358);  query: -adv / (n * tasks)  (trpo.policy_loss + the mean over
        tasks, rl.py:469-472).  adv = the normalised GAE advantages (``ch.normalize``, rl.py:355) or, with
        ``normalize=False``, the raw ones (``vpg_a2c_loss`` does not normalise, rl.py:208-226)."""
        if not normalize and getattr(self, 'adv_raw', None) is None:
            self.adv_raw = torch.zeros(2, self.tasks, self.n, dtype=torch.float32, device=self.device)
        for k, scale in ((0, -1.0 / self.n), (1, -1.0 / (self.n * self.total_tasks))):
            a = XmRlAdvArgs()
            a.replays, a.n, a.state_dim = self.tasks, self.n, self.sd
            a.gamma, a.tau, a.reg, a.coef_scale = self.gamma, self.tau, self.reg, scale
            a.states, a.next_states = _p(self.states[k]), _p(self.next_states[k])
            a.rewards, a.dones, a.coef = _p(self.rewards[k]), _p(self.dones[k]), _p(self.coef[k])
            if not normalize:
                a.advantages = _p(self.adv_raw[k])
            _lib.check(self.lib.xm_rl_advantages(ctypes.byref(a), self._stream()), 'xm_rl_advantages')
            if not normalize:
                torch.mul(self.adv_raw[k], scale, out=self.coef[k])      # the kernel's raw-advantage output, rescaled

    # ---- sweeps -----------------------------------------------------------------------------------------------------------
    def _sweep_args(self, loss, what, k):
        a = XmRlSweepArgs()
        a.tasks, a.n, a.in_dim, a.out_dim = self.tasks, self.n, self.sd, self.ad
        a.h1, a.h2, a.activation, a.loss, a.what = self.hiddens[0], self.hiddens[1], self.act, loss, what
        a.states, a.actions, a.coef = _p(self.states[k]), _p(self.actions[k]), _p(self.coef[k])
        a.kl_scale = 1.0 / (self.n * self.ad * getattr(self, 'total_tasks', self.tasks))
        if getattr(self, 'partial', None) is not None:
            a.partial, a.partial_bytes = self.partial.data_ptr(), self.partial.numel() * 8
        return a

    def _launch(self, a):
        _lib.check(self.lib.xm_rl_sweep(ctypes.byref(a), self._stream()), 'xm_rl_sweep')

    def adapt(self, theta, out=None, stride=0, head_only=0):
        """theta'_t = theta_t - lr * grad_theta [ -mean(log_prob * adv) ] on the support replay: ``trpo_update``
        (rl.py:361-374) for every task.  ``theta`` [P] shared (stride 0) or [tasks, P].  ``head_only=1``: the ANIL
        policy (body frozen in the inner loop)."""
        out = self.theta_prime if out is None else out
        a = self._sweep_args(XM_RL_A2C, XM_RL_GRAD, 0)
        a.theta, a.theta_task_stride = _p(theta), stride
        a.out, a.out_task_stride = _p(out), self.P
        a.base, a.base_task_stride, a.scale = _p(theta), stride, -self.lr
        a.head_only = head_only
        self._launch(a)
        return out

    def a2c_loss(self, theta, stride=0, k=0):
        """-mean(log_prob * normalised advantage) per task on replay k (``trpo_a2c_loss``, rl.py:346-358), values only."""
        a = self._sweep_args(XM_RL_A2C, XM_RL_FORWARD, k)
        a.theta, a.theta_task_stride = _p(theta), stride
        a.task_loss = _p(self.task_loss)
        self._launch(a)
        return self.task_loss if k == 0 else self.task_loss * self.total_tasks

    def hvp(self, theta, v, v_stride, out, head_only=0):
        """out_t = v_t - lr * H_t(theta) v_t: the cotangent (or tangent -- H is symmetric) through the adaptation step
        (``head_only=3``: v_t - lr * M H_t (M v_t), the ANIL inner-loop graph holds the head / sigma only)."""
        a = self._sweep_args(XM_RL_A2C, XM_RL_HVP, 0)
        a.theta, a.theta_task_stride = _p(theta), 0
        a.theta_dot, a.theta_dot_task_stride = _p(v), v_stride
        a.out, a.out_task_stride = _p(out), self.P
        a.base, a.base_task_stride, a.scale = _p(v), v_stride, -self.lr
        a.head_only = head_only
        self._launch(a)
        return out

    def set_old_policies(self, old_theta):
        """Outputs of the stored (first-order adapted) policies on the query states: ``old_policy.density(states)``
        (rl.py:459), computed once."""
        a = self._sweep_args(XM_RL_A2C, XM_RL_FORWARD, 1)
        a.theta, a.theta_task_stride = _p(old_theta), self.P
        a.mu_out = _p(self.mu_old)
        self._launch(a)
        self.logstd_old.copy_(torch.clamp(old_theta[:, :self.ad], min=LOG_EPS))

    def surrogate(self, theta_prime, grad_out=None):
        """Per-task surrogate loss / KL of the adapted policies on the query replay (already divided by the number of
        tasks); with ``grad_out`` also d/d theta'_t.  Returns (task_loss [tasks], task_kl [tasks])."""
        a = self._sweep_args(XM_RL_SURROGATE, XM_RL_FORWARD if grad_out is None else XM_RL_GRAD, 1)
        a.theta, a.theta_task_stride = _p(theta_prime), self.P
        a.mu_old, a.logstd_old = _p(self.mu_old), _p(self.logstd_old)
        a.task_loss, a.task_kl = _p(self.task_loss), _p(self.task_kl)
        if grad_out is not None:
            a.out, a.out_task_stride, a.scale = _p(grad_out), self.P, 1.0
        self._launch(a)
        return self.task_loss, self.task_kl

    def fisher(self, theta_prime, theta_dot, out):
        a = self._sweep_args(XM_RL_FISHER, XM_RL_GRAD, 1)
        a.theta, a.theta_task_stride = _p(theta_prime), self.P
        a.theta_dot, a.theta_dot_task_stride = _p(theta_dot), self.P
        a.logstd_old = _p(self.logstd_old)
        a.out, a.out_task_stride, a.scale = _p(out), self.P, 1.0
        self._launch(a)
        return out

    # ---- MAML / ANIL - PPO inner loop (core_functions/rl.py:264-316) ---------------------------------------------------------
    def _ppo_args(self, what, e, clip):
        a = self._sweep_args(XM_RL_SURROGATE, what, 0)
        a.theta, a.theta_task_stride = _p(self.ppo_thetas[e]), self.P
        a.mu_old, a.logstd_old, a.clip = _p(self.mu_old_s), _p(self.logstd_old_s), clip
        return a

    def ppo_adapt(self, theta, epochs, clip, anil=False):
        """The inner loop of ``fast_adapt_ppo`` (rl.py:268-293) for every task: ``epochs`` steps
        theta_{e+1} = theta_e - lr * M grad L_ppo(theta_e) on the support replay, old log-probabilities fixed at theta_0
        (no_grad, :281-282), M = identity (MAML) or the head / sigma mask (ANIL: the body runs under no_grad and
        ``allow_unused`` skips its update).  Returns the adapted parameters [tasks, P]."""
        B, P = self.tasks, self.P
        if getattr(self, 'ppo_thetas', None) is None or self.ppo_thetas.shape[0] != epochs + 1:
            self.ppo_thetas = torch.zeros(epochs + 1, B, P, dtype=torch.float32, device=self.device)
            self.mu_old_s = torch.zeros(B, self.n, self.ad, dtype=torch.float32, device=self.device)
            self.logstd_old_s = torch.zeros(B, self.ad, dtype=torch.float32, device=self.device)
        th = self.ppo_thetas
        th[0].copy_(theta.expand(B, P) if theta.dim() == 1 else theta)
        self.ppo_set_old()
        for e in range(epochs):
            a = self._ppo_args(XM_RL_GRAD, e, clip)
            a.out, a.out_task_stride = _p(th[e + 1]), P
            a.base, a.base_task_stride, a.scale = _p(th[e]), P, -self.lr
            a.head_only = 1 if anil else 0
            self._launch(a)
        return th[epochs]

    def ppo_set_old(self):
        """Outputs of the un-adapted policies ``ppo_thetas[0]`` on the loaded support states: the fixed old log-probabilities
        of an adaptation step (no_grad, rl.py:281-282)."""
        th0, P = self.ppo_thetas[0], self.P
        a = self._sweep_args(XM_RL_A2C, XM_RL_FORWARD, 0)
        a.theta, a.theta_task_stride, a.mu_out = _p(th0), P, _p(self.mu_old_s)
        self._launch(a)
        self.logstd_old_s.copy_(torch.clamp(th0[:, :self.ad], min=LOG_EPS))

    def ppo_backprop(self, cur, nxt, epochs, clip, anil=False):
        """Carries the cotangent ``cur`` [tasks, P] of theta_E back to theta_0 of ONE adaptation step (``ppo_thetas``,
        old outputs and support replay of that step loaded): bar_e = bar_{e+1} - lr * M H_e (M bar_{e+1}).  Returns the
        buffer holding bar_0 (``cur`` and ``nxt`` are used alternately)."""
        P = self.P
        for e in reversed(range(epochs)):
            a = self._ppo_args(XM_RL_HVP, e, clip)
            a.theta_dot, a.theta_dot_task_stride = _p(cur), P
            a.out, a.out_task_stride = _p(nxt), P
            a.base, a.base_task_stride, a.scale = _p(cur), P, -self.lr
            a.head_only = 3 if anil else 0          # the inner-loop graph holds the head / sigma only (body under no_grad)
            self._launch(a)
            cur, nxt = nxt, cur
        return cur

    def ppo_outer(self, epochs, clip, anil=False):
        """Validation loss of the adapted policies on the query replay (rl.py:297-309: the PPO objective against the
        adapted policy's OWN detached log-probabilities -- the ratio is identically 1, so the value is sum(coef) and the
        gradient is the a2c gradient at theta_E) and its second-order gradient w.r.t. theta_0:
        bar_e = bar_{e+1} - lr * M H_e (M bar_{e+1}), H_e v = the tangent of the PPO gradient sweep at theta_e.
        Returns (valid loss per task (already / total tasks), sum over tasks of the gradient [P])."""
        P = self.P
        a = self._sweep_args(XM_RL_A2C, XM_RL_GRAD, 1)
        a.theta, a.theta_task_stride = _p(self.ppo_thetas[epochs]), P
        a.out, a.out_task_stride, a.scale = _p(self.bar), P, 1.0
        self._launch(a)
        valid_loss = self.coef[1].sum(dim=1)
        cur = self.ppo_backprop(self.bar, self.pertask, epochs, clip, anil)
        self.ppo_task_grads = cur
        return valid_loss, self._sum_tasks(cur)

    def ppo_meta_gradient(self, theta, epochs, clip, anil=False):
        """``fast_adapt_ppo`` for every task + the gradient of the mean validation loss (rl/maml_ppo.py:103-129,
        rl/anil_ppo.py:106-130).  Returns (valid loss per task, gradient [P], adapted parameters [tasks, P])."""
        adapted = self.ppo_adapt(theta, epochs, clip, anil)
        valid, grad = self.ppo_outer(epochs, clip, anil)
        return valid, grad, adapted

    # ---- MAML / ANIL - VPG (core_functions/rl.py:208-254) ---------------------------------------------------------------------
    def vpg_meta_gradient(self, theta, anil=False, first_order=False):
        """``fast_adapt_vpg`` for every task (one adaptation step; replays loaded with ``normalize=False``) + the gradient
        of the mean validation loss w.r.t. the shared initial parameters ``theta`` [P]:
        theta'_t = theta - lr * M grad L_a2c(theta; support_t)            (``learner.adapt``, rl.py:241)
        valid_t  = L_a2c(theta'_t; query_t)                                (rl.py:250)
        grad     = sum_t [ bar_t - lr * M H_t(theta) (M bar_t) ],  bar_t = d valid_t / d theta'_t   (``first_order``: bar_t)
        M = identity (MAML) or the head / sigma mask (ANIL).  Returns (valid loss per task (already / total tasks),
        gradient [P], adapted parameters [tasks, P])."""
        P = self.P
        adapted = self.adapt(theta, head_only=1 if anil else 0)
        valid = self.a2c_loss(adapted, stride=P, k=1).clone() / self.total_tasks
        self.a2c_grad(adapted, P, 1, self.bar)
        if first_order:
            return valid, self._sum_tasks(self.bar), adapted
        self.hvp(theta, self.bar, P, self.pertask, head_only=3 if anil else 0)
        return valid, self._sum_tasks(self.pertask), adapted

    def a2c_grad(self, theta, stride, k, out):
        """out_t = d/d theta_t of the a2c loss of replay k (weights ``coef[k]``) at ``theta`` ([P] shared, stride 0, or
        [tasks, P])."""
        a = self._sweep_args(XM_RL_A2C, XM_RL_GRAD, k)
        a.theta, a.theta_task_stride = _p(theta), stride
        a.out, a.out_task_stride, a.scale = _p(out), self.P, 1.0
        self._launch(a)
        return out

    def _sum_tasks(self, per_task):
        out = torch.empty(self.P, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.xm_accumulate_tasks(_p(per_task), self.P, self.tasks, self.P, _p(out), 0, self._stream()),
                   'xm_accumulate_tasks')
        return out

    # ---- the quantities meta_optimize_trpo needs ------------------------------------------------------------------------------
    def loss_and_kl(self, theta):
        """``meta_surrogate_loss`` (rl.py:441-473) at ``theta``: (mean surrogate loss, mean KL) as device scalars."""
        self.adapt(theta)
        loss, kl = self.surrogate(self.theta_prime)
        return loss.sum(), kl.sum()

    def gradient(self, theta):
        """(loss, kl, d loss / d theta) with the second-order term through the adaptation step (rl.py:413-416)."""
        self.adapt(theta)
        loss, kl = self.surrogate(self.theta_prime, grad_out=self.bar)
        loss, kl = loss.sum(), kl.sum()
        self.hvp(theta, self.bar, self.P, self.pertask)
        return loss, kl, self._sum_tasks(self.pertask)

    def fvp(self, theta, v, damping=1e-5):
        """Fisher-vector product of the mean KL at theta (``trpo.hessian_vector_product``, rl.py:417); requires
        ``theta_prime`` = adapt(theta) from a preceding ``gradient`` call."""
        self.hvp(theta, v, 0, self.tdot)                       # tangent of theta'_t
        self.fisher(self.theta_prime, self.tdot, self.bar)     # F J v per task, pulled back to theta'_t
        self.hvp(theta, self.bar, self.P, self.pertask)        # ... and through the adaptation step
        return self._sum_tasks(self.pertask) + damping * v

    def meta_optimize(self, theta, max_kl=0.01, ls_max_steps=15, backtrack_factor=0.5, outer_lr=1.0, cg_iters=10,
                      damping=1e-5, cg_tol=1e-10, cg_eps=1e-8):
        """``meta_optimize_trpo`` (rl.py:409-438) on the loaded replays / stored old policies.  ``theta`` [P] is not
        modified; returns (new theta, diagnostics)."""
        old_loss, old_kl, grad = self.gradient(theta)
        # cherry.algorithms.trpo.conjugate_gradient(Fvp, grad)
        x = torch.zeros_like(grad)
        r, p = grad.clone(), grad.clone()
        r_dot_old = torch.dot(r, r)
        for _ in range(cg_iters):
            Ap = self.fvp(theta, p, damping)
            alpha = r_dot_old / (torch.dot(p, Ap) + cg_eps)
            x = x + alpha * p
            r = r - alpha * Ap
            r_dot_new = torch.dot(r, r)
            p = r + (r_dot_new / r_dot_old) * p
            r_dot_old = r_dot_new
            if r_dot_new.item() < cg_tol:
                break
        shs = 0.5 * torch.dot(x, self.fvp(theta, x, damping))
        step = x / torch.sqrt(shs / max_kl)
        old_loss_h = float(old_loss)
        diag = {'old_loss': old_loss_h, 'old_kl': float(old_kl), 'grad': grad, 'step': step, 'ls_step': -1}
        new_theta = theta.clone()
        for ls in range(ls_max_steps):
            stepsize = backtrack_factor ** ls * outer_lr
            cand = theta - stepsize * step
            new_loss, kl = self.loss_and_kl(cand)
            new_loss, kl = float(new_loss), float(kl)
            if new_loss < old_loss_h and kl < max_kl:
                new_theta = cand
                diag.update(ls_step=ls, new_loss=new_loss, new_kl=kl)
                break
        return new_theta, diag
