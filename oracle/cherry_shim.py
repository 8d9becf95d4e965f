"""Restatement of the cherry-rl functions the reference's RL hot path calls (TEST INFRASTRUCTURE).

cherry-rl is listed unpinned in the reference's requirements.txt:5 and is not installed here (no network).  The
reference's call sites: core_functions/rl.py:1-3 (imports), :96 td.discount, :105 pg.generalized_advantage,
:355 ch.normalize, :358 a2c.policy_loss, :417 trpo.hessian_vector_product, :418 trpo.conjugate_gradient,
:469 trpo.policy_loss; rl/maml_trpo.py:85 ch.models.robotics.LinearValue.  What follows restates the published
algorithms of cherry >= 0.1.0 (from memory -- PARITY UNPINNED: the reference ships no fixtures
for this path); ``install()`` registers them under the module names the reference imports so that
core_functions/rl.py and core_functions/policies.py load unmodified (oracle/rl_ref_loader.py).
"""
import sys
import types

import torch
from torch import nn


def _col(x):
    return x.view(-1, 1) if x.dim() == 1 else x


def discount(gamma, rewards, dones, bootstrap=0.0):
    """cherry.td.discount: reverse scan R <- r_t + gamma * R * (1 - d_t); output [T, 1]."""
    rewards, dones = _col(rewards), _col(dones)
    R = torch.zeros_like(rewards[0]) + bootstrap
    out = torch.zeros_like(rewards)
    for t in reversed(range(out.size(0))):
        R = R * (1.0 - dones[t])
        R = rewards[t] + gamma * R
        out[t] = out[t] + R[0]
    return out


def temporal_difference(gamma, rewards, dones, values, next_values):
    rewards, dones, values, next_values = _col(rewards), _col(dones), _col(values), _col(next_values)
    return rewards + gamma * (1.0 - dones) * next_values - values


def generalized_advantage(gamma, tau, rewards, dones, values, next_value):
    """cherry.pg.generalized_advantage."""
    rewards, dones, values = _col(rewards), _col(dones), _col(values)
    next_values = torch.cat((values[1:], _col(next_value)), dim=0)
    td = temporal_difference(gamma, rewards, dones, values, next_values)
    return discount(tau * gamma, td, dones)


def normalize(tensor, epsilon=1e-8):
    """cherry.normalize: (x - mean) / (unbiased std + eps); identity for <= 1 element."""
    if tensor.numel() <= 1:
        return tensor
    return (tensor - tensor.mean()) / (tensor.std() + epsilon)


def a2c_policy_loss(log_probs, advantages):
    return -torch.mean(log_probs * advantages)


def trpo_policy_loss(new_log_probs, old_log_probs, advantages):
    return -torch.mean(torch.exp(new_log_probs - old_log_probs) * advantages)


def ppo_policy_loss(new_log_probs, old_log_probs, advantages, clip=0.1):
    ratios = torch.exp(new_log_probs - old_log_probs)
    return -torch.mean(torch.min(ratios * advantages, ratios.clamp(1.0 - clip, 1.0 + clip) * advantages))


def hessian_vector_product(loss, parameters, damping=1e-5):
    """cherry.algorithms.trpo.hessian_vector_product: v -> flat(grad(grad(loss) . v)) + damping * v."""
    parameters = list(parameters)
    grad = torch.autograd.grad(loss, parameters, create_graph=True, retain_graph=True)
    flat = torch.nn.utils.parameters_to_vector(grad)

    def hvp(v, retain_graph=True):
        hv = torch.autograd.grad(torch.dot(flat, v), parameters, retain_graph=retain_graph)
        return torch.nn.utils.parameters_to_vector(hv) + damping * v
    return hvp


def conjugate_gradient(Ax, b, num_iterations=10, tol=1e-10, eps=1e-8):
    """cherry.algorithms.trpo.conjugate_gradient: textbook CG from x = 0."""
    x = torch.zeros_like(b)
    r = b
    p = r
    r_dot_old = torch.dot(r, r)
    for _ in range(num_iterations):
        Ap = Ax(p)
        alpha = r_dot_old / (torch.dot(p, Ap) + eps)
        x = x + alpha * p
        r = r - alpha * Ap
        r_dot_new = torch.dot(r, r)
        p = r + (r_dot_new / r_dot_old) * p
        r_dot_old = r_dot_new
        if r_dot_new.item() < tol:
            break
    return x


class LinearValue(nn.Module):
    """cherry.models.robotics.LinearValue: ridge regression on [s, s^2, t, t^2, t^3, 1], t = arange(N) / 100."""

    def __init__(self, input_size, reg=1e-5):
        super().__init__()
        self.linear = nn.Linear(2 * input_size + 4, 1, bias=False)
        self.reg = reg

    def _features(self, states):
        n = states.size(0)
        ones = torch.ones(n, 1, dtype=states.dtype, device=states.device)
        al = torch.arange(n, dtype=states.dtype, device=states.device).view(-1, 1) / 100.0
        return torch.cat([states, states ** 2, al, al ** 2, al ** 3, ones], dim=1)

    def fit(self, states, returns):
        f = self._features(states)
        A = f.t() @ f + self.reg * torch.eye(f.size(1), dtype=f.dtype, device=f.device)
        b = f.t() @ returns
        coeffs = torch.linalg.lstsq(A, b).solution
        self.linear.weight.data = coeffs.t().to(self.linear.weight.dtype)

    def forward(self, states):
        return self.linear(self._features(states))


class Replay:
    """Stand-in for cherry.ExperienceReplay as the reference reads it (core_functions/rl.py:49-56)."""

    def __init__(self, states, actions, rewards, dones, next_states):
        self._s, self._a, self._r, self._d, self._n = states, actions, _col(rewards), _col(dones), next_states

    def state(self):
        return self._s

    def action(self):
        return self._a

    def reward(self):
        return self._r

    def done(self):
        return self._d

    def next_state(self):
        return self._n


def install():
    """Registers the restatement under the module names core_functions/rl.py imports."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    a2c = mod('cherry.algorithms.a2c', policy_loss=a2c_policy_loss)
    trpo = mod('cherry.algorithms.trpo', policy_loss=trpo_policy_loss, hessian_vector_product=hessian_vector_product,
               conjugate_gradient=conjugate_gradient)
    ppo = mod('cherry.algorithms.ppo', policy_loss=ppo_policy_loss)
    algorithms = mod('cherry.algorithms', a2c=a2c, trpo=trpo, ppo=ppo)
    td = mod('cherry.td', discount=discount, temporal_difference=temporal_difference)
    pg = mod('cherry.pg', generalized_advantage=generalized_advantage)
    robotics = mod('cherry.models.robotics', LinearValue=LinearValue)
    models = mod('cherry.models', robotics=robotics)
    mod('cherry', algorithms=algorithms, td=td, pg=pg, models=models, normalize=normalize, ExperienceReplay=Replay)
