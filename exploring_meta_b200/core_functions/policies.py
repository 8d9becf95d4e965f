"""Mirror of the reference's ``core_functions/policies.py`` for the policy-MLP hot path (config 5).

``DiagNormalPolicy`` keeps the reference's constructor, attribute names (``mean`` Sequential, ``sigma``), parameter
order (``sigma`` first: a module's own parameters precede its children's) and ``state_dict`` keys, so checkpoints are
interchangeable.  The adaptation / meta-optimisation arithmetic does not run through these modules: the functions of
``core_functions/rl.py`` read the flat parameter vector and launch libxmeta's kernels.  ``density`` / ``log_prob`` /
``forward`` are the environment-interaction side (sampling actions during rollouts, out of scope: SURVEY 8) and stay
plain torch calls on whatever device the module lives on."""
import math

import torch
from torch import nn
from torch.distributions import Normal

EPSILON = 1e-6


def linear_init(module):
    """policies.py:17-21"""
    if isinstance(module, nn.Linear):
        nn.init.xavier_uniform_(module.weight)
        module.bias.data.zero_()
    return module


class DiagNormalPolicy(nn.Module):
    """policies.py:30-67.  ``activation`` defaults to 'relu' as in the reference (its MAML-TRPO driver forgets to pass
    'tanh', rl/maml_trpo.py:86); both run on the kernels."""

    def __init__(self, input_size, output_size, hiddens=None, activation='relu'):
        super().__init__()
        if hiddens is None:
            hiddens = [100, 100]
        self.activation = activation
        act = {'relu': nn.ReLU, 'tanh': nn.Tanh}[activation]
        layers = [linear_init(nn.Linear(input_size, hiddens[0])), act()]
        for i, o in zip(hiddens[:-1], hiddens[1:]):
            layers.append(linear_init(nn.Linear(i, o)))
            layers.append(act())
        layers.append(linear_init(nn.Linear(hiddens[-1], output_size)))
        self.mean = nn.Sequential(*layers)
        self.sigma = nn.Parameter(torch.Tensor(output_size))
        self.sigma.data.fill_(math.log(1))
        self.input_size, self.output_size, self.hiddens = input_size, output_size, list(hiddens)

    def density(self, state):
        loc = self.mean(state)
        scale = torch.exp(torch.clamp(self.sigma, min=math.log(EPSILON)))
        return Normal(loc=loc, scale=scale)

    def log_prob(self, state, action):
        return self.density(state).log_prob(action).mean(dim=1, keepdim=True)

    def forward(self, state):
        return self.density(state).sample()

    def get_representation(self, x, layer=-1):
        modules = list(self.mean.modules())
        for layer_i in modules[1:layer]:
            x = layer_i(x)
        return x

    # ---- flat parameter vector in parameters() order (what the kernels read) ------------------------------------------
    def flat_parameters(self):
        return torch.cat([p.detach().reshape(-1).float() for p in self.parameters()])

    def load_flat_parameters(self, flat):
        o = 0
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(flat[o:o + p.numel()].view_as(p))
                o += p.numel()
        return self


class DiagNormalPolicyANIL(nn.Module):
    """policies.py:70-126: the same network split into ``body`` (all layers but the last) and ``head``; during the inner
    loop the body runs under no_grad (``turn_off_body_grads``), so only ``sigma`` and the head adapt.  Same flat
    parameter layout as ``DiagNormalPolicy`` (sigma, body.0, body.2, head)."""

    def __init__(self, input_size, output_size, fc_neurons, hiddens=None):
        super().__init__()
        if hiddens is None:
            hiddens = [100, 100]
        self.activation = 'tanh'
        self.fc_neurons = fc_neurons
        layers = [linear_init(nn.Linear(input_size, hiddens[0])), nn.Tanh()]
        for i, o in zip(hiddens[:-1], hiddens[1:]):
            layers.append(linear_init(nn.Linear(i, o)))
            layers.append(nn.Tanh())
        self.body = nn.Sequential(*layers)
        self.head = linear_init(nn.Linear(fc_neurons, output_size))
        self.sigma = nn.Parameter(torch.Tensor(output_size))
        self.sigma.data.fill_(math.log(1))
        self.features_no_grad = False
        self.input_size, self.output_size, self.hiddens = input_size, output_size, list(hiddens)

    def turn_on_body_grads(self):
        self.features_no_grad = False

    def turn_off_body_grads(self):
        self.features_no_grad = True

    def forward_pass(self, state):
        if self.features_no_grad:
            with torch.no_grad():
                state_features = self.body(state)
        else:
            state_features = self.body(state)
        return self.head(state_features)

    def density(self, state):
        loc = self.forward_pass(state)
        scale = torch.exp(torch.clamp(self.sigma, min=math.log(EPSILON)))
        return Normal(loc=loc, scale=scale)

    def log_prob(self, state, action):
        return self.density(state).log_prob(action).mean(dim=1, keepdim=True)

    def forward(self, state):
        return self.density(state).sample()

    flat_parameters = DiagNormalPolicy.flat_parameters
    load_flat_parameters = DiagNormalPolicy.load_flat_parameters


class LinearValue(nn.Module):
    """Mirror of ``cherry.models.robotics.LinearValue`` (the baseline the reference passes around,
    rl/maml_trpo.py:85): ridge regression on [s, s^2, t, t^2, t^3, 1].  Inside the kernels the fit is part of
    ``xm_rl_advantages``; this object carries ``reg`` (and, after ``fit``, the coefficients) through the reference's
    call signatures."""

    def __init__(self, input_size, reg=1e-5):
        super().__init__()
        self.linear = nn.Linear(2 * input_size + 4, 1, bias=False)
        self.reg = reg

    def _features(self, states):
        n = states.size(0)
        ones = torch.ones(n, 1, dtype=states.dtype, device=states.device)
        al = torch.arange(n, dtype=states.dtype, device=states.device).view(-1, 1) / 100.0
        return torch.cat([states, states ** 2, al, al ** 2, al ** 3, ones], dim=1)

    def forward(self, states):
        return self.linear(self._features(states))
