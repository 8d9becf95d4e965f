"""Restatement of the learn2learn functions the reference's hot path calls (TEST INFRASTRUCTURE).

learn2learn is an un-vendored, un-pinned dependency of the reference (call sites:
``core_functions/maml.py:8-9,45``, ``core_functions/vision.py:13``, ``vision/maml_vision.py:9,84,104``,
``vision/anil_vision.py:9,94,116``).  It is not installed here, so its published algorithm
(learn2learn >= 0.1.2: ``learn2learn/utils.py::clone_module/update_module`` and
``learn2learn/algorithms/maml.py::maml_update/MAML``) is restated below in our own words.
``install()`` registers the restatement under the ``learn2learn`` module names so that the
reference's files import unmodified (see ``oracle/ref_loader.py``).
"""
import sys
import types

import torch
from torch.autograd import grad as _autograd_grad


def clone_module(module, memo=None):
    """Differentiable structural copy: every parameter becomes ``p.clone()`` (a non-leaf that
    back-propagates into ``p``); buffers are shared unless they require grad; child modules recurse."""
    if memo is None:
        memo = {}
    if not isinstance(module, torch.nn.Module):
        return module
    twin = module.__new__(type(module))
    twin.__dict__ = module.__dict__.copy()
    twin._parameters = twin._parameters.copy()
    twin._buffers = twin._buffers.copy()
    twin._modules = twin._modules.copy()
    for name, p in module._parameters.items():
        if p is None:
            continue
        key = p.data_ptr()
        if key not in memo:
            memo[key] = p.clone()
        twin._parameters[name] = memo[key]
    for name, b in module._buffers.items():
        if b is None or not b.requires_grad:
            continue                      # BN running stats stay shared with the master
        key = b.data_ptr()
        if key not in memo:
            memo[key] = b.clone()
        twin._buffers[name] = memo[key]
    for name, child in twin._modules.items():
        twin._modules[name] = clone_module(child, memo)
    if hasattr(twin, 'flatten_parameters'):
        twin = twin._apply(lambda t: t)
    return twin


def update_module(module, updates=None, memo=None):
    """Re-binds each parameter to ``p + p.update`` (out of place, non-leaf)."""
    if memo is None:
        memo = {}
    if updates is not None:
        plist = list(module.parameters())
        if len(plist) != len(list(updates)):
            print('WARNING:update_module(): Parameters and updates have different length.')
        for p, u in zip(plist, updates):
            p.update = u
    for name, p in module._parameters.items():
        if p is None:
            continue
        if p in memo:
            module._parameters[name] = memo[p]
        elif getattr(p, 'update', None) is not None:
            fresh = p + p.update
            p.update = None
            memo[p] = fresh
            module._parameters[name] = fresh
    for name, b in module._buffers.items():
        if b is None:
            continue
        if b in memo:
            module._buffers[name] = memo[b]
        elif getattr(b, 'update', None) is not None:
            fresh = b + b.update
            b.update = None
            memo[b] = fresh
            module._buffers[name] = fresh
    for name, child in module._modules.items():
        module._modules[name] = update_module(child, updates=None, memo=memo)
    if hasattr(module, 'flatten_parameters'):
        module._apply(lambda t: t)
    return module


def maml_update(model, lr, grads=None):
    if grads is not None:
        plist = list(model.parameters())
        if len(grads) != len(plist):
            print('WARNING:maml_update(): Parameters and gradients have different length.')
        for p, g in zip(plist, grads):
            if g is not None:
                p.update = -lr * g
    return update_module(model)


def magic_box(x):
    return torch.exp(x - x.detach())


class BaseLearner(torch.nn.Module):
    def __init__(self, module=None):
        super().__init__()
        self.module = module

    def __getattr__(self, attr):
        try:
            return super().__getattr__(attr)
        except AttributeError:
            return getattr(self.__dict__['_modules']['module'], attr)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


class MAML(BaseLearner):
    def __init__(self, model, lr, first_order=False, allow_unused=None, allow_nograd=False):
        super().__init__()
        self.module = model
        self.lr = lr
        self.first_order = first_order
        self.allow_nograd = allow_nograd
        if allow_unused is None:
            allow_unused = allow_nograd
        self.allow_unused = allow_unused

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def adapt(self, loss, first_order=None, allow_unused=None, allow_nograd=None):
        if first_order is None:
            first_order = self.first_order
        if allow_unused is None:
            allow_unused = self.allow_unused
        if allow_nograd is None:
            allow_nograd = self.allow_nograd
        second_order = not first_order
        if allow_nograd:
            diff = [p for p in self.module.parameters() if p.requires_grad]
            got = _autograd_grad(loss, diff, retain_graph=second_order,
                                 create_graph=second_order, allow_unused=allow_unused)
            grads, i = [], 0
            for p in self.module.parameters():
                if p.requires_grad:
                    grads.append(got[i])
                    i += 1
                else:
                    grads.append(None)
        else:
            grads = _autograd_grad(loss, self.module.parameters(), retain_graph=second_order,
                                   create_graph=second_order, allow_unused=allow_unused)
        self.module = maml_update(self.module, self.lr, grads)

    def clone(self, first_order=None, allow_unused=None, allow_nograd=None):
        if first_order is None:
            first_order = self.first_order
        if allow_unused is None:
            allow_unused = self.allow_unused
        if allow_nograd is None:
            allow_nograd = self.allow_nograd
        return MAML(clone_module(self.module), lr=self.lr, first_order=first_order,
                    allow_unused=allow_unused, allow_nograd=allow_nograd)


def install():
    """Register the restatement under the learn2learn names the reference imports."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    placeholder = type('Placeholder', (), {'__init__': lambda self, *a, **k: None})
    root = mod('learn2learn', clone_module=clone_module, update_module=update_module,
               magic_box=magic_box)
    root.algorithms = mod('learn2learn.algorithms', MAML=MAML)
    root.algorithms.maml = mod('learn2learn.algorithms.maml', MAML=MAML, maml_update=maml_update)
    root.utils = mod('learn2learn.utils', clone_module=clone_module, update_module=update_module)
    root.data = mod('learn2learn.data')
    root.data.transforms = mod('learn2learn.data.transforms',
                               **{n: placeholder for n in ('NWays', 'KShots', 'LoadData', 'RemapLabels',
                                                           'ConsecutiveLabels', 'FilterLabels')})
    root.vision = mod('learn2learn.vision')
    root.vision.transforms = mod('learn2learn.vision.transforms', RandomClassRotation=placeholder)
    return root
