"""``MAML`` wrapper with the reference's surface (``core_functions/maml.py:12-49`` on top of learn2learn's
``MAML`` -- ``__init__(model, lr, first_order, allow_unused, allow_nograd)``, ``forward``, ``clone``, ``adapt``,
attribute pass-through, ``get_rep`` / ``get_rep_i``).

learn2learn is not a dependency here: ``clone()`` and ``adapt()`` implement its published semantics directly.
``clone()`` returns a wrapper around a structural copy of the module whose parameters are differentiable
``.clone()``s of the originals (so a loss on the copy back-propagates into the master's ``.grad``) and whose
non-differentiable buffers -- the BatchNorm running statistics -- stay shared with the master.  ``adapt(loss)``
takes ``torch.autograd.grad(loss, parameters, create_graph=not first_order)`` and re-binds every parameter to
``p + (-lr * g)`` out of place, keeping the graph for the outer gradient.  With the models of
``core_functions/vision_models.py`` every derivative involved (first and second order) executes on libxmeta's
CUDA kernels (see exploring_meta_b200/functional.py).
"""
import traceback

import torch


def _copy_shell(module):
    """A new module object of the same class sharing nothing mutable with ``module`` at the top level."""
    twin = module.__new__(type(module))
    twin.__dict__ = dict(module.__dict__)
    for slot in ('_parameters', '_buffers', '_modules'):
        twin.__dict__[slot] = dict(module.__dict__[slot])
    return twin


def clone_module(module, _seen=None):
    """Differentiable structural copy (learn2learn ``clone_module`` semantics)."""
    if not isinstance(module, torch.nn.Module):
        return module
    seen = {} if _seen is None else _seen
    twin = _copy_shell(module)
    for name, param in module._parameters.items():
        if param is not None:
            twin._parameters[name] = seen.setdefault(param.data_ptr(), param.clone())
    for name, buf in module._buffers.items():
        if buf is not None and buf.requires_grad:          # running statistics do not: they stay shared
            twin._buffers[name] = seen.setdefault(buf.data_ptr(), buf.clone())
    for name, child in module._modules.items():
        twin._modules[name] = clone_module(child, seen)
    return twin


def update_module(module, updates=None, _done=None):
    """Re-binds every parameter that carries an ``.update`` to ``p + p.update`` (out of place)."""
    done = {} if _done is None else _done
    if updates is not None:
        params = list(module.parameters())
        updates = list(updates)
        if len(params) != len(updates):
            print('WARNING:update_module(): Parameters and updates have different length. ('
                  + str(len(params)) + ' vs ' + str(len(updates)) + ')')
        for p, u in zip(params, updates):
            p.update = u
    for store in (module._parameters, module._buffers):
        for name, t in store.items():
            if t is None:
                continue
            if t in done:
                store[name] = done[t]
            elif getattr(t, 'update', None) is not None:
                new = t + t.update
                t.update = None
                done[t] = new
                store[name] = new
    for name, child in module._modules.items():
        module._modules[name] = update_module(child, None, done)
    return module


def maml_update(model, lr, grads=None):
    """theta' = theta + (-lr * g) for every parameter with a gradient (learn2learn ``maml_update``)."""
    if grads is not None:
        params = list(model.parameters())
        if len(grads) != len(params):
            print('WARNING:maml_update(): Parameters and gradients have different length. ('
                  + str(len(params)) + ' vs ' + str(len(grads)) + ')')
        for p, g in zip(params, grads):
            if g is not None:
                p.update = -lr * g
    return update_module(model)


class MAML(torch.nn.Module):
    def __init__(self, model, lr, first_order=False, allow_unused=None, allow_nograd=False):
        super().__init__()
        self.module = model
        self.lr = lr
        self.first_order = first_order
        self.allow_nograd = allow_nograd
        self.allow_unused = allow_nograd if allow_unused is None else allow_unused

    def __getattr__(self, attr):
        try:
            return super().__getattr__(attr)
        except AttributeError:
            return getattr(self.__dict__['_modules']['module'], attr)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    # core_functions/maml.py:15-19
    def get_rep(self, input_d):
        return self.get_base_representation(input_d)

    def get_rep_i(self, input_d, layer_i):
        return self.get_rep_layer(input_d, layer_i)

    def adapt(self, loss, first_order=None, allow_unused=None, allow_nograd=None):
        first_order = self.first_order if first_order is None else first_order
        allow_unused = self.allow_unused if allow_unused is None else allow_unused
        allow_nograd = self.allow_nograd if allow_nograd is None else allow_nograd
        second_order = not first_order
        self.__dict__['_xm_fresh'] = None                   # no longer an un-adapted clone
        if allow_nograd:
            wanted = [p for p in self.module.parameters() if p.requires_grad]
            got = list(torch.autograd.grad(loss, wanted, retain_graph=second_order, create_graph=second_order,
                                           allow_unused=allow_unused))
            grads = [got.pop(0) if p.requires_grad else None for p in self.module.parameters()]
        else:
            try:
                grads = torch.autograd.grad(loss, self.module.parameters(), retain_graph=second_order,
                                            create_graph=second_order, allow_unused=allow_unused)
            except RuntimeError:
                traceback.print_exc()
                print('MAML.adapt(): maybe try with allow_nograd=True and/or allow_unused=True ?')
                raise
        self.module = maml_update(self.module, self.lr, grads)

    def clone(self, first_order=None, allow_unused=None, allow_nograd=None):
        """core_functions/maml.py:23-49."""
        first_order = self.first_order if first_order is None else first_order
        allow_unused = self.allow_unused if allow_unused is None else allow_unused
        allow_nograd = self.allow_nograd if allow_nograd is None else allow_nograd
        twin = MAML(clone_module(self.module), lr=self.lr, first_order=first_order, allow_unused=allow_unused,
                    allow_nograd=allow_nograd)
        # remembered so that fast_adapt can hand a fresh clone of a known network to the task-batched engine
        twin.__dict__['_xm_fresh'] = self.module
        return twin
