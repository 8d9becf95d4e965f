"""Self-contained CPU restatement of the reference's MAML / ANIL hot path (TEST INFRASTRUCTURE).

PARITY UNPINNED by the reference itself (no tests / goldens exist, learn2learn is un-pinned; see
``oracle/__init__.py``).  This file is checked against the reference's own unmodified files
(``oracle/ref_loader.py``) by ``tests/golden/make_golden.py`` and ``tests/test_oracle.py``.

Everything is functional over a flat list of parameter tensors in ``module.parameters()`` order
(per block: ``normalize.weight, normalize.bias, conv.weight, conv.bias`` -- BN is registered before
the conv, ``core_functions/vision_models.py:168-185`` -- then ``linear.weight, linear.bias``), runs
on the CPU through ``torch.autograd`` and works in fp32 or fp64.  The tolerance contract is stated
against the fp64 run.
"""
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class NetSpec:
    """Shape of a reference network: ``ConvBase`` body (+ optional linear head)."""
    in_c: int
    in_h: int
    in_w: int
    hidden: int
    ways: int
    layers: int = 4
    pool: bool = True        # True: stride-1 conv + MaxPool2d(2,2)   False: stride-2 conv, no pool
    head: str = 'flatten'    # 'flatten' (MiniImagenetCNN :107-110) | 'mean' (OmniglotCNN :51-55)

    def out_hw(self):
        h, w = self.in_h, self.in_w
        for _ in range(self.layers):
            if self.pool:
                h, w = h // 2, w // 2
            else:
                h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        return h, w

    def feat_dim(self):
        h, w = self.out_hw()
        return self.hidden * h * w if self.head == 'flatten' else self.hidden

    def param_shapes(self, with_head=True):
        shapes, cin = [], self.in_c
        for _ in range(self.layers):
            shapes += [(self.hidden,), (self.hidden,), (self.hidden, cin, 3, 3), (self.hidden,)]
            cin = self.hidden
        if with_head:
            shapes += [(self.ways, self.feat_dim()), (self.ways,)]
        return shapes


def miniimagenet_spec(ways=5, hidden=32):
    """``MiniImagenetCNN(ways)`` -- core_functions/vision_models.py:91-105."""
    return NetSpec(3, 84, 84, hidden, ways, 4, True, 'flatten')


def omniglot_spec(ways=5, hidden=64):
    """``OmniglotCNN(ways)`` -- core_functions/vision_models.py:38-49."""
    return NetSpec(1, 28, 28, hidden, ways, 4, False, 'mean')


def init_params(spec, seed=42, with_head=True, dtype=torch.float32):
    """Reproduces the RNG consumption of the reference constructors under ``torch.manual_seed``:
    per block BatchNorm2d (no RNG) -> uniform_(gamma) (:175) -> Conv2d default init -> xavier_uniform_
    + zero bias (:186, :204-207); then Linear default init followed by xavier/zero
    (MiniImagenetCNN :103-104) or normal_()/zero (OmniglotCNN :47-49)."""
    torch.manual_seed(seed)
    params, cin = [], spec.in_c
    for _ in range(spec.layers):
        bn = torch.nn.BatchNorm2d(spec.hidden, affine=True)
        torch.nn.init.uniform_(bn.weight)
        conv = torch.nn.Conv2d(cin, spec.hidden, (3, 3), stride=1 if spec.pool else 2, padding=1, bias=True)
        torch.nn.init.xavier_uniform_(conv.weight.data, gain=1.0)
        torch.nn.init.constant_(conv.bias.data, 0.0)
        params += [bn.weight, bn.bias, conv.weight, conv.bias]
        cin = spec.hidden
    if with_head:
        lin = torch.nn.Linear(spec.feat_dim(), spec.ways, bias=True)
        if spec.head == 'flatten':
            torch.nn.init.xavier_uniform_(lin.weight.data, gain=1.0)
            torch.nn.init.constant_(lin.bias.data, 0.0)
        else:
            lin.weight.data.normal_()
            lin.bias.data.mul_(0.0)
        params += [lin.weight, lin.bias]
    return [p.detach().clone().to(dtype) for p in params]


def init_anil_params(spec, seed=42, dtype=torch.float32):
    """ANIL parameters as ``vision/anil_vision.py:86-94`` creates them under ``torch.manual_seed``: the
    ``ConvBase`` body (same RNG consumption as the blocks of ``init_params``) and then a default-initialised
    ``torch.nn.Linear(fc_neurons, ways)`` head.  Returns (body list, head list)."""
    body = init_params(spec, seed=seed, with_head=False, dtype=dtype)     # leaves the RNG after the blocks
    h, w = spec.out_hw()
    lin = torch.nn.Linear(spec.hidden * h * w, spec.ways)
    return body, [lin.weight.detach().clone().to(dtype), lin.bias.detach().clone().to(dtype)]


def body_forward(params, x, spec, bn_log=None):
    """``ConvBase.forward`` = 4x ``ConvBlock.forward`` (vision_models.py:188-193): conv -> BN with
    per-call batch statistics (no .eval() exists anywhere in the reference) -> ReLU -> max-pool."""
    h = x.reshape(-1, spec.in_c, spec.in_h, spec.in_w)
    for l in range(spec.layers):
        gamma, beta, w, b = params[4 * l:4 * l + 4]
        h = F.conv2d(h, w, b, stride=1 if spec.pool else 2, padding=1)
        if bn_log is not None:
            with torch.no_grad():
                n = h.numel() // h.size(1)
                mean = h.mean(dim=(0, 2, 3))
                var_unbiased = h.var(dim=(0, 2, 3), unbiased=False) * (n / (n - 1))
                bn_log.append((l, mean.clone(), var_unbiased.clone()))
        h = F.batch_norm(h, None, None, gamma, beta, training=True, momentum=0.1, eps=1e-5)
        h = F.relu(h)
        if spec.pool:
            h = F.max_pool2d(h, 2, 2)
    return h


def head_features(h, spec):
    if spec.head == 'flatten':
        return h.reshape(h.size(0), -1)
    return h.mean(dim=[2, 3])


def net_forward(params, x, spec, bn_log=None):
    feats = head_features(body_forward(params, x, spec, bn_log), spec)
    return F.linear(feats, params[4 * spec.layers], params[4 * spec.layers + 1])


def split_task(x, y):
    """``prepare_batch`` (utils/data_pre.py:121-127): even rows adapt, odd rows evaluate."""
    return x[0::2], y[0::2], x[1::2], y[1::2]


def count_correct(logits, targets):
    """``accuracy`` (core_functions/vision.py:21-23) without the division: argmax == target count."""
    return int((logits.argmax(dim=1).view(targets.shape) == targets).sum().item())


def maml_task(master, x, y, spec, steps, lr, first_order=False, bn_log=None):
    """One ``maml.clone()`` + ``fast_adapt`` (vision.py:6-18, learn2learn MAML.adapt): returns the
    query loss (carrying the graph back to ``master``), the correct-count, and theta_T."""
    xs, ys, xq, yq = split_task(x, y)
    fast = [p.clone() for p in master]
    for _ in range(steps):
        loss = F.cross_entropy(net_forward(fast, xs, spec, bn_log), ys)
        grads = torch.autograd.grad(loss, fast, retain_graph=not first_order,
                                    create_graph=not first_order)
        fast = [p + (-lr * g) for p, g in zip(fast, grads)]
    logits = net_forward(fast, xq, spec, bn_log)
    return F.cross_entropy(logits, yq), count_correct(logits, yq), fast


def anil_task(body, head, x, y, spec, steps, lr, first_order=False, bn_log=None):
    """ANIL (vision/anil_vision.py:116-122 + data_pre.py:118-119): body forward ONCE over all 2S
    rows (one BN batch), then only the linear head is adapted on the feature rows."""
    feats = head_features(body_forward(body, x, spec, bn_log), spec)
    fs, ys, fq, yq = split_task(feats, y)
    fast = [p.clone() for p in head]
    for _ in range(steps):
        loss = F.cross_entropy(F.linear(fs, fast[0], fast[1]), ys)
        grads = torch.autograd.grad(loss, fast, retain_graph=not first_order,
                                    create_graph=not first_order)
        fast = [p + (-lr * g) for p, g in zip(fast, grads)]
    logits = F.linear(fq, fast[0], fast[1])
    return F.cross_entropy(logits, yq), count_correct(logits, yq), fast


def compose_running_stats(running_mean, running_var, calls, momentum=0.1):
    """Sequential BN EMA side effect on the SHARED master buffers, one update per forward call
    (learn2learn clones share buffers): r <- (1-m) r + m s, unbiased variance."""
    rm = [t.clone() for t in running_mean]
    rv = [t.clone() for t in running_var]
    for l, mean, var_u in calls:
        rm[l] = (1 - momentum) * rm[l] + momentum * mean
        rv[l] = (1 - momentum) * rv[l] + momentum * var_u
    return rm, rv


def meta_iteration(master, X, Y, spec, steps, lr, first_order=False, anil_head=None):
    """The train half of one outer iteration (vision/maml_vision.py:95-112): for every task clone,
    fast_adapt, ``eval_loss.backward()`` accumulating into the master ``.grad`` in task order.

    ``master``: list of tensors (body [+ head] parameters); for ANIL pass the body list as
    ``master`` and the two head tensors as ``anil_head``.  Returns a dict with per-task query loss,
    correct counts, theta_T, the summed meta-gradient (NOT yet scaled by 1/B) and the per-call BN
    batch statistics in call order."""
    leaves = [p.detach().clone().requires_grad_(True) for p in master]
    head_leaves = None
    if anil_head is not None:
        head_leaves = [p.detach().clone().requires_grad_(True) for p in anil_head]
    losses, corrects, adapted, bn_log = [], [], [], []
    for t in range(X.size(0)):
        if anil_head is None:
            loss, correct, fast = maml_task(leaves, X[t], Y[t], spec, steps, lr, first_order, bn_log)
        else:
            loss, correct, fast = anil_task(leaves, head_leaves, X[t], Y[t], spec, steps, lr,
                                            first_order, bn_log)
        loss.backward()
        losses.append(float(loss.item()))
        corrects.append(correct)
        adapted.append([p.detach().clone() for p in fast])
    out = {
        'loss': torch.tensor(losses, dtype=torch.float64),
        'correct': torch.tensor(corrects, dtype=torch.int64),
        'adapted': adapted,
        'grad': [p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for p in leaves],
        'bn_calls': bn_log,
    }
    if head_leaves is not None:
        out['head_grad'] = [p.grad.detach().clone() for p in head_leaves]
    return out


def adam_step(params, grads, state, lr=0.003, betas=(0.9, 0.999), eps=1e-8):
    """``torch.optim.Adam`` defaults as used at vision/maml_vision.py:85,141 (no weight decay,
    no amsgrad).  ``state`` = dict(step, m, v); returns the new parameter list."""
    state['step'] += 1
    t = state['step']
    b1, b2 = betas
    out = []
    for i, (p, g) in enumerate(zip(params, grads)):
        state['m'][i] = b1 * state['m'][i] + (1 - b1) * g
        state['v'][i] = b2 * state['v'][i] + (1 - b2) * g * g
        denom = state['v'][i].sqrt() / (1 - b2 ** t) ** 0.5 + eps
        out.append(p - (lr / (1 - b1 ** t)) * state['m'][i] / denom)
    return out


def new_adam_state(params):
    return {'step': 0, 'm': [torch.zeros_like(p) for p in params], 'v': [torch.zeros_like(p) for p in params]}


def flatten(tensors):
    return torch.cat([t.reshape(-1) for t in tensors])


def rel_l2(a, b):
    """||a-b|| / ||b|| in float64."""
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def conv_bias_mask(spec, with_head=True):
    """Boolean mask over the flat parameter vector selecting the four ``conv.bias`` tensors, whose
    gradient is analytically zero under train-mode BN and is compared absolutely."""
    parts = []
    for i, shp in enumerate(spec.param_shapes(with_head)):
        n = 1
        for s in shp:
            n *= s
        is_cb = (i < 4 * spec.layers) and (i % 4 == 3)
        parts.append(torch.full((n,), is_cb, dtype=torch.bool))
    return torch.cat(parts)
