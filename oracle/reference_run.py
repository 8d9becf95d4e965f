"""Replay of the reference's train loop body on the reference's OWN files (build container only).

Follows vision/maml_vision.py:84-86,95-112 and vision/anil_vision.py:86-99,109-122 literally --
``MAML(model, lr, first_order=False)``, ``maml.clone()``, the reference ``fast_adapt``,
``eval_loss.backward()`` -- with ``train_tasks.sample()`` replaced by given synthetic tensors.
Used to pin ``oracle/maml_oracle.py`` and to generate ``tests/golden``.
"""
import torch

from . import ref_loader


def build_model(kind, ways, seed=42, dtype=torch.float32):
    ns = ref_loader.load()
    torch.manual_seed(seed)
    if kind == 'min':
        model = ns.MiniImagenetCNN(ways)
    elif kind == 'omni':
        model = ns.OmniglotCNN(ways)
    else:
        raise ValueError(kind)
    return model.to(dtype)


def maml_iteration(model, X, Y, ways, shots, steps, lr):
    ns = ref_loader.load()
    maml = ns.MAML(model, lr=lr, first_order=False)
    loss_fn = torch.nn.CrossEntropyLoss(reduction='mean')
    for p in maml.parameters():
        p.grad = None
    losses, accs, adapted = [], [], []
    for t in range(X.size(0)):
        learner = maml.clone()
        eval_loss, eval_acc = ns.fast_adapt((X[t], Y[t]), learner, loss_fn, steps, shots, ways,
                                            torch.device('cpu'))
        eval_loss.backward()
        losses.append(eval_loss.item())
        accs.append(eval_acc.item())
        adapted.append([p.detach().clone() for p in learner.parameters()])
    return {'loss': torch.tensor(losses, dtype=torch.float64), 'acc': torch.tensor(accs, dtype=torch.float64),
            'adapted': adapted, 'grad': [p.grad.detach().clone() for p in maml.parameters()]}


def build_anil(kind, ways, seed=42, dtype=torch.float32):
    """vision/anil_vision.py:86-94: ConvBase body (+ view) and a MAML-wrapped Linear head."""
    ns = ref_loader.load()
    torch.manual_seed(seed)
    if kind == 'omni':
        body, fc = ns.ConvBase(output_size=64, hidden=32, channels=1, max_pool=False), 128
    else:
        body, fc = ns.ConvBase(output_size=64, channels=3, max_pool=True), 1600
    features = torch.nn.Sequential(body, _View(fc)).to(dtype)
    head = torch.nn.Linear(fc, ways).to(dtype)
    return features, head


class _View(torch.nn.Module):
    def __init__(self, n):
        super().__init__()
        self.n = n

    def forward(self, x):
        return x.view(-1, self.n)


def anil_iteration(features, head_module, X, Y, ways, shots, steps, lr):
    ns = ref_loader.load()
    head = ns.MAML(head_module, lr=lr)
    loss_fn = torch.nn.CrossEntropyLoss(reduction='mean')
    allp = list(features.parameters()) + list(head.parameters())
    for p in allp:
        p.grad = None
    losses, accs = [], []
    for t in range(X.size(0)):
        learner = head.clone()
        eval_loss, eval_acc = ns.fast_adapt((X[t], Y[t]), learner, loss_fn, steps, shots, ways,
                                            torch.device('cpu'), features=features)
        eval_loss.backward()
        losses.append(eval_loss.item())
        accs.append(eval_acc.item())
    return {'loss': torch.tensor(losses, dtype=torch.float64), 'acc': torch.tensor(accs, dtype=torch.float64),
            'grad': [p.grad.detach().clone() for p in features.parameters()],
            'head_grad': [p.grad.detach().clone() for p in head.parameters()]}
