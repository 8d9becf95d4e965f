"""``fast_adapt`` / ``accuracy`` / ``evaluate`` with the reference's signatures (``core_functions/vision.py:6-42``).

Two execution routes, same results:

* generic (any learner, any loss, optional ``features``): the reference's own sequence -- ``prepare_batch``; for each
  adaptation step ``learner.adapt(loss(learner(adapt_data), adapt_labels))``; query loss / accuracy -- where every
  ConvBlock forward, backward and double-backward runs on the CUDA kernels through the autograd functions of
  ``exploring_meta_b200.functional``;
* engine route, taken when the learner is an un-adapted ``MAML.clone()`` of one of this package's CNNs and the loss
  is a plain mean cross-entropy: the whole task (inner loop + query pass) is one static launch program of the
  task-batched engine with ``tasks = 1``; the returned loss is attached to the master parameters by an autograd node
  whose backward launches the second-order program, so ``valid_loss.backward()`` accumulates the meta-gradient into
  the master ``.grad`` exactly like the reference.  ``evaluate`` batches its ``meta_batch_size`` tasks into one
  launch program.
"""
import torch

from ..utils.data_pre import prepare_batch
from .maml import MAML
from .vision_models import net_spec_of

_ENGINES = {}


def _engine(spec, tasks, shots, steps, lr, mode, device):
    from ..engine import MamlEngine
    key = (spec, int(tasks), int(shots), int(steps), float(lr), mode, str(device))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) > 16:
            _ENGINES.clear()
        eng = _ENGINES[key] = MamlEngine(spec, tasks, shots, steps, lr, mode=mode, device=device)
    return eng


def _plain_cross_entropy(loss):
    return (isinstance(loss, torch.nn.CrossEntropyLoss) and loss.reduction == 'mean' and loss.weight is None
            and loss.ignore_index == -100 and getattr(loss, 'label_smoothing', 0.0) == 0.0)


def _on_kernel_device(device):
    from .. import engine as _engine
    try:
        _engine._require_cuda(torch.device(device))
        return True
    except Exception:
        return False


def _master_of(learner):
    return learner.__dict__.get('_xm_fresh') if isinstance(learner, MAML) else None


def _bn_layers(module):
    return [blk.normalize for blk in module.base.children()]


def _apply_bn_side_effects(engine, module):
    bns = _bn_layers(module)
    n = engine.update_running_stats([b.running_mean for b in bns], [b.running_var for b in bns])
    for b in bns:
        b.num_batches_tracked += n


class _EngineTask(torch.autograd.Function):
    """loss(theta) of one task through the engine; backward = the second-order launch program."""

    @staticmethod
    def forward(ctx, x, y, cfg, *params):
        spec, shots, steps, lr, first_order = cfg
        theta = torch.cat([p.detach().reshape(-1) for p in params])
        eng = _engine(spec, 1, shots, steps, lr, 'eval', x.device)
        eng.run(x.unsqueeze(0), y.unsqueeze(0), theta)
        ctx.cfg, ctx.x, ctx.y, ctx.theta = cfg, x, y, theta
        ctx.shapes = [p.shape for p in params]
        ctx.mark_non_differentiable(eng.correct)
        return eng.loss[0].clone(), eng.correct[0].clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gloss, _gcorrect):
        spec, shots, steps, lr, first_order = ctx.cfg
        eng = _engine(spec, 1, shots, steps, lr, 'first' if first_order else 'second', ctx.x.device)
        eng.run(ctx.x.unsqueeze(0), ctx.y.unsqueeze(0), ctx.theta)
        flat = eng.grad * gloss
        outs, o = [], 0
        for shp in ctx.shapes:
            n = 1
            for s in shp:
                n *= s
            outs.append(flat[o:o + n].reshape(shp))
            o += n
        return (None, None, None) + tuple(outs)


def _engine_route(batch, learner, loss, adaptation_steps, shots, ways, device, features):
    master = _master_of(learner)
    if master is None or features is not None or not _plain_cross_entropy(loss):
        return None
    spec = net_spec_of(master)
    data, labels = batch
    if (spec is None or spec.ways < ways or data.dtype != torch.float32 or data.size(0) != 2 * shots * ways
            or not _on_kernel_device(device) or not master.training):
        return None
    if spec.ways != ways:
        return None
    return master, spec


def fast_adapt(batch, learner, loss, adaptation_steps, shots, ways, device, features=None):
    route = _engine_route(batch, learner, loss, adaptation_steps, shots, ways, device, features)
    if route is not None:
        master, spec = route
        data, labels = batch
        x = data.to(device).reshape(data.size(0), spec.in_c, spec.in_h, spec.in_w).contiguous()
        y = labels.to(device).contiguous()
        cfg = (spec, shots, adaptation_steps, float(learner.lr), bool(learner.first_order))
        valid_loss, correct = _EngineTask.apply(x, y, cfg, *master.parameters())
        eng = _engine(spec, 1, shots, adaptation_steps, float(learner.lr), 'eval', x.device)
        with torch.no_grad():
            _apply_bn_side_effects(eng, master)
            if adaptation_steps > 0:            # leave the learner holding theta_T, like the reference does
                flat, o = eng.theta_steps[adaptation_steps - 1, 0], 0
                for name_holder, name, p in _named_param_slots(learner.module):
                    n = p.numel()
                    name_holder._parameters[name] = flat[o:o + n].reshape(p.shape).clone()
                    o += n
        learner.__dict__['_xm_fresh'] = None
        return valid_loss, correct.float() / (shots * ways)

    adapt_data, adapt_labels, eval_data, eval_labels = prepare_batch(batch, shots, ways, device, features=features)
    for _step in range(adaptation_steps):
        train_loss = loss(learner(adapt_data), adapt_labels)
        learner.adapt(train_loss)
    predictions = learner(eval_data)
    valid_loss = loss(predictions, eval_labels)
    valid_accuracy = accuracy(predictions, eval_labels)
    return valid_loss, valid_accuracy


def _named_param_slots(module):
    """(owner module, name, tensor) for every parameter in ``parameters()`` order."""
    for sub in module.modules():
        for name, p in sub._parameters.items():
            if p is not None:
                yield sub, name, p


def accuracy(predictions, targets):
    predictions = predictions.argmax(dim=1).view(targets.shape)
    return (predictions == targets).sum().float() / targets.size(0)


def evaluate(params, test_tasks, model, loss, device, features=None):
    """Meta-test loop (core_functions/vision.py:26-42).  With one of this package's CNNs behind ``model`` and a
    plain cross-entropy, the ``meta_batch_size`` sampled tasks run as ONE task-batched launch program."""
    B = params['meta_batch_size']
    spec = net_spec_of(model.module) if isinstance(model, MAML) else None
    if (spec is not None and features is None and _plain_cross_entropy(loss) and spec.ways == params['ways']
            and _on_kernel_device(device) and model.module.training):
        batches = [test_tasks.sample() for _ in range(B)]
        x = torch.stack([b[0] for b in batches]).to(device).reshape(B, -1, spec.in_c, spec.in_h, spec.in_w)
        y = torch.stack([b[1] for b in batches]).to(device)
        if x.dtype == torch.float32 and x.size(1) == 2 * params['shots'] * params['ways']:
            eng = _engine(spec, B, params['shots'], params['adapt_steps'], float(model.lr), 'eval', x.device)
            theta = torch.cat([p.detach().reshape(-1) for p in model.module.parameters()])
            eng.run(x.contiguous(), y.contiguous(), theta)
            with torch.no_grad():
                _apply_bn_side_effects(eng, model.module)
            meta_test_accuracy = float(eng.correct.sum().item()) / (B * params['shots'] * params['ways'])
            print('Meta Test Accuracy', meta_test_accuracy)
            return meta_test_accuracy
        test_tasks = _Replay(batches)
    meta_test_loss = 0.0
    meta_test_accuracy = 0.0
    for _task in range(B):
        learner = model.clone()
        batch = test_tasks.sample()
        eval_loss, eval_acc = fast_adapt(batch, learner, loss, params['adapt_steps'], params['shots'], params['ways'],
                                         device, features=features)
        meta_test_loss += eval_loss.item()
        meta_test_accuracy += eval_acc.item()
    meta_test_accuracy = meta_test_accuracy / B
    print('Meta Test Accuracy', meta_test_accuracy)
    return meta_test_accuracy


class _Replay:
    def __init__(self, batches):
        self._it = iter(batches)

    def sample(self):
        return next(self._it)
