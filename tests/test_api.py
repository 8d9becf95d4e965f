"""The reference-facing Python API (exploring_meta_b200.core_functions / utils), on the CPU emulator of the C ABI
(-m "not gpu") and on the real library on cuda:0 (-m gpu):
ConvBlock autograd (forward / backward / double-backward), model surface, MAML.clone()/adapt(), fast_adapt (generic
and engine routes) and evaluate, all against the oracle."""
import pytest
import torch
import torch.nn.functional as F

from exploring_meta_b200 import functional as XF
from exploring_meta_b200.core_functions import MAML, MiniImagenetCNN, OmniglotCNN, ConvBase, accuracy, evaluate, fast_adapt
from exploring_meta_b200.synthetic import SyntheticTasks, make_tasks
from exploring_meta_b200.utils import prepare_batch
from oracle import maml_oracle as mo


def _ref_block(x, gamma, beta, w, b, stride, pool):
    z = F.conv2d(x, w, b, stride=stride, padding=1)
    z = F.batch_norm(z, None, None, gamma, beta, training=True, eps=1e-5)
    z = F.relu(z)
    return F.max_pool2d(z, 2, 2) if pool else z


@pytest.mark.parametrize('cin,cout,h,stride,pool,nchw_image', [(3, 8, 12, 1, True, True), (8, 8, 9, 1, True, False),
                                                               (1, 8, 14, 2, False, True), (8, 4, 7, 2, False, False)])
def test_conv_block_autograd_through_double_backward(kdev, cin, cout, h, stride, pool, nchw_image):
    torch.manual_seed(0)
    x = torch.randn(5, cin, h, h).to(kdev)
    if not nchw_image:
        x = x.contiguous(memory_format=torch.channels_last)
    gamma, beta = (torch.rand(cout) + 0.5).to(kdev), (torch.randn(cout) * 0.1).to(kdev)
    w, b = (torch.randn(cout, cin, 3, 3) * 0.3).to(kdev), torch.zeros(cout).to(kdev)
    ins = [t.clone().requires_grad_(True) for t in (x, gamma, beta, w, b)]
    ref = [t.detach().cpu().double().requires_grad_(True) for t in (x, gamma, beta, w, b)]
    out, stats = XF.conv_block(*ins, stride=stride, pool=pool)
    out_ref = _ref_block(*ref, stride, pool)
    assert out.shape == out_ref.shape
    assert torch.allclose(out.cpu().double(), out_ref, atol=1e-5)
    probe = torch.randn_like(out)
    g = torch.autograd.grad((out * probe).sum(), ins, create_graph=True)
    g_ref = torch.autograd.grad((out_ref * probe.cpu().double()).sum(), ref, create_graph=True)
    for a, r in zip(g[:4], g_ref[:4]):
        assert mo.rel_l2(a.detach().cpu(), r.detach()) < 1e-4
    assert g[4].abs().max() == 0                                     # conv bias: cancelled by train-mode BN
    # double backward: d/d(inputs) of a random functional of the first-order gradients
    probes = [torch.randn_like(t) for t in g[:4]]
    s = sum((a * p).sum() for a, p in zip(g[:4], probes))
    s_ref = sum((a * p.cpu().double()).sum() for a, p in zip(g_ref[:4], probes))
    gg = torch.autograd.grad(s, ins[:4])
    gg_ref = torch.autograd.grad(s_ref, ref[:4])
    for a, r in zip(gg, gg_ref):
        assert mo.rel_l2(a.cpu(), r) < 2e-4


def test_model_surface_matches_reference_layout(kdev):
    m = MiniImagenetCNN(5)
    names = [n for n, _ in m.named_parameters()]
    assert names[:4] == ['base.0.normalize.weight', 'base.0.normalize.bias', 'base.0.conv.weight', 'base.0.conv.bias']
    assert names[-2:] == ['linear.weight', 'linear.bias']
    assert sum(p.numel() for p in m.parameters()) == 32901
    assert sum(p.numel() for p in OmniglotCNN(5).parameters()) == 112261
    assert sum(p.numel() for p in OmniglotCNN(20).parameters()) == 113236
    keys = set(m.state_dict())
    for i in range(4):
        for k in ('normalize.weight', 'normalize.bias', 'normalize.running_mean', 'normalize.running_var',
                  'normalize.num_batches_tracked', 'conv.weight', 'conv.bias'):
            assert 'base.%d.%s' % (i, k) in keys
    # same seed -> same initial parameters as the reference constructors (oracle.init_params restates them)
    torch.manual_seed(42)
    ours = [p.detach() for p in MiniImagenetCNN(5).parameters()]
    for a, b in zip(ours, mo.init_params(mo.miniimagenet_spec(5), seed=42)):
        assert torch.equal(a, b)
    body = ConvBase(output_size=64, channels=3, max_pool=True)       # ANIL body: hidden defaults to 64
    assert body[0].conv.weight.shape == (64, 3, 3, 3)


def test_forward_and_running_stats(kdev):
    torch.manual_seed(42)
    m = OmniglotCNN(5).to(kdev)
    x = torch.randn(6, 1, 28, 28)
    params = [p.detach().cpu().double() for p in m.parameters()]
    log = []
    ref = mo.net_forward(params, x.double(), mo.omniglot_spec(5), bn_log=log)
    x = x.to(kdev)
    out = m(x)
    assert torch.allclose(out.cpu().double(), ref, atol=1e-4)
    rm, rv = mo.compose_running_stats([torch.zeros(64).double()] * 4, [torch.ones(64).double()] * 4, log)
    for l, blk in enumerate(m.base):
        assert int(blk.normalize.num_batches_tracked) == 1
        assert torch.allclose(blk.normalize.running_mean.cpu().double(), rm[l], atol=1e-5)
        assert torch.allclose(blk.normalize.running_var.cpu().double(), rv[l], atol=1e-5)
    assert m.get_base_representation(x).shape == (6, 64, 2, 2)
    assert m.get_rep_layer(x, 2).shape == (6, 64, 7, 7)


def _reference_iteration(kind, ways, shots, steps, lr, X, Y, first_order=False):
    ospec = mo.omniglot_spec(ways) if kind == 'omni' else mo.miniimagenet_spec(ways)
    params = mo.init_params(ospec, seed=42)
    return ospec, mo.meta_iteration([p.double() for p in params], X.double(), Y, ospec, steps, lr, first_order=first_order)


@pytest.mark.parametrize('route', ['generic', 'engine'])
@pytest.mark.parametrize('kind,ways,shots,steps,lr', [('omni', 5, 1, 2, 0.4), ('min', 5, 1, 1, 0.05)])
def test_clone_adapt_fast_adapt_meta_gradient(kdev, route, kind, ways, shots, steps, lr):
    """The reference's train loop body (vision/maml_vision.py:95-112) written against this package's API."""
    shape = (1, 28, 28) if kind == 'omni' else (3, 84, 84)
    X, Y = make_tasks(2, ways, shots, shape, seed=5)
    ospec, ref = _reference_iteration(kind, ways, shots, steps, lr, X, Y)
    torch.manual_seed(42)
    model = (OmniglotCNN(ways) if kind == 'omni' else MiniImagenetCNN(ways)).to(kdev)
    maml = MAML(model, lr=lr, first_order=False)
    loss = torch.nn.CrossEntropyLoss(reduction='mean') if route == 'engine' else \
        (lambda logits, y: F.cross_entropy(logits, y))             # a plain callable forces the generic route
    losses, accs = [], []
    for t in range(X.size(0)):
        learner = maml.clone()
        eval_loss, eval_acc = fast_adapt((X[t], Y[t]), learner, loss, steps, shots, ways, kdev)
        eval_loss.backward()
        losses.append(eval_loss.item())
        accs.append(eval_acc.item())
        adapted = mo.flatten([p.detach().cpu() for p in learner.module.parameters()])
        mask = ~mo.conv_bias_mask(ospec)
        assert mo.rel_l2(adapted[mask], mo.flatten(ref['adapted'][t])[mask]) < 1e-4
    grad = mo.flatten([p.grad.cpu() for p in maml.parameters()])
    mask = ~mo.conv_bias_mask(ospec)
    assert mo.rel_l2(grad[mask], mo.flatten(ref['grad'])[mask]) < 5e-4
    assert torch.allclose(torch.tensor(losses, dtype=torch.float64), ref['loss'], rtol=1e-4, atol=1e-5)
    assert [round(a * ways * shots) for a in accs] == ref['correct'].tolist()
    # BN buffers are shared between the clones and the master: T support forwards + 1 query forward per task
    assert int(model.base[0].normalize.num_batches_tracked) == X.size(0) * (steps + 1)


def test_first_order_adapt(kdev):
    X, Y = make_tasks(1, 5, 1, (1, 28, 28), seed=9)
    ospec, ref = _reference_iteration('omni', 5, 1, 2, 0.3, X, Y, first_order=True)
    torch.manual_seed(42)
    maml = MAML(OmniglotCNN(5).to(kdev), lr=0.3, first_order=True)
    learner = maml.clone()
    loss, _ = fast_adapt((X[0], Y[0]), learner, lambda a, b: F.cross_entropy(a, b), 2, 1, 5, kdev)
    loss.backward()
    grad = mo.flatten([p.grad.cpu() for p in maml.parameters()])
    mask = ~mo.conv_bias_mask(ospec)
    assert mo.rel_l2(grad[mask], mo.flatten(ref['grad'])[mask]) < 5e-4


def test_prepare_batch_and_accuracy(kdev):
    data, labels = torch.arange(20.).view(10, 2), torch.arange(10)
    a, al, e, el = prepare_batch((data, labels), shots=1, ways=5, device=kdev)
    assert al.tolist() == [0, 2, 4, 6, 8] and el.tolist() == [1, 3, 5, 7, 9]
    assert torch.equal(a.cpu(), data[0::2]) and torch.equal(e.cpu(), data[1::2])
    logits = torch.tensor([[1., 1., 0.], [0., 2., 2.], [3., 0., 0.]])
    assert float(accuracy(logits, torch.tensor([0, 2, 0]))) == pytest.approx(2 / 3)     # ties -> lowest index


def test_evaluate_batched_equals_loop(kdev, capsys):
    params = {'meta_batch_size': 3, 'adapt_steps': 1, 'shots': 1, 'ways': 5}
    torch.manual_seed(42)
    maml = MAML(OmniglotCNN(5).to(kdev), lr=0.4)
    ce = torch.nn.CrossEntropyLoss(reduction='mean')
    acc_engine = evaluate(params, SyntheticTasks(5, 1, (1, 28, 28), seed=3), maml, ce, kdev)
    acc_loop = evaluate(params, SyntheticTasks(5, 1, (1, 28, 28), seed=3), maml, lambda a, b: F.cross_entropy(a, b), kdev)
    assert acc_engine == pytest.approx(acc_loop)
    assert 'Meta Test Accuracy' in capsys.readouterr().out
    X, Y = SyntheticTasks(5, 1, (1, 28, 28), seed=3).sample_batch(3)
    ospec = mo.omniglot_spec(5)
    ref = mo.meta_iteration([p.double() for p in mo.init_params(ospec, seed=42)], X.double(), Y, ospec, 1, 0.4)
    assert acc_engine == pytest.approx(float(ref['correct'].sum()) / 15)
