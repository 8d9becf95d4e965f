"""The oracle against the committed golden vectors (generated from the reference's own files by
tests/golden/make_golden.py) and, when /root/reference is present, against a live run of those files.

PARITY UNPINNED by the reference itself (it ships no tests or fixtures and does not pin learn2learn);
these fixtures are outputs of the reference's unmodified hot-path files run on top of the learn2learn
restatement (oracle/l2l_shim.py)."""
import pytest
import torch

import golden_util as gu
from oracle import l2l_shim, maml_oracle as mo, ref_loader

# every fixture is cheap enough for the CPU suite in fp64
CPU_CASES = list(gu.NAMES)


@pytest.mark.parametrize('name', CPU_CASES)
def test_oracle_reproduces_golden(name):
    g = gu.Golden(name)
    X, Y, params, head = g.inputs()
    out = mo.meta_iteration([p.double() for p in params], X.double(), Y, g.ospec(), g.steps, g.inner_lr,
                            anil_head=None if head is None else [h.double() for h in head])
    mask = g.grad_mask()
    assert mo.rel_l2(mo.flatten(out['grad'])[mask], g.t('grad64_as_f32')[mask]) < 5e-7     # fixture is fp32-rounded
    assert torch.allclose(out['loss'], g.t('loss64'), rtol=1e-11, atol=1e-12)
    assert out['correct'].tolist() == g.t('correct').tolist()
    if g.algo == 'maml':
        assert mo.rel_l2(mo.flatten(out['adapted'][0]), g.t('adapted0_64_as_f32')) < 5e-7
    else:
        assert mo.rel_l2(mo.flatten(out['head_grad']), g.t('head_grad64_as_f32')) < 5e-7


def test_golden_records_reference_noise_floor():
    """The fixtures carry the reference's own fp32-vs-fp64 deviation; the headline configuration is in the
    chaotic regime, the calm ones are not."""
    assert float(gu.Golden('maml_min_5w5s_t5_headline').z['e_ref_grad']) > 1e-2
    assert float(gu.Golden('maml_min_5w1s_t2_calm').z['e_ref_grad']) < 1e-4
    assert float(gu.Golden('maml_omni_5w1s_t1').z['e_ref_grad']) < 1e-3


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present (GPU box)')
def test_restatement_matches_reference_files_live():
    from oracle import reference_run as rr
    from exploring_meta_b200.synthetic import make_tasks
    X, Y = make_tasks(2, 5, 1, (1, 28, 28), seed=11)
    model = rr.build_model('omni', 5, seed=42, dtype=torch.float64)
    ref = rr.maml_iteration(model, X.double(), Y, 5, 1, 2, 0.4)
    ospec = mo.omniglot_spec(5)
    params = mo.init_params(ospec, seed=42, dtype=torch.float64)
    out = mo.meta_iteration(params, X.double(), Y, ospec, 2, 0.4)
    assert mo.rel_l2(mo.flatten(out['grad']), mo.flatten(ref['grad'])) < 1e-12
    assert torch.allclose(out['loss'], ref['loss'], rtol=1e-12)
    for a, b in zip(out['adapted'][1], ref['adapted'][1]):
        assert torch.allclose(a, b, rtol=1e-12, atol=1e-14)


def test_l2l_restatement_semantics():
    """clone(): parameters become differentiable non-leaf copies, BN buffers stay shared with the master;
    adapt(): out-of-place p + (-lr * g), second-order graph kept."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3, padding=1), torch.nn.BatchNorm2d(4), torch.nn.ReLU(),
                              torch.nn.Flatten(), torch.nn.Linear(4 * 36, 3)).double()
    maml = l2l_shim.MAML(net, lr=0.1)
    learner = maml.clone()
    for (n0, p0), (n1, p1) in zip(net.named_parameters(), learner.module.named_parameters()):
        assert n0 == n1 and not p1.is_leaf and torch.equal(p0, p1)
    assert learner.module[1].running_mean is net[1].running_mean
    x, y = torch.randn(6, 1, 6, 6, dtype=torch.float64), torch.tensor([0, 1, 2, 0, 1, 2])
    before = [p for p in learner.module.parameters()]
    loss = torch.nn.functional.cross_entropy(learner(x), y)
    g = torch.autograd.grad(loss, before, create_graph=True)
    learner.adapt(loss)
    assert int(net[1].num_batches_tracked) == 1                  # the clone's forward updated the master's buffer
    for p_old, p_new, gi in zip(before, learner.module.parameters(), g):
        assert p_new is not p_old and torch.allclose(p_new, p_old - 0.1 * gi)
    q = torch.nn.functional.cross_entropy(learner(x), y)
    q.backward()
    assert all(p.grad is not None for p in net.parameters())
    # first-order clone: no graph through the inner gradient
    fo = maml.clone(first_order=True)
    loss = torch.nn.functional.cross_entropy(fo(x), y)
    fo.adapt(loss)
    assert all(p.grad_fn is not None for p in fo.module.parameters())
