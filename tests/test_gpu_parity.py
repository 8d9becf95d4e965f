"""Task-level parity on the GPU: the product's launch programs on libxmeta.so vs the oracle
(oracle/maml_oracle.py, CPU, fp64 and fp32) on the same seeded synthetic tasks.

Tolerance contract: compare with the fp64 oracle; e_new = rel-L2(ours, fp64),
e_ref = rel-L2(oracle fp32, fp64); require e_new <= max(tau, 4*e_ref) with tau = 1e-4 for the query
loss / adapted weights and 1e-3 for the meta-gradient; the four conv.bias gradients (analytically zero)
are bounded absolutely; arg-max correct counts must match exactly."""
import pytest
import torch

from exploring_meta_b200 import engine as eng
from exploring_meta_b200 import spec as pspec
from exploring_meta_b200.synthetic import make_tasks
from oracle import maml_oracle as mo

pytestmark = pytest.mark.gpu

TAU_W, TAU_G = 1e-4, 1e-3


def _ospec(s):
    return mo.NetSpec(s.in_c, s.in_h, s.in_w, s.hidden, s.ways, s.layers, s.pool,
                      s.head if s.head != 'none' else 'flatten')


def _check_maml(spec, shots, steps, lr, tasks, seed, mode='second'):
    ospec = _ospec(spec)
    params = mo.init_params(ospec, seed=42)
    X, Y = make_tasks(tasks, spec.ways, shots, (spec.in_c, spec.in_h, spec.in_w), seed=seed)
    fo = mode == 'first'
    r64 = mo.meta_iteration([p.double() for p in params], X.double(), Y, ospec, steps, lr, first_order=fo)
    r32 = mo.meta_iteration(params, X, Y, ospec, steps, lr, first_order=fo)
    e = eng.MamlEngine(spec, tasks, shots, steps, lr, mode=mode, device='cuda')
    e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda())
    torch.cuda.synchronize()
    mask = ~mo.conv_bias_mask(ospec)
    g64, g32, g = mo.flatten(r64['grad']), mo.flatten(r32['grad']), e.grad.cpu()
    e_ref, e_new = mo.rel_l2(g32[mask], g64[mask]), mo.rel_l2(g[mask], g64[mask])
    assert e_new <= max(TAU_G, 4 * e_ref), 'meta-grad: e_new %.3e e_ref %.3e' % (e_new, e_ref)
    assert g[~mask].abs().max() <= 1e-5 * g64.abs().max()
    assert torch.allclose(e.loss.cpu().double(), r64['loss'], rtol=1e-4, atol=1e-5)
    assert e.correct.cpu().tolist() == r64['correct'].tolist()
    for t in range(tasks):
        th64, th32 = mo.flatten(r64['adapted'][t]), mo.flatten(r32['adapted'][t])
        ew_ref = mo.rel_l2(th32[mask], th64[mask])
        ew = mo.rel_l2(e.theta_steps[steps - 1, t].cpu()[mask], th64[mask])
        assert ew <= max(TAU_W, 4 * ew_ref), 'theta_T task %d: %.3e (ref %.3e)' % (t, ew, ew_ref)
    return e_new, e_ref


def test_omniglot_5w1s_config1_shape():
    """BASELINE config 1 shapes (OmniglotCNN 64 filters, 5-way 1-shot, 1 step, lr 0.5), 4 tasks."""
    _check_maml(pspec.omniglot_spec(5), 1, 1, 0.5, 4, seed=0)


def test_omniglot_20w5s_config4_shape():
    _check_maml(pspec.omniglot_spec(20), 5, 1, 0.5, 2, seed=1)


def test_miniimagenet_5w5s_one_step():
    _check_maml(pspec.miniimagenet_spec(5), 5, 1, 0.5, 2, seed=2)


def test_miniimagenet_5w1s_three_steps_calm_lr():
    """Config 2 shape at the calm inner lr (at lr 0.5 / 5 steps the reference itself
    deviates from fp64 by 0.27 on random-init data, so the tight check runs at a small lr)."""
    _check_maml(pspec.miniimagenet_spec(5), 1, 3, 0.001, 2, seed=3)


def test_miniimagenet_5w5s_five_steps_config2_shape_calm_lr():
    """BASELINE config 2 EXACTLY (S = 25 support rows, T = 5, second order) at the calm inner lr: a second seed
    next to the committed golden ``maml_min_5w5s_t5_calm`` (there e_ref = 2.6e-5; here the reference's own fp32
    run has e_ref = 3.8e-4, which the contract's 4 * e_ref term absorbs)."""
    e_new, e_ref = _check_maml(pspec.miniimagenet_spec(5), 5, 5, 0.001, 2, seed=14)
    print('cfg-2 shape, calm lr: e_new %.3e e_ref %.3e' % (e_new, e_ref))


def test_config2_shape_decision_flips_no_more_frequent_than_reference_fp32():
    """At S = 25, T = 5 a single task evaluates 5 x 4 layers x up to 1.4 M ReLU / max-pool decisions; a pre-activation
    pair that ties to fp32 rounding sends the gradient to a different winner and moves the meta-gradient by 1e-4..4e-2
    -- in the reference's OWN fp32 run as often as in ours (profiles/r02_flip_rate.txt: 11 vs 10 of 16 seeds).  Per
    seed the tolerance contract therefore cannot be tight; what must hold: the query loss and the correct count agree
    with fp64 on every seed, seeds where neither run flips agree to rounding, and over 12 seeds our flip count does not
    exceed the reference's by more than four (which seeds flip depends on the last bit of every reduction: a
    different summation order in ONE kernel moves individual seeds in and out of the set; measured 10 vs 11 of 16 and
    11 vs 8 of 12 with two versions of the head kernel).  The exact per-seed statements are the loss, the count and the
    calm seeds."""
    spec, ospec = pspec.miniimagenet_spec(5), _ospec(pspec.miniimagenet_spec(5))
    params = mo.init_params(ospec, seed=42)
    mask = ~mo.conv_bias_mask(ospec)
    e = eng.MamlEngine(spec, 1, 5, 5, 0.001, mode='second', device='cuda')
    flips_ref = flips_new = 0
    for seed in range(100, 112):
        X, Y = make_tasks(1, 5, 5, (3, 84, 84), seed=seed)
        r64 = mo.meta_iteration([p.double() for p in params], X.double(), Y, ospec, 5, 0.001)
        r32 = mo.meta_iteration(params, X, Y, ospec, 5, 0.001)
        e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda())
        g64 = mo.flatten(r64['grad'])[mask]
        e_ref, e_new = mo.rel_l2(mo.flatten(r32['grad'])[mask], g64), mo.rel_l2(e.grad.cpu()[mask], g64)
        assert torch.allclose(e.loss.cpu().double(), r64['loss'], rtol=1e-4, atol=1e-5)
        assert e.correct.cpu().tolist() == r64['correct'].tolist()
        flips_ref += e_ref > 1e-4
        flips_new += e_new > 1e-4
        if e_ref <= 1e-4 and e_new <= 1e-4:
            assert e_new <= max(2e-5, 4 * e_ref)
        assert e_new <= 0.1                      # a flip, not a wrong kernel: the largest observed is 3.6e-2
    print('decision flips over 12 seeds: reference fp32 %d, CUDA path %d' % (flips_ref, flips_new))
    assert flips_new <= flips_ref + 4, (flips_new, flips_ref)


def test_omniglot_20w5s_config4_four_tasks():
    _check_maml(pspec.omniglot_spec(20), 5, 1, 0.5, 4, seed=21)


def test_first_order_and_eval_modes():
    spec = pspec.miniimagenet_spec(5)
    _check_maml(spec, 1, 2, 0.01, 2, seed=4, mode='first')
    ospec = _ospec(spec)
    params = mo.init_params(ospec, seed=42)
    X, Y = make_tasks(2, 5, 1, (3, 84, 84), seed=4)
    r = mo.meta_iteration([p.double() for p in params], X.double(), Y, ospec, 2, 0.01, first_order=True)
    e = eng.MamlEngine(spec, 2, 1, 2, 0.01, mode='eval', device='cuda')
    e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda())
    assert torch.allclose(e.loss.cpu().double(), r['loss'], rtol=1e-4, atol=1e-5)
    assert e.correct.cpu().tolist() == r['correct'].tolist()


def test_cuda_graph_replay_is_deterministic_enough():
    spec = pspec.omniglot_spec(5)
    params = mo.init_params(_ospec(spec), seed=42)
    X, Y = make_tasks(4, 5, 1, (1, 28, 28), seed=0)
    e = eng.MamlEngine(spec, 4, 1, 1, 0.5, device='cuda')
    e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda())
    g0 = e.grad.clone()
    e.capture()
    e.launch()
    torch.cuda.synchronize()
    assert mo.rel_l2(e.grad.cpu(), g0.cpu()) < 1e-5


def test_bn_running_stats():
    spec = pspec.omniglot_spec(5)
    ospec = _ospec(spec)
    params = mo.init_params(ospec, seed=42)
    X, Y = make_tasks(3, 5, 1, (1, 28, 28), seed=5)
    ref = mo.meta_iteration(params, X, Y, ospec, 1, 0.5)
    rm_ref, rv_ref = mo.compose_running_stats([torch.zeros(64)] * 4, [torch.ones(64)] * 4, ref['bn_calls'])
    e = eng.MamlEngine(spec, 3, 1, 1, 0.5, device='cuda')
    e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda())
    rm = [torch.zeros(64, device='cuda') for _ in range(4)]
    rv = [torch.ones(64, device='cuda') for _ in range(4)]
    e.update_running_stats(rm, rv)
    for l in range(4):
        assert torch.allclose(rm[l].cpu(), rm_ref[l], rtol=1e-4, atol=1e-5)
        assert torch.allclose(rv[l].cpu(), rv_ref[l], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('first_order,tasks', [(False, 2), (True, 2), (False, 4)])
def test_anil_config3_shape(first_order, tasks):
    spec = pspec.anil_body_spec('min', 5)
    ospec = _ospec(spec)
    body = mo.init_params(ospec, seed=42, with_head=False)
    torch.manual_seed(7)
    head = [torch.randn(5, 1600) * 0.03, torch.zeros(5)]
    X, Y = make_tasks(tasks, 5, 5, (3, 84, 84), seed=6)
    r64 = mo.meta_iteration([p.double() for p in body], X.double(), Y, ospec, 1, 0.5, first_order=first_order,
                            anil_head=[h.double() for h in head])
    r32 = mo.meta_iteration(body, X, Y, ospec, 1, 0.5, first_order=first_order, anil_head=head)
    e = eng.AnilEngine(spec, tasks, 5, 1, 0.5, first_order=first_order, device='cuda')
    e.run(X.cuda(), Y.cuda(), mo.flatten(body).cuda(), mo.flatten(head).cuda())
    mask = ~mo.conv_bias_mask(ospec, with_head=False)
    g64, g32 = mo.flatten(r64['grad']), mo.flatten(r32['grad'])
    e_ref, e_new = mo.rel_l2(g32[mask], g64[mask]), mo.rel_l2(e.grad.cpu()[mask], g64[mask])
    assert e_new <= max(TAU_G, 4 * e_ref), 'body grad: %.3e (ref %.3e)' % (e_new, e_ref)
    h64 = mo.flatten(r64['head_grad'])
    assert mo.rel_l2(e.head_grad.cpu(), h64) <= max(TAU_G, 4 * mo.rel_l2(mo.flatten(r32['head_grad']), h64))
    assert torch.allclose(e.loss.cpu().double(), r64['loss'], rtol=1e-4, atol=1e-5)
    assert e.correct.cpu().tolist() == r64['correct'].tolist()
