"""Host-side launch programs (exploring_meta_b200/engine.py) run against the CPU emulator of the C ABI
and compared with the oracle: checks the forward-over-reverse algorithm, buffer plumbing and launch
order without a GPU.  (The -m gpu twins in test_gpu_parity.py run the same programs on libxmeta.so.)"""
import pytest
import torch

from exploring_meta_b200 import engine as eng
from exploring_meta_b200 import spec as pspec
from exploring_meta_b200.synthetic import make_tasks
from oracle import maml_oracle as mo


def _ospec(s):
    return mo.NetSpec(s.in_c, s.in_h, s.in_w, s.hidden, s.ways, s.layers, s.pool,
                      s.head if s.head != 'none' else 'flatten')


CASES = [
    # (spec, shots, steps, lr, tasks)
    (pspec.NetSpec(3, 12, 12, 8, 3, 2, True, 'flatten'), 2, 2, 0.1, 3),
    (pspec.NetSpec(1, 14, 14, 8, 4, 3, False, 'mean'), 1, 1, 0.4, 2),
    (pspec.NetSpec(2, 21, 21, 4, 2, 4, True, 'flatten'), 1, 3, 0.05, 2),   # odd sizes: floor pooling
    (pspec.NetSpec(3, 12, 16, 32, 3, 2, True, 'flatten'), 2, 2, 0.1, 2),   # fused image-block path (xm_img_*)
    (pspec.NetSpec(1, 14, 14, 32, 4, 3, False, 'mean'), 1, 2, 0.3, 2),     # Omniglot image block (stride 2, no pool)
]


@pytest.mark.parametrize('spec,shots,steps,lr,tasks', CASES)
def test_maml_second_order_matches_oracle(emulated_lib, spec, shots, steps, lr, tasks):
    ospec = _ospec(spec)
    params = mo.init_params(ospec, seed=3)
    X, Y = make_tasks(tasks, spec.ways, shots, (spec.in_c, spec.in_h, spec.in_w), seed=1)
    ref = mo.meta_iteration([p.double() for p in params], X.double(), Y, ospec, steps, lr)
    e = eng.MamlEngine(spec, tasks, shots, steps, lr, mode='second', device='cpu')
    e.run(X, Y, mo.flatten(params))
    g_ref = mo.flatten(ref['grad'])
    mask = ~mo.conv_bias_mask(ospec)
    assert mo.rel_l2(e.grad[mask], g_ref[mask]) < 2e-4
    assert e.grad[~mask].abs().max() <= 1e-5 * g_ref.abs().max()
    assert torch.allclose(e.loss.double(), ref['loss'], rtol=1e-4, atol=1e-5)
    assert e.correct.tolist() == ref['correct'].tolist()
    for t in range(tasks):
        th = mo.flatten(ref['adapted'][t])
        assert mo.rel_l2(e.theta_steps[steps - 1, t][mask], th[mask]) < 1e-4


@pytest.mark.parametrize('case', [0, 3])
def test_maml_first_order_and_eval(emulated_lib, case):
    spec, shots, steps, lr, tasks = CASES[case]
    ospec = _ospec(spec)
    params = mo.init_params(ospec, seed=5)
    X, Y = make_tasks(tasks, spec.ways, shots, (spec.in_c, spec.in_h, spec.in_w), seed=2)
    ref = mo.meta_iteration([p.double() for p in params], X.double(), Y, ospec, steps, lr, first_order=True)
    e = eng.MamlEngine(spec, tasks, shots, steps, lr, mode='first', device='cpu')
    e.run(X, Y, mo.flatten(params))
    mask = ~mo.conv_bias_mask(ospec)
    assert mo.rel_l2(e.grad[mask], mo.flatten(ref['grad'])[mask]) < 2e-4
    e2 = eng.MamlEngine(spec, tasks, shots, steps, lr, mode='eval', device='cpu')
    e2.run(X, Y, mo.flatten(params))
    assert torch.allclose(e2.loss.double(), ref['loss'], rtol=1e-4, atol=1e-5)
    assert e2.correct.tolist() == ref['correct'].tolist()


def test_bn_running_stats(emulated_lib):
    spec, shots, steps, lr, tasks = CASES[0]
    ospec = _ospec(spec)
    params = mo.init_params(ospec, seed=5)
    X, Y = make_tasks(tasks, spec.ways, shots, (spec.in_c, spec.in_h, spec.in_w), seed=2)
    ref = mo.meta_iteration(params, X, Y, ospec, steps, lr)
    rm0 = [torch.zeros(spec.hidden) for _ in range(spec.layers)]
    rv0 = [torch.ones(spec.hidden) for _ in range(spec.layers)]
    rm_ref, rv_ref = mo.compose_running_stats(rm0, rv0, ref['bn_calls'])
    e = eng.MamlEngine(spec, tasks, shots, steps, lr, mode='second', device='cpu')
    e.run(X, Y, mo.flatten(params))
    rm = [t.clone() for t in rm0]
    rv = [t.clone() for t in rv0]
    n = e.update_running_stats(rm, rv)
    assert n == tasks * (steps + 1)
    for l in range(spec.layers):
        assert torch.allclose(rm[l], rm_ref[l], rtol=1e-4, atol=1e-5)
        assert torch.allclose(rv[l], rv_ref[l], rtol=1e-4, atol=1e-5)


def test_image_block_path_is_taken(emulated_lib):
    spec, shots, steps, lr, tasks = CASES[3]
    e = eng.MamlEngine(spec, tasks, shots, steps, lr, mode='second', device='cpu')
    names = [name for _fn, _a, name in e.prog.calls]
    assert e.img and names.count('xm_img_gram') == 2
    assert names.count('xm_img_fwd') == steps + 1 and names.count('xm_img_bwd') == steps + 1
    assert names.count('xm_img_dual_fwd') == steps and names.count('xm_img_dual_bwd') == steps
    e8 = eng.MamlEngine(CASES[0][0], tasks, shots, steps, lr, mode='second', device='cpu')
    assert not e8.img and not any(n.startswith('xm_img') for _f, _a, n in e8.prog.calls)
    eo = eng.MamlEngine(CASES[4][0], 2, 1, 2, 0.3, mode='second', device='cpu')
    assert eo.img and sum(n == 'xm_img_dual_bwd' for _f, _a, n in eo.prog.calls) == 2


@pytest.mark.parametrize('first_order', [False, True])
@pytest.mark.parametrize('hidden', [8, 32])
def test_anil_matches_oracle(emulated_lib, first_order, hidden):
    spec = pspec.NetSpec(3, 12, 12, hidden, 3, 2, True, 'none')
    shots, steps, lr, tasks = 2, 2, 0.3, 3
    ospec = _ospec(spec)
    body = mo.init_params(ospec, seed=7, with_head=False)
    torch.manual_seed(11)
    D = spec.hidden * 3 * 3
    head = [torch.randn(spec.ways, D) * 0.1, torch.randn(spec.ways) * 0.1]
    X, Y = make_tasks(tasks, spec.ways, shots, (3, 12, 12), seed=4)
    ref = mo.meta_iteration([p.double() for p in body], X.double(), Y, ospec, steps, lr,
                            first_order=first_order, anil_head=[h.double() for h in head])
    e = eng.AnilEngine(spec, tasks, shots, steps, lr, first_order=first_order, device='cpu')
    e.run(X, Y, mo.flatten(body), mo.flatten(head))
    mask = ~mo.conv_bias_mask(ospec, with_head=False)
    assert mo.rel_l2(e.grad[mask], mo.flatten(ref['grad'])[mask]) < 2e-4
    assert mo.rel_l2(e.head_grad, mo.flatten(ref['head_grad'])) < 2e-4
    assert torch.allclose(e.loss.double(), ref['loss'], rtol=1e-4, atol=1e-5)
    assert e.correct.tolist() == ref['correct'].tolist()
