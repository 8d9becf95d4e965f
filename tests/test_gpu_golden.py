"""The CUDA path (through the C ABI) against the committed golden vectors of the reference's own files
(tests/golden/*.npz, made by tests/golden/make_golden.py).

Tolerance contract: compare with the reference's fp64 result; e_ref = the
reference's own fp32-vs-fp64 deviation stored in the fixture; require e_new <= max(tau, 4 * e_ref) with
tau = 1e-3 for meta-gradients, 1e-4 for adapted weights and query losses; conv.bias gradients (analytically
zero) bounded absolutely; correct counts exact (on the chaotic headline case: equal to the fp64 OR the fp32
reference count)."""
import pytest
import torch

import golden_util as gu
from exploring_meta_b200 import engine as eng
from exploring_meta_b200 import spec as pspec
from oracle import maml_oracle as mo

pytestmark = pytest.mark.gpu
TAU_W, TAU_G = 1e-4, 1e-3


def _pspec(g):
    o = g.ospec()
    head = 'none' if g.algo == 'anil' else o.head
    return pspec.NetSpec(o.in_c, o.in_h, o.in_w, o.hidden, o.ways, o.layers, o.pool, head)


@pytest.mark.parametrize('name', gu.NAMES)
def test_cuda_path_matches_golden(name):
    g = gu.Golden(name)
    X, Y, params, head = g.inputs()
    mask = g.grad_mask()
    g64 = g.t('grad64_as_f32')
    e_ref = float(g.z['e_ref_grad'])
    if g.algo == 'maml':
        e = eng.MamlEngine(_pspec(g), g.tasks, g.shots, g.steps, g.inner_lr, mode='second', device='cuda')
        e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda())
    else:
        e = eng.AnilEngine(_pspec(g), g.tasks, g.shots, g.steps, g.inner_lr, device='cuda')
        e.run(X.cuda(), Y.cuda(), mo.flatten(params).cuda(), mo.flatten(head).cuda())
    torch.cuda.synchronize()
    grad = e.grad.cpu()
    e_new = mo.rel_l2(grad[mask], g64[mask])
    assert e_new <= max(TAU_G, 4 * e_ref), '%s meta-grad: e_new %.3e vs e_ref %.3e' % (name, e_new, e_ref)
    if (~mask).any():
        assert grad[~mask].abs().max() <= 1e-5 * g64.abs().max()
    loss64, loss32 = g.t('loss64'), g.t('loss32').double()
    tol = torch.clamp(4 * (loss32 - loss64).abs(), min=0) + 1e-4 * loss64.abs() + 1e-5
    assert ((e.loss.cpu().double() - loss64).abs() <= tol).all(), (e.loss.cpu().tolist(), loss64.tolist())
    got = e.correct.cpu().tolist()
    if 'headline' in name:
        assert all(c in (a, b) for c, a, b in zip(got, g.t('correct').tolist(), g.t('correct32').tolist()))
    else:
        assert got == g.t('correct').tolist()
    if g.algo == 'maml':
        th = e.theta_steps[g.steps - 1, 0].cpu()
        ew = mo.rel_l2(th[mask], g.t('adapted0_64_as_f32')[mask])
        assert ew <= max(TAU_W, 4 * float(g.z['e_ref_adapted0'])), 'theta_T: %.3e' % ew
    else:
        eh = mo.rel_l2(e.head_grad.cpu(), g.t('head_grad64_as_f32'))
        assert eh <= max(TAU_G, 4 * float(g.z['e_ref_head_grad'])), 'head grad: %.3e' % eh
    print('%s: e_new %.3e e_ref %.3e' % (name, e_new, e_ref))
