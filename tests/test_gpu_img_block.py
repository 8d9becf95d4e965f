"""Fused image block (xm_img_*: first ConvBlock without the pre-BN map, Gram-matrix closed forms + sparse winner
gather) on the GPU against the dense evaluation of the same contract in tests/cabi_emulator.py."""
import ctypes

import pytest
import torch

from exploring_meta_b200 import _lib
from exploring_meta_b200._lib import XmBlockGeom, XmImgArgs
from test_gpu_kernels import Pair

pytestmark = pytest.mark.gpu

CASES = [
    # tasks, n, cin, cout, H, W, shared weights
    (2, 3, 3, 32, 12, 16, False),
    (3, 2, 3, 64, 84, 84, False),     # Mini-ImageNet resolution, ANIL body width (two channel tiles)
    (2, 2, 1, 32, 28, 28, False),
    (2, 3, 4, 32, 10, 14, False),     # W % 4 == 2: half-filled last column group
    (33, 2, 3, 32, 30, 30, True),     # many tasks, master weights shared (task stride 0)
    (2, 25, 3, 32, 84, 84, True),     # config-2 shape per task
]


def _geom(tasks, n, cin, cout, H, W):
    return XmBlockGeom(tasks, n, cin, cout, H, W, H, W, H // 2, W // 2, 1, 1)


def _setup(case, seed):
    tasks, n, cin, cout, H, W, shared = case
    g = _geom(tasks, n, cin, cout, H, W)
    K = 9 * cin
    Pn = 3 * cout + cout * K                    # [gamma, beta, w, b] per task
    torch.manual_seed(seed)
    P = Pair()
    rows = 2 * n
    P.add('x', torch.randn(tasks, rows, cin, H, W) + 0.3)
    wt = 1 if shared else tasks
    th = torch.zeros(wt, Pn)
    th[:, :cout] = torch.rand(wt, cout) + 0.1
    th[:, cout:2 * cout] = torch.randn(wt, cout) * 0.3 - 0.5
    th[:, 2 * cout:2 * cout + cout * K] = torch.randn(wt, cout * K) * 0.3
    P.add('theta', th)
    P.add('v', torch.randn(tasks, Pn) * 0.5)                       # tangent direction / axpy base
    P.out('gram', (tasks, K * K + K), torch.float64)
    pshape = (tasks, n, H // 2, W // 2, cout)
    P.out('p', pshape).out('zsel', pshape).out('sel', pshape, torch.uint8)
    P.out('pdot', pshape).out('zdsel', pshape)
    P.add('gp', torch.randn(*pshape)).add('gpd', torch.randn(*pshape))
    for name in ('mi', 'cs', 'br', 'dr'):
        P.out(name, (tasks, 2, cout))
    P.out('ssum', (tasks, cout, K + 3), torch.float64).out('scratch', (tasks, cout, K + 3), torch.float64)
    P.out('out', (tasks, Pn))
    return g, P, K, Pn, (0 if shared else Pn), rows


def _args(g, P, ptr, K, Pn, tstride, rows):
    cout = g.cout
    a = XmImgArgs()
    a.g, a.eps = g, 1e-5
    a.row0, a.row_step, a.rows_per_task = 1, 2, rows
    a.x, a.gram = ptr('x'), ptr('gram')
    a.gamma, a.beta, a.gb_task_stride = ptr('theta'), ptr('theta', cout), tstride
    a.w, a.w_task_stride = ptr('theta', 2 * cout), tstride
    a.mean_invstd, a.scratch = ptr('mi'), ptr('scratch')
    a.zsel, a.sel = ptr('zsel'), ptr('sel')
    return a


def _dual(a, ptr, cout, Pn):
    a.gamma_dot, a.beta_dot, a.gbdot_task_stride = ptr('v'), ptr('v', cout), Pn
    a.w_dot, a.wdot_task_stride = ptr('v', 2 * cout), Pn


def _outs(a, ptr, cout, K, Pn, base):
    a.out_gamma, a.out_beta, a.out_w, a.out_b = ptr('out'), ptr('out', cout), ptr('out', 2 * cout), ptr('out', 2 * cout + cout * K)
    a.out_task_stride = Pn
    if base:
        a.base_gamma, a.base_beta = ptr('v'), ptr('v', cout)
        a.base_w, a.base_b = ptr('v', 2 * cout), ptr('v', 2 * cout + cout * K)
        a.base_task_stride = Pn
    a.scale = -0.37


def _sync(P, *names):
    for nme in names:
        P.gpu[nme].copy_(P.cpu[nme])


@pytest.mark.parametrize('case', CASES)
def test_image_block_chain(case):
    g, P, K, Pn, ts, rows = _setup(case, 0)
    cout = g.cout
    lib = _lib.load()
    assert lib.xm_img_supported(ctypes.byref(g)) == 1
    assert lib.xm_img_gram_bytes(ctypes.byref(g)) == P.cpu['gram'].numel() * 8
    assert lib.xm_img_scratch_bytes(ctypes.byref(g)) == P.cpu['scratch'].numel() * 8

    P.run('xm_img_gram', lambda ptr: _args(g, P, ptr, K, Pn, ts, rows))
    P.close('gram', 1e-12, 0)

    def fwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        a.call_stats, a.p = ptr('cs'), ptr('p')
        return a
    P.run('xm_img_fwd', fwd)
    P.close('mi', 2e-6)
    P.close('cs', 2e-6)
    P.close('p', 2e-6)
    P.close('zsel', 2e-6)
    mism = (P.gpu['sel'].cpu() != P.cpu['sel']).sum().item()
    assert mism <= max(2, P.cpu['sel'].numel() // 200000), 'winner positions differ in %d elements' % mism
    dead = (P.cpu['sel'] == 255).float().mean().item()
    assert 0.02 < dead < 0.98                                   # the case exercises both live and ReLU-dead windows
    _sync(P, 'gram', 'mi', 'zsel', 'sel')

    def bwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        a.gp, a.bwd_red, a.ssum = ptr('gp'), ptr('br'), ptr('ssum')
        _outs(a, ptr, cout, K, Pn, base=True)
        return a
    P.run('xm_img_bwd', bwd)
    P.close('br', 5e-6)
    P.close('ssum', 5e-6)
    P.close('out', 5e-6)
    _sync(P, 'br', 'ssum')

    def dfwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        _dual(a, ptr, cout, Pn)
        a.dual_red, a.pdot, a.zdsel = ptr('dr'), ptr('pdot'), ptr('zdsel')
        return a
    P.run('xm_img_dual_fwd', dfwd)
    P.close('dr', 5e-6)
    P.close('pdot', 5e-6)
    P.close('zdsel', 5e-6)
    _sync(P, 'dr', 'zdsel')

    def dbwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        _dual(a, ptr, cout, Pn)
        a.gp, a.gpdot, a.bwd_red, a.dual_red = ptr('gp'), ptr('gpd'), ptr('br'), ptr('dr')
        a.zdsel, a.ssum = ptr('zdsel'), ptr('ssum')
        _outs(a, ptr, cout, K, Pn, base=True)
        return a
    P.out('out', (g.tasks, Pn))
    P.run('xm_img_dual_bwd', dbwd)
    P.close('out', 1e-5)


def test_image_block_no_base_no_gpdot():
    """NULL base (plain gradient out = scale * g) and NULL gpdot (tangent of the cotangent is zero)."""
    case = CASES[0]
    g, P, K, Pn, ts, rows = _setup(case, 3)
    cout = g.cout
    P.run('xm_img_gram', lambda ptr: _args(g, P, ptr, K, Pn, ts, rows))

    def fwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        a.p = ptr('p')
        return a
    P.run('xm_img_fwd', fwd)
    _sync(P, 'gram', 'mi', 'zsel', 'sel')

    def bwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        a.gp, a.bwd_red, a.ssum = ptr('gp'), ptr('br'), ptr('ssum')
        _outs(a, ptr, cout, K, Pn, base=False)
        return a
    P.run('xm_img_bwd', bwd)
    P.close('out', 5e-6)
    _sync(P, 'br', 'ssum')

    def dfwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        _dual(a, ptr, cout, Pn)
        a.dual_red, a.pdot, a.zdsel = ptr('dr'), ptr('pdot'), ptr('zdsel')
        return a
    P.run('xm_img_dual_fwd', dfwd)
    _sync(P, 'dr', 'zdsel')

    def dbwd(ptr):
        a = _args(g, P, ptr, K, Pn, ts, rows)
        _dual(a, ptr, cout, Pn)
        a.gp, a.bwd_red, a.dual_red = ptr('gp'), ptr('br'), ptr('dr')
        a.zdsel, a.ssum = ptr('zdsel'), ptr('ssum')
        _outs(a, ptr, cout, K, Pn, base=False)
        return a
    P.run('xm_img_dual_bwd', dbwd)
    P.close('out', 1e-5)


def test_image_block_rejects_uncovered_geometry():
    lib = _lib.load()
    bad = [XmBlockGeom(2, 3, 3, 32, 21, 21, 21, 21, 10, 10, 1, 1),     # odd map
           XmBlockGeom(2, 3, 3, 64, 28, 28, 14, 14, 14, 14, 2, 0),     # stride 2 with three input channels
           XmBlockGeom(2, 3, 32, 32, 42, 42, 42, 42, 21, 21, 1, 1),    # not an image layer
           XmBlockGeom(2, 3, 3, 8, 12, 12, 12, 12, 6, 6, 1, 1)]        # narrow
    for g in bad:
        assert lib.xm_img_supported(ctypes.byref(g)) == 0
        a = XmImgArgs()
        a.g = g
        assert lib.xm_img_fwd(ctypes.byref(a), None) < 0
        assert b'image-block' in lib.xm_last_error()


# ---- Omniglot image block: stride 2, no pool, one input channel (img_flat.cu) --------------------------------------
FLAT_CASES = [
    # tasks, n, cout, H, shared weights
    (2, 5, 64, 28, False),
    (3, 4, 32, 28, False),
    (40, 3, 64, 14, True),
    (2, 100, 64, 28, True),          # config-4 shape per task
]


@pytest.mark.parametrize('case', FLAT_CASES)
def test_flat_image_block_chain(case):
    tasks, n, cout, H, shared = case
    hz = (H + 2 - 3) // 2 + 1
    g = XmBlockGeom(tasks, n, 1, cout, H, H, hz, hz, hz, hz, 2, 0)
    K, Pn = 9, 3 * cout + cout * 9
    lib = _lib.load()
    assert lib.xm_img_supported(ctypes.byref(g)) == 1
    torch.manual_seed(7)
    P = Pair()
    rows = 2 * n
    P.add('x', torch.rand(tasks, rows, 1, H, H))
    wt = 1 if shared else tasks
    th = torch.zeros(wt, Pn)
    th[:, :cout] = torch.rand(wt, cout) + 0.1
    th[:, cout:2 * cout] = torch.randn(wt, cout) * 0.3
    th[:, 2 * cout:2 * cout + cout * K] = torch.randn(wt, cout * K) * 0.5
    P.add('theta', th)
    P.add('v', torch.randn(tasks, Pn) * 0.5)
    P.out('gram', (tasks, K * K + K), torch.float64)
    pshape = (tasks, n, hz, hz, cout)
    P.out('p', pshape).out('pdot', pshape)
    P.add('gp', torch.randn(*pshape)).add('gpd', torch.randn(*pshape))
    for name in ('mi', 'cs', 'br', 'dr'):
        P.out(name, (tasks, 2, cout))
    P.out('ssum', (tasks, cout, K + 3), torch.float64).out('scratch', (tasks, cout, K + 3), torch.float64)
    P.out('out', (tasks, Pn))
    ts = 0 if shared else Pn

    def base(ptr):
        a = XmImgArgs()
        a.g, a.eps = g, 1e-5
        a.row0, a.row_step, a.rows_per_task = 0, 2, rows
        a.x, a.gram = ptr('x'), ptr('gram')
        a.gamma, a.beta, a.gb_task_stride = ptr('theta'), ptr('theta', cout), ts
        a.w, a.w_task_stride = ptr('theta', 2 * cout), ts
        a.mean_invstd, a.scratch = ptr('mi'), ptr('scratch')
        return a

    P.run('xm_img_gram', base)
    P.close('gram', 1e-12, 0)

    def fwd(ptr):
        a = base(ptr)
        a.call_stats, a.p = ptr('cs'), ptr('p')
        return a
    P.run('xm_img_fwd', fwd)
    P.close('mi', 2e-6)
    P.close('cs', 2e-6)
    P.close('p', 2e-6)
    alive = (P.cpu['p'] > 0).float().mean().item()
    assert 0.05 < alive < 0.95
    _sync(P, 'gram', 'mi')

    def bwd(ptr):
        a = base(ptr)
        a.gp, a.bwd_red, a.ssum = ptr('gp'), ptr('br'), ptr('ssum')
        _outs(a, ptr, cout, K, Pn, base=True)
        return a
    P.run('xm_img_bwd', bwd)
    P.close('br', 5e-6)
    P.close('ssum', 5e-6)
    P.close('out', 5e-6)
    _sync(P, 'br', 'ssum')

    def dfwd(ptr):
        a = base(ptr)
        _dual(a, ptr, cout, Pn)
        a.dual_red, a.pdot = ptr('dr'), ptr('pdot')
        return a
    P.run('xm_img_dual_fwd', dfwd)
    P.close('dr', 5e-6)
    P.close('pdot', 5e-6)
    _sync(P, 'dr')

    def dbwd(ptr):
        a = base(ptr)
        _dual(a, ptr, cout, Pn)
        a.gp, a.gpdot, a.bwd_red, a.dual_red, a.ssum = ptr('gp'), ptr('gpd'), ptr('br'), ptr('dr'), ptr('ssum')
        _outs(a, ptr, cout, K, Pn, base=True)
        return a
    P.out('out', (tasks, Pn))
    P.run('xm_img_dual_bwd', dbwd)
    P.close('out', 1e-5)
