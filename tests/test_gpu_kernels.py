"""Per-kernel parity on the GPU: every libxmeta entry point against the CPU emulator of its contract
(tests/cabi_emulator.py, plain torch ops) on the same random inputs, called through the C ABI."""
import ctypes

import pytest
import torch

import cabi_emulator as emu
from exploring_meta_b200 import _lib
from exploring_meta_b200._lib import (XmAnilHeadArgs, XmBlockGeom, XmBnArgs, XmConvArgs, XmHeadArgs, XmWgradArgs)

pytestmark = pytest.mark.gpu


def geom(tasks, n, cin, cout, hin, stride, pool):
    hz = (hin + 2 - 3) // stride + 1
    hp = hz // 2 if pool else hz
    return XmBlockGeom(tasks, n, cin, cout, hin, hin, hz, hz, hp, hp, stride, 1 if pool else 0)


class Pair:
    """The same named buffers on the CPU (for the emulator) and on the GPU (for the library)."""

    def __init__(self):
        self.cpu, self.gpu = {}, {}

    def add(self, name, t):
        self.cpu[name] = t.contiguous().clone()
        self.gpu[name] = t.contiguous().clone().cuda()
        return self

    def out(self, name, shape, dtype=torch.float32):
        return self.add(name, torch.full(shape, 7.0).to(dtype))

    def run(self, fn_name, make_args):
        """make_args(ptr) -> args struct, where ptr(name, offset_elems=0) -> address."""
        def ptr_of(store):
            def ptr(name, off=0):
                if name is None:
                    return None
                t = store[name]
                return t.data_ptr() + off * t.element_size()
            return ptr
        a_cpu = make_args(ptr_of(self.cpu))
        rc = getattr(emu.EmulatedLib(), fn_name)(ctypes.byref(a_cpu), None)
        assert rc == 0
        lib = _lib.load()
        a_gpu = make_args(ptr_of(self.gpu))
        rc = getattr(lib, fn_name)(ctypes.byref(a_gpu), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, fn_name)
        torch.cuda.synchronize()

    def close(self, name, rtol=1e-5, atol=1e-6):
        a, b = self.gpu[name].cpu().double(), self.cpu[name].double()
        scale = b.abs().max().clamp_min(1e-30)
        err = ((a - b).abs().max() / scale).item()
        assert err <= rtol + atol, '%s: max rel-to-max error %.3e' % (name, err)


CONV_CASES = [
    # tasks, n, cin, cout, hin, stride, pool
    (2, 3, 3, 32, 20, 1, True),      # image layer (NCHW source)
    (2, 3, 32, 32, 21, 1, True),     # odd map, floor pooling
    (1, 2, 64, 64, 10, 1, True),     # ANIL body width: two cin chunks, two cout tiles
    (2, 5, 1, 64, 28, 2, False),     # Omniglot image layer
    (2, 5, 64, 64, 14, 2, False),    # Omniglot stride-2 layer
    (1, 9, 64, 64, 4, 2, False),     # tiny maps: several images per tile
    (2, 2, 8, 12, 9, 1, True),       # unusual channel counts
    (2, 3, 32, 32, 42, 1, True),     # tcgen05 path: one tile per CTA
    (40, 3, 32, 32, 42, 1, True),    # tcgen05 path: ~15 tiles per CTA (pipeline phases wrap)
    (3, 25, 32, 32, 10, 1, True),    # tcgen05 path: small maps, many images per tile
    (2, 6, 32, 32, 5, 1, True),
    (2, 2, 3, 32, 84, 1, True),      # image layer at full Mini-ImageNet resolution (tcgen05 image variant)
    (33, 2, 3, 32, 30, 1, True),     # image layer, several tiles per CTA
]


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('dual', [False, True])
def test_conv_forward(case, dual):
    tasks, n, cin, cout, hin, stride, pool = case
    g = geom(tasks, n, cin, cout, hin, stride, pool)
    nchw = cin <= 3
    torch.manual_seed(0)
    P = Pair()
    rows = 2 * n
    if nchw:
        P.add('x', torch.randn(tasks, rows, cin, hin, hin))
    else:
        P.add('x', torch.randn(tasks, n, hin, hin, cin))
        P.add('xd', torch.randn(tasks, n, hin, hin, cin))
    P.add('w', torch.randn(tasks, cout, cin, 3, 3) * 0.2).add('wd', torch.randn(tasks, cout, cin, 3, 3) * 0.2)
    P.add('aux', torch.randn(tasks, n, g.hz, g.wz, cout))
    P.out('out', (tasks, n, g.hz, g.wz, cout)).out('stats', (tasks, 2, cout), torch.float64)
    ws_bytes = max(int(_lib.load().xm_conv_workspace_bytes(ctypes.byref(g))), 0)   # stride-2 layers on the tcgen05 path
    P.out('ws', (max(ws_bytes // 4, 1),))

    def mk(ptr):
        a = XmConvArgs()
        a.g, a.mode = g, 0
        if ws_bytes:
            a.workspace, a.workspace_bytes = ptr('ws'), ws_bytes
        a.stat_mode = 2 if dual else 1
        if nchw:
            a.src_nchw, a.row0, a.row_step, a.rows_per_task = 1, 1, 2, rows
        a.src1, a.w1, a.w1_task_stride = ptr('x'), ptr('w'), cout * cin * 9
        if dual and not nchw:
            a.src2, a.w2, a.w2_task_stride = ptr('xd'), ptr('wd'), cout * cin * 9
        a.out, a.aux, a.stats = ptr('out'), ptr('aux'), ptr('stats')
        return a
    P.run('xm_conv', mk)
    P.close('out', 2e-6)
    P.close('stats', 1e-5)


@pytest.mark.parametrize('case', [c for c in CONV_CASES if c[2] > 3])
@pytest.mark.parametrize('dual', [False, True])
def test_conv_dgrad(case, dual):
    tasks, n, cin, cout, hin, stride, pool = case
    g = geom(tasks, n, cin, cout, hin, stride, pool)
    torch.manual_seed(1)
    P = Pair()
    P.add('gz', torch.randn(tasks, n, g.hz, g.wz, cout)).add('gzd', torch.randn(tasks, n, g.hz, g.wz, cout))
    P.add('w', torch.randn(tasks, cout, cin, 3, 3) * 0.2).add('wd', torch.randn(tasks, cout, cin, 3, 3) * 0.2)
    P.out('out', (tasks, n, hin, hin, cin))
    ws_bytes = max(int(_lib.load().xm_conv_workspace_bytes(ctypes.byref(g))), 0)
    P.out('ws', (max(ws_bytes // 4, 1),))

    def mk(ptr):
        a = XmConvArgs()
        a.g, a.mode, a.stat_mode = g, 1, 0
        if ws_bytes:
            a.workspace, a.workspace_bytes = ptr('ws'), ws_bytes
        a.src1, a.w1, a.w1_task_stride = ptr('gz'), ptr('w'), cout * cin * 9
        if dual:
            a.src2, a.w2, a.w2_task_stride = ptr('gzd'), ptr('wd'), cout * cin * 9
        a.out = ptr('out')
        return a
    P.run('xm_conv', mk)
    P.close('out', 2e-6)


def test_conv_shared_weights_stride0():
    g = geom(3, 2, 32, 32, 10, 1, True)
    torch.manual_seed(2)
    P = Pair()
    P.add('x', torch.randn(3, 2, 10, 10, 32)).add('w', torch.randn(32, 32, 3, 3) * 0.2)
    P.out('out', (3, 2, 10, 10, 32)).out('stats', (3, 2, 32), torch.float64)

    def mk(ptr):
        a = XmConvArgs()
        a.g, a.mode, a.stat_mode = g, 0, 1
        a.src1, a.w1, a.w1_task_stride = ptr('x'), ptr('w'), 0
        a.out, a.stats = ptr('out'), ptr('stats')
        return a
    P.run('xm_conv', mk)
    P.close('out', 2e-6)


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('dual', [False, True])
def test_wgrad(case, dual):
    tasks, n, cin, cout, hin, stride, pool = case
    g = geom(tasks, n, cin, cout, hin, stride, pool)
    nchw = cin <= 3
    if nchw and dual:
        pytest.skip('image sources carry no tangent')
    torch.manual_seed(3)
    P = Pair()
    rows = 2 * n
    if nchw:
        P.add('x', torch.randn(tasks, rows, cin, hin, hin))
    else:
        P.add('x', torch.randn(tasks, n, hin, hin, cin)).add('xd', torch.randn(tasks, n, hin, hin, cin))
    P.add('g', torch.randn(tasks, n, g.hz, g.wz, cout)).add('gd', torch.randn(tasks, n, g.hz, g.wz, cout))
    Pn = cout * cin * 9 + cout
    P.add('base', torch.randn(tasks, Pn)).out('outp', (tasks, Pn))
    lib = _lib.load()
    nbytes = int(lib.xm_wgrad_scratch_bytes(ctypes.byref(g)))
    assert nbytes > 0
    P.out('partial', (nbytes // 4,))

    def mk(ptr):
        a = XmWgradArgs()
        a.g = g
        if nchw:
            a.src_nchw, a.row0, a.row_step, a.rows_per_task = 1, 0, 2, rows
        a.x1, a.g1 = ptr('x'), ptr('g')
        if dual:
            a.x1, a.g1, a.x2, a.g2 = ptr('x'), ptr('gd'), ptr('xd'), ptr('g')
        a.out_w, a.out_b, a.out_task_stride = ptr('outp'), ptr('outp', cout * cin * 9), Pn
        a.base_w, a.base_b, a.base_task_stride = ptr('base'), ptr('base', cout * cin * 9), Pn
        a.scale = -0.37
        a.partial, a.partial_bytes = ptr('partial'), nbytes
        return a
    P.run('xm_wgrad', mk)
    P.close('outp', 3e-6)


BN_CASES = [
    (2, 3, 32, 20, 1, True), (2, 3, 32, 21, 1, True), (1, 2, 64, 10, 1, True),
    (2, 5, 64, 14, 2, False), (3, 4, 64, 4, 2, False), (2, 2, 12, 9, 1, True), (2, 2, 6, 9, 1, True),
]


def _bn_buffers(case, seed):
    tasks, n, C, hin, stride, pool = case
    g = geom(tasks, n, 4, C, hin, stride, pool)
    torch.manual_seed(seed)
    P = Pair()
    z = torch.randn(tasks, n, g.hz, g.wz, C) * 1.5 + 0.3
    zd = torch.randn(tasks, n, g.hz, g.wz, C)
    P.add('z', z).add('zd', zd)
    sums = torch.stack([z.double().sum(dim=(1, 2, 3)), (z.double() ** 2).sum(dim=(1, 2, 3))], dim=1)
    dsums = torch.stack([zd.double().sum(dim=(1, 2, 3)), (zd.double() * z.double()).sum(dim=(1, 2, 3))], dim=1)
    P.add('sums', sums).add('dsums', dsums)
    P.add('theta', torch.rand(tasks, 2 * C) + 0.1).add('v', torch.randn(tasks, 2 * C))
    P.add('base', torch.randn(tasks, 2 * C)).out('outp', (tasks, 2 * C))
    P.out('mi', (tasks, 2, C)).out('cs', (tasks, 2, C)).out('br', (tasks, 2, C)).out('dr', (tasks, 2, C))
    P.out('p', (tasks, n, g.hp, g.wp, C)).out('pd', (tasks, n, g.hp, g.wp, C))
    P.add('gp', torch.randn(tasks, n, g.hp, g.wp, C)).add('gpd', torch.randn(tasks, n, g.hp, g.wp, C))
    P.out('gz', (tasks, n, g.hz, g.wz, C)).out('gzd', (tasks, n, g.hz, g.wz, C))
    P.out('scratch', (tasks * 4 * C,), torch.float64)
    return g, P, C


def _bn_args(g, C, ptr, gpdot=True):
    b = XmBnArgs()
    b.g, b.eps = g, 1e-5
    b.z, b.zdot, b.sums, b.dsums = ptr('z'), ptr('zd'), ptr('sums'), ptr('dsums')
    b.gamma, b.beta, b.gb_task_stride = ptr('theta'), ptr('theta', C), 2 * C
    b.gamma_dot, b.beta_dot, b.gbdot_task_stride = ptr('v'), ptr('v', C), 2 * C
    b.mean_invstd, b.call_stats, b.bwd_red, b.dual_red = ptr('mi'), ptr('cs'), ptr('br'), ptr('dr')
    b.p, b.pdot, b.gp = ptr('p'), ptr('pd'), ptr('gp')
    b.gpdot = ptr('gpd') if gpdot else None
    b.gz, b.gzdot = ptr('gz'), ptr('gzd')
    b.out_gamma, b.out_beta, b.out_task_stride = ptr('outp'), ptr('outp', C), 2 * C
    b.base_gamma, b.base_beta, b.base_task_stride = ptr('base'), ptr('base', C), 2 * C
    b.scale = -0.4
    b.scratch = ptr('scratch')
    return b


@pytest.mark.parametrize('case', BN_CASES)
def test_bn_chain(case):
    """fwd -> bwd -> dual_fwd -> dual_bwd, each consuming the side buffers of the previous call."""
    g, P, C = _bn_buffers(case, 5)
    P.run('xm_bn_fwd', lambda ptr: _bn_args(g, C, ptr))
    for name in ('mi', 'cs', 'p'):
        P.close(name, 3e-6)
    P.run('xm_bn_bwd', lambda ptr: _bn_args(g, C, ptr))
    for name in ('br', 'gz', 'outp'):
        P.close(name, 1e-5)
    P.run('xm_bn_dual_fwd', lambda ptr: _bn_args(g, C, ptr))
    for name in ('dr', 'pd'):
        P.close(name, 1e-5)
    P.run('xm_bn_dual_bwd', lambda ptr: _bn_args(g, C, ptr))
    for name in ('gz', 'gzd', 'outp'):
        P.close(name, 2e-5)
    P.run('xm_bn_dual_bwd', lambda ptr: _bn_args(g, C, ptr, gpdot=False))
    for name in ('gz', 'gzd', 'outp'):
        P.close(name, 2e-5)


HEAD_CASES = [(3, 25, 5, 32, 25, 0), (2, 5, 5, 64, 4, 1), (2, 100, 20, 64, 4, 1), (2, 6, 3, 8, 9, 0)]


@pytest.mark.parametrize('case', HEAD_CASES)
@pytest.mark.parametrize('dual', [0, 1])
def test_head(case, dual):
    tasks, n, ways, c, hw, mode = case
    D = c * hw if mode == 0 else c
    torch.manual_seed(6)
    P = Pair()
    P.add('feat', torch.randn(tasks, n, hw, c)).add('featd', torch.randn(tasks, n, hw, c))
    P.add('labels', torch.randint(0, ways, (tasks, 2 * n), dtype=torch.int64))
    Pn = ways * D + ways
    P.add('theta', torch.randn(tasks, Pn) * 0.1).add('v', torch.randn(tasks, Pn)).add('base', torch.randn(tasks, Pn))
    P.out('outp', (tasks, Pn)).out('loss', (tasks,)).out('correct', (tasks,), torch.int32)
    P.out('logits', (tasks, n, ways)).out('gf', (tasks, n, hw, c)).out('gfd', (tasks, n, hw, c))

    def mk(ptr):
        h = XmHeadArgs()
        h.tasks, h.n, h.ways, h.c, h.hw, h.mode, h.dual = tasks, n, ways, c, hw, mode, dual
        h.feat = ptr('feat')
        h.labels, h.label_row0, h.label_row_step, h.labels_per_task = ptr('labels'), 1, 2, 2 * n
        h.w, h.b, h.wb_task_stride = ptr('theta'), ptr('theta', ways * D), Pn
        h.loss, h.correct, h.logits = ptr('loss'), ptr('correct'), ptr('logits')
        if dual:
            h.feat_dot = ptr('featd')
            h.w_dot, h.b_dot, h.wbdot_task_stride = ptr('v'), ptr('v', ways * D), Pn
            h.g_feat_dot = ptr('gfd')
        else:
            h.g_feat = ptr('gf')
        h.out_w, h.out_b, h.out_task_stride = ptr('outp'), ptr('outp', ways * D), Pn
        h.base_w, h.base_b, h.base_task_stride = ptr('base'), ptr('base', ways * D), Pn
        h.scale = -0.5
        return h
    P.run('xm_head', mk)
    for name in ('loss', 'logits', 'outp', 'gfd' if dual else 'gf'):
        P.close(name, 1e-5)
    assert P.gpu['correct'].cpu().tolist() == P.cpu['correct'].tolist()


@pytest.mark.parametrize('case', [(3, 10, 5, 64, 25, 0, 1), (2, 12, 3, 32, 4, 0, 3), (2, 8, 4, 16, 4, 1, 2)])
@pytest.mark.parametrize('first_order', [0, 1])
def test_anil_head(case, first_order):
    tasks, rows, ways, c, hw, mode, steps = case
    D = c * hw if mode == 0 else c
    torch.manual_seed(8)
    P = Pair()
    P.add('feat', torch.randn(tasks, rows, hw, c))
    P.add('labels', torch.randint(0, ways, (tasks, rows), dtype=torch.int64))
    P.add('w', torch.randn(ways, D) * 0.1).add('b', torch.randn(ways) * 0.1)
    Pn = ways * D + ways
    P.out('loss', (tasks,)).out('correct', (tasks,), torch.int32).out('gf', (tasks, rows, hw, c))
    P.out('g', (tasks, Pn))
    lib = _lib.load()
    probe = XmAnilHeadArgs()
    probe.tasks, probe.rows, probe.ways, probe.c, probe.hw, probe.mode, probe.steps = tasks, rows, ways, c, hw, mode, steps
    nbytes = int(lib.xm_anil_head_scratch_bytes(ctypes.byref(probe)))
    P.out('scratch', (nbytes // 4,))

    def mk(ptr):
        h = XmAnilHeadArgs()
        h.tasks, h.rows, h.ways, h.c, h.hw, h.mode = tasks, rows, ways, c, hw, mode
        h.steps, h.first_order, h.lr = steps, first_order, 0.3
        h.feat, h.labels, h.w, h.b = ptr('feat'), ptr('labels'), ptr('w'), ptr('b')
        h.loss, h.correct, h.g_feat = ptr('loss'), ptr('correct'), ptr('gf')
        h.g_w, h.g_b, h.g_task_stride = ptr('g'), ptr('g', ways * D), Pn
        h.scratch, h.scratch_bytes = ptr('scratch'), nbytes
        return h
    P.run('xm_anil_head', mk)
    for name in ('loss', 'gf', 'g'):
        P.close(name, 2e-5)
    assert P.gpu['correct'].cpu().tolist() == P.cpu['correct'].tolist()


def test_outer_step_helpers():
    lib = _lib.load()
    e = emu.EmulatedLib()
    torch.manual_seed(9)
    stream = torch.cuda.current_stream().cuda_stream
    src = torch.randn(5, 1000)
    dst = torch.randn(777)
    d_cpu, d_gpu, s_gpu = dst.clone(), dst.clone().cuda(), src.clone().cuda()
    for acc in (0, 1):
        e.xm_accumulate_tasks(src.data_ptr(), 1000, 5, 777, d_cpu.data_ptr(), acc, None)
        _lib.check(lib.xm_accumulate_tasks(s_gpu.data_ptr(), 1000, 5, 777, d_gpu.data_ptr(), acc, stream), 'acc')
        assert torch.equal(d_gpu.cpu(), d_cpu)          # same order of fp32 additions -> bit-exact
    th, g, m, v = torch.randn(500), torch.randn(500), torch.zeros(500), torch.zeros(500)
    cpu = [t.clone() for t in (th, g, m, v)]
    gpu = [t.clone().cuda() for t in (th, g, m, v)]
    for step in (1, 2, 3):
        e.xm_adam_step(*[t.data_ptr() for t in cpu], 500, 1 / 32, 0.003, 0.9, 0.999, 1e-8, step, None)
        _lib.check(lib.xm_adam_step(*[t.data_ptr() for t in gpu], 500, 1 / 32, 0.003, 0.9, 0.999, 1e-8, step, stream), 'adam')
    for a, b in zip(cpu, gpu):
        assert torch.allclose(b.cpu(), a, rtol=1e-5, atol=1e-7)
    rm, rv, st = torch.zeros(32), torch.ones(32), torch.randn(3, 4, 2, 32)
    c = [rm.clone(), rv.clone()]
    gq = [rm.clone().cuda(), rv.clone().cuda()]
    stg = st.cuda()
    e.xm_bn_ema(c[0].data_ptr(), c[1].data_ptr(), st.data_ptr(), 4, 64, 3, 256, 32, 0.1, None)
    _lib.check(lib.xm_bn_ema(gq[0].data_ptr(), gq[1].data_ptr(), stg.data_ptr(), 4, 64, 3, 256, 32, 0.1, stream), 'ema')
    assert torch.allclose(gq[0].cpu(), c[0], rtol=1e-6, atol=1e-7) and torch.allclose(gq[1].cpu(), c[1], rtol=1e-6, atol=1e-7)


def test_bad_arguments_fail_loudly():
    lib = _lib.load()
    a = XmConvArgs()
    a.g = geom(1, 1, 4, 4, 8, 1, True)
    rc = lib.xm_conv(ctypes.byref(a), None)
    assert rc < 0 and b'null' in lib.xm_last_error()
    with pytest.raises(_lib.XmetaError):
        _lib.check(rc, 'xm_conv')
