"""bench.py's work accounting (algorithmic FLOPs / bytes per C-ABI call) on a launch program built against the
C-ABI emulator: every entry point of the config-2 program is accounted for and the totals match SURVEY App. C."""
import bench
from exploring_meta_b200 import engine as eng
from exploring_meta_b200 import spec as pspec


def test_program_work_covers_the_config2_program(emulated_lib):
    spec = pspec.miniimagenet_spec(5)
    e = eng.MamlEngine(spec, 1, 5, 5, 0.5, mode='second', device='cpu')
    assert e.img
    names = {name for _fn, _a, name in e.prog.calls}
    assert {'xm_img_gram', 'xm_img_fwd', 'xm_img_bwd', 'xm_img_dual_fwd', 'xm_img_dual_bwd', 'xm_conv', 'xm_wgrad',
            'xm_bn_fwd', 'xm_bn_bwd', 'xm_bn_dual_fwd', 'xm_bn_dual_bwd', 'xm_head', 'xm_accumulate_tasks'} == names
    assert len(e.prog.calls) == 201
    work = bench.program_work(e)
    # tensor-core families: 2 FLOP/MAC x positions x 9 cin cout x pairs over the 32 -> 32 layers
    m = [hz * wz * spec.hidden * cin * 9 for (cin, _h, _w, hz, wz, _hp, _wp) in spec.block_dims()]
    rest = sum(m[1:])
    conv_calls = 25 * rest * (5 * (1 + 1) + (1 + 1) + 5 * (2 + 2))      # fwd + dgrad per step / query, x2 pairs in the dual sweeps
    wgrad_calls = 25 * rest * (5 + 1 + 5 * 2)
    assert abs(work['xm_conv']['flops'] - 2 * conv_calls) <= 1e-6 * work['xm_conv']['flops']
    assert abs(work['xm_wgrad']['flops'] - 2 * wgrad_calls) <= 1e-6 * work['xm_wgrad']['flops']
    # every image-block family has bytes and flops; the streaming BN families have bytes
    for fam in ('xm_img_fwd', 'xm_img_bwd', 'xm_img_dual_fwd', 'xm_img_dual_bwd', 'xm_img_gram'):
        assert work[fam]['flops'] > 0 and work[fam]['bytes'] > 0
    for fam in ('xm_bn_fwd', 'xm_bn_bwd', 'xm_bn_dual_fwd', 'xm_bn_dual_bwd'):
        assert work[fam]['bytes'] > 0
    assert bench.family('xm_conv', e.prog.calls[3][1]) == 'xm_conv'
