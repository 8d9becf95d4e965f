"""Role-level cycle breakdown of conv_tc_kernel (build with XM_NVCC_EXTRA=-DXM_TC_TIMING): one 42x42 forward launch."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_meta_b200 import _lib
from exploring_meta_b200._lib import XmBlockGeom, XmConvArgs
lib = _lib.load()
H = int(sys.argv[1]) if len(sys.argv) > 1 else 42
lib.xm_set_precision(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
g = XmBlockGeom(32, 25, 32, 32, H, H, H, H, H // 2, H // 2, 1, 1)
x = torch.randn(32, 25, H, H, 32, device='cuda'); w = torch.randn(32, 32 * 32 * 9, device='cuda') * 0.05
out = torch.empty_like(x); stats = torch.zeros(32, 2, 32, dtype=torch.float64, device='cuda')
a = XmConvArgs(); a.g = g; a.mode = 0; a.stat_mode = 1
a.src1, a.w1, a.w1_task_stride, a.out, a.stats = x.data_ptr(), w.data_ptr(), 9216, out.data_ptr(), stats.data_ptr()
for _ in range(2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    _lib.check(lib.xm_conv(ctypes.byref(a), torch.cuda.current_stream().cuda_stream), 'conv')
    ev1.record()
    torch.cuda.synchronize()
    Q = 25 * (H + 1) * (H + 1)
    tiles = (Q + 125) // 126
    print('--- kernel %.1f us; %d tiles per task, %.1f per SM -> %.3f us per tile' % (
        ev0.elapsed_time(ev1) * 1e3, tiles, 32 * tiles / 148.0, ev0.elapsed_time(ev1) * 1e3 / (32 * tiles / 148.0)))
if os.environ.get('XM_TIMING_REPS'):
    # steady-state figure: mean over back-to-back launches on two alternating input / output sets (> L2 together)
    reps = int(os.environ['XM_TIMING_REPS'])
    x2, out2 = torch.randn_like(x), torch.empty_like(x)
    b = XmConvArgs(); b.g = g; b.mode = 0; b.stat_mode = 1
    b.src1, b.w1, b.w1_task_stride, b.out, b.stats = x2.data_ptr(), w.data_ptr(), 9216, out2.data_ptr(), stats.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    for i in range(4):
        lib.xm_conv(ctypes.byref(a if i & 1 else b), st)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(reps):
        lib.xm_conv(ctypes.byref(a if i & 1 else b), st)
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / reps
    print('=== steady state: %.1f us per launch (memset + kernel), %.3f us per tile and SM' % (us, us / (32 * tiles / 148.0)))
