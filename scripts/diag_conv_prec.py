"""xm_conv (tcgen05) accuracy and speed per contraction precision: 1 = 3xTF32, 2 = 3xFP16-split (per-tile power-of-two
scaling), against a float64 torch evaluation.  Inputs: (a) post-ReLU-like activations, (b) BN-backward-like cotangents
(zero-mean, mostly pool-winner entries) scaled to 1e-5 (cotangent magnitudes of a mean loss), (c) a tensor whose
images differ in scale by 2^20 (dynamic range across tiles).
  python scripts/diag_conv_prec.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from exploring_meta_b200 import _lib
from exploring_meta_b200._lib import XmBlockGeom, XmConvArgs

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def inputs(kind, tasks, n, hw):
    torch.manual_seed(3)
    if kind == 'relu':
        x = torch.relu(torch.randn(tasks, n, hw, hw, 32, device='cuda') * 0.7 + 0.1)
    elif kind == 'bnbwd':
        g = torch.randn(tasks, n, hw, hw, 32, device='cuda') * (torch.rand(tasks, n, hw, hw, 32, device='cuda') < 0.25)
        x = (g - g.mean(dim=(1, 2, 3), keepdim=True)) * 1e-5
    else:
        x = torch.randn(tasks, n, hw, hw, 32, device='cuda')
        x = x * (2.0 ** (-20.0 * torch.arange(n, device='cuda') / max(n - 1, 1))).view(1, n, 1, 1, 1)
    w = torch.randn(tasks, 32, 32, 3, 3, device='cuda') * 0.06
    return x, w


def run(kind, tasks, n, hw, mode):
    g = XmBlockGeom(tasks, n, 32, 32, hw, hw, hw, hw, hw // 2, hw // 2, 1, 1)
    x, w = inputs(kind, tasks, n, hw)
    out = torch.zeros(tasks, n, hw, hw, 32, device='cuda')
    stats = torch.zeros(tasks, 2, 32, dtype=torch.float64, device='cuda')
    a = XmConvArgs()
    a.g, a.mode, a.stat_mode = g, mode, (1 if mode == 0 else 0)
    a.src1, a.w1, a.w1_task_stride = x.data_ptr(), w.data_ptr(), 9216
    a.out, a.stats = out.data_ptr(), stats.data_ptr()
    xd = x.double().permute(0, 1, 4, 2, 3)
    if mode == 0:
        ref = torch.stack([F.conv2d(xd[t], w[t].double(), padding=1) for t in range(tasks)])
        r32 = torch.stack([F.conv2d(xd[t].float(), w[t], padding=1) for t in range(tasks)])
    else:
        ref = torch.stack([torch.nn.grad.conv2d_input((n, 32, hw, hw), w[t].double(), xd[t], padding=1) for t in range(tasks)])
        r32 = torch.stack([torch.nn.grad.conv2d_input((n, 32, hw, hw), w[t], xd[t].float(), padding=1) for t in range(tasks)])
    ref, r32 = ref.permute(0, 1, 3, 4, 2), r32.permute(0, 1, 3, 4, 2).double()
    # per-image errors: relative to that image's own output norm (shows what per-tile scaling preserves)
    def errs(got):
        d = (got.double() - ref).flatten(2).norm(dim=2) / ref.flatten(2).norm(dim=2).clamp_min(1e-300)
        return float(((got.double() - ref).norm() / ref.norm())), float(d.max())
    line = '%-6s %s tasks %2d %2dx%2d | torch-fp32 %.1e/%.1e' % ((kind, 'fwd  ' if mode == 0 else 'dgrad', tasks, hw, hw) + errs(r32))
    for prec in (1, 2):
        lib.xm_set_precision(prec)
        out.zero_()
        _lib.check(lib.xm_conv(ctypes.byref(a), st), 'xm_conv')
        torch.cuda.synchronize()
        e_all, e_img = errs(out)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            lib.xm_conv(ctypes.byref(a), st)
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 100.0
        fl = 2.0 * tasks * n * hw * hw * 9 * 32 * 32
        line += ' | prec %d: %.1e/%.1e %6.1f us %5.1f TF' % (prec, e_all, e_img, us, fl / us / 1e6)
    lib.xm_set_precision(1)
    print(line, flush=True)


print('errors: rel-L2 over the tensor / worst per-image rel-L2')
for kind in ('relu', 'bnbwd', 'range'):
    for tasks, hw in ((32, 42), (32, 21), (32, 10), (4, 42)):
        for mode in (0, 1):
            run(kind, tasks, 25, hw, mode)
