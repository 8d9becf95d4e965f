"""xm_wgrad / xm_conv accuracy against a float64 torch evaluation on the GPU, per geometry and task count."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from exploring_meta_b200 import _lib
from exploring_meta_b200._lib import XmBlockGeom, XmWgradArgs, XmConvArgs

lib = _lib.load()
if len(sys.argv) > 1:
    lib.xm_set_precision(int(sys.argv[1]))
    print('precision', sys.argv[1])
st = torch.cuda.current_stream().cuda_stream


def run(tasks, n, hw, realistic):
    g = XmBlockGeom(tasks, n, 32, 32, hw, hw, hw, hw, hw // 2, hw // 2, 1, 1)
    torch.manual_seed(0)
    x = torch.relu(torch.randn(tasks, n, hw, hw, 32, device='cuda')) * (1.0 if realistic else 1.0)
    gz = torch.randn(tasks, n, hw, hw, 32, device='cuda')
    if realistic:   # BN-backward-like cotangent: zero mean per channel, mostly zeros (pool winners only)
        mask = (torch.rand_like(gz) < 0.25).float()
        gz = gz * mask
        gz = gz - gz.mean(dim=(1, 2, 3), keepdim=True)
    nbytes = int(lib.xm_wgrad_scratch_bytes(ctypes.byref(g)))
    part = torch.empty(nbytes // 4, device='cuda')
    out = torch.zeros(tasks, 32 * 32 * 9 + 32, device='cuda')
    a = XmWgradArgs()
    a.g = g
    a.x1, a.g1 = x.data_ptr(), gz.data_ptr()
    a.out_w, a.out_b, a.out_task_stride = out.data_ptr(), out.data_ptr() + 4 * 9216, out.shape[1]
    a.scale = 1.0
    a.partial, a.partial_bytes = part.data_ptr(), nbytes
    _lib.check(lib.xm_wgrad(ctypes.byref(a), st), 'xm_wgrad')
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        lib.xm_wgrad(ctypes.byref(a), st)
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 100.0
    xd = x.double().permute(0, 1, 4, 2, 3)
    gd = gz.double().permute(0, 1, 4, 2, 3)
    ref = torch.stack([torch.nn.grad.conv2d_weight(xd[t], (32, 32, 3, 3), gd[t], padding=1) for t in range(tasks)])
    ref32 = torch.stack([torch.nn.grad.conv2d_weight(xd[t].float(), (32, 32, 3, 3), gd[t].float(), padding=1) for t in range(tasks)])
    got = out[:, :9216].view(tasks, 32, 32, 3, 3).double()
    e = ((got - ref).norm() / ref.norm()).item()
    e32 = ((ref32.double() - ref).norm() / ref.norm()).item()
    es = ((got.sum(0) - ref.sum(0)).norm() / ref.sum(0).norm()).item()
    print('wgrad tasks %3d n %2d %2dx%2d %s: rel-L2 ours %.2e  torch-fp32 %.2e   task-sum ours %.2e   %7.1f us (incl. reduce) %5.1f TF' % (
        tasks, n, hw, hw, 'bn-like' if realistic else 'randn ', e, e32, es, us, 2.0 * tasks * n * hw * hw * 9 * 32 * 32 / us / 1e6))


for realistic in (False, True):
    for tasks in (4, 32):
        for hw in (10, 42):
            run(tasks, 25, hw, realistic)


def run_conv(tasks, n, hw, mode, reps=3):
    g = XmBlockGeom(tasks, n, 32, 32, hw, hw, hw, hw, hw // 2, hw // 2, 1, 1)
    torch.manual_seed(1)
    x = torch.randn(tasks, n, hw, hw, 32, device='cuda')
    w = torch.randn(tasks, 32, 32, 3, 3, device='cuda') * 0.1
    out = torch.zeros(tasks, n, hw, hw, 32, device='cuda')
    stats = torch.zeros(tasks, 2, 32, dtype=torch.float64, device='cuda')
    a = XmConvArgs()
    a.g, a.mode, a.stat_mode = g, mode, (1 if mode == 0 else 0)
    a.src1, a.w1, a.w1_task_stride = x.data_ptr(), w.data_ptr(), 9216
    a.out, a.stats = out.data_ptr(), stats.data_ptr()
    xd = x.double().permute(0, 1, 4, 2, 3)
    if mode == 0:
        ref = torch.stack([F.conv2d(xd[t], w[t].double(), padding=1) for t in range(tasks)])
    else:
        ref = torch.stack([torch.nn.grad.conv2d_input((n, 32, hw, hw), w[t].double(), xd[t], padding=1) for t in range(tasks)])
    ref = ref.permute(0, 1, 3, 4, 2)
    worst = 0.0
    for _ in range(reps):
        out.zero_()
        _lib.check(lib.xm_conv(ctypes.byref(a), st), 'xm_conv')
        torch.cuda.synchronize()
        per_task = (out.double() - ref).flatten(1).norm(dim=1) / ref.flatten(1).norm(dim=1)
        worst = max(worst, per_task.max().item())
        bad = (per_task > 1e-5).nonzero().flatten().tolist()
        if bad:
            print('   bad tasks', bad, [float(per_task[b]) for b in bad])
        if mode == 0:
            s_ref = ref.sum(dim=(1, 2, 3))
            es = ((stats[:, 0] - s_ref).abs().max() / s_ref.abs().max()).item()
            if es > 1e-6:
                print('   stats off', es)
    print('conv %s tasks %3d n %2d %2dx%2d: worst per-task rel-L2 %.2e' % ('fwd  ' if mode == 0 else 'dgrad', tasks, n, hw, hw, worst))


for tasks in (4, 32):
    for hw in (10, 21, 42):
        for mode in (0, 1):
            run_conv(tasks, 25, hw, mode)
