import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import cabi_emulator as emu
from exploring_meta_b200 import _lib
from exploring_meta_b200._lib import XmBlockGeom, XmWgradArgs

def run(H, n, tasks, use_tc):
    lib = _lib.load()
    lib.xm_set_tcgen05(use_tc)
    g = XmBlockGeom(tasks, n, 32, 32, H, H, H, H, H // 2, H // 2, 1, 1)
    torch.manual_seed(0)
    x = torch.randn(tasks, n, H, H, 32); gz = torch.randn(tasks, n, H, H, 32)
    nb = int(lib.xm_wgrad_scratch_bytes(ctypes.byref(g)))
    outs = {}
    for dev in ('cpu', 'cuda'):
        X, G = x.to(dev), gz.to(dev)
        out = torch.zeros(tasks, 32 * 32 * 9 + 32, device=dev); part = torch.zeros(nb // 4, device=dev)
        a = XmWgradArgs(); a.g = g
        a.x1, a.g1 = X.data_ptr(), G.data_ptr()
        a.out_w, a.out_b, a.out_task_stride = out.data_ptr(), out.data_ptr() + 4 * 9216, 9216 + 32
        a.scale = 1.0; a.partial, a.partial_bytes = part.data_ptr(), nb
        if dev == 'cpu':
            emu.EmulatedLib().xm_wgrad(ctypes.byref(a), None)
        else:
            _lib.check(lib.xm_wgrad(ctypes.byref(a), torch.cuda.current_stream().cuda_stream), 'wgrad'); torch.cuda.synchronize()
        outs[dev] = out.cpu()[:, :9216].view(tasks, 32, 32, 9)      # [co][ci][tap]
    return outs['cpu'], outs['cuda']

E, A = run(5, 1, 1, 1)
print('rel err', ((A - E).norm() / E.norm()).item(), 'norms', A.norm().item(), E.norm().item(), 'nan', torch.isnan(A).any().item())
E0, A0 = E[0], A[0]
def corr(a, b): return (a * b).sum() / (a.norm() * b.norm() + 1e-30)
print('tap x tap correlation (rows: actual tap, cols: expected tap), same (co,ci) orientation')
for ta in range(9):
    print(' '.join('%6.2f' % corr(A0[:, :, ta], E0[:, :, te]) for te in range(9)))
print('transposed orientation')
for ta in range(9):
    print(' '.join('%6.2f' % corr(A0[:, :, ta].t(), E0[:, :, te]) for te in range(9)))
# per-(co) and per-(ci) row correlations for tap 4
print('tap4 per-co corr', [round(float(corr(A0[co, :, 4], E0[co, :, 4])), 2) for co in range(0, 32, 4)])
print('tap4 per-ci corr', [round(float(corr(A0[:, ci, 4], E0[:, ci, 4])), 2) for ci in range(0, 32, 4)])
print('A tap4 block [0:4,0:4]\n', A0[:4, :4, 4], '\nE\n', E0[:4, :4, 4])
E2, A2 = run(5, 1, 1, 0)
print('old kernel rel err', ((A2 - E2).norm() / E2.norm()).item())
