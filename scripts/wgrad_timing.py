"""Role-level cycle breakdown of wgrad_tc_kernel (build with XM_NVCC_EXTRA=-DXM_TC_TIMING): one 42x42 launch, precision 2."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_meta_b200 import _lib
from exploring_meta_b200._lib import XmBlockGeom, XmWgradArgs
lib = _lib.load()
H = int(sys.argv[1]) if len(sys.argv) > 1 else 42
lib.xm_set_precision(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
g = XmBlockGeom(32, 25, 32, 32, H, H, H, H, H // 2, H // 2, 1, 1)
x = torch.relu(torch.randn(32, 25, H, H, 32, device='cuda')); gz = torch.randn(32, 25, H, H, 32, device='cuda')
nbytes = int(lib.xm_wgrad_scratch_bytes(ctypes.byref(g)))
part = torch.empty(nbytes // 4, device='cuda'); out = torch.zeros(32, 9216 + 32, device='cuda')
a = XmWgradArgs(); a.g = g
a.x1, a.g1 = x.data_ptr(), gz.data_ptr()
a.out_w, a.out_b, a.out_task_stride = out.data_ptr(), out.data_ptr() + 4 * 9216, out.shape[1]
a.scale = 1.0; a.partial, a.partial_bytes = part.data_ptr(), nbytes
for _ in range(2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    _lib.check(lib.xm_wgrad(ctypes.byref(a), torch.cuda.current_stream().cuda_stream), 'wgrad')
    ev1.record(); torch.cuda.synchronize()
    print('--- %.1f us' % (ev0.elapsed_time(ev1) * 1e3))
