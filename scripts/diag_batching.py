"""Per-parameter-tensor difference between the 32-task meta-gradient and the sum of eight 4-task ones (config 2
shapes, calm inner lr): where does batching-dependent fp32 noise enter?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

lr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.001
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else 'second'
spec = pspec.miniimagenet_spec(5)
theta = pspec.init_flat_params(spec, seed=42).cuda()
X, Y = make_tasks(32, 5, 5, (3, 84, 84), seed=0)
X, Y = X.cuda(), Y.cuda()
big = eng.MamlEngine(spec, 32, 5, T, lr, mode=mode, device='cuda')
big.run(X, Y, theta); torch.cuda.synchronize()
g32 = big.grad.clone().double(); rows32 = big.bar[0].clone().double() if mode == 'first' else None; del big
small = eng.MamlEngine(spec, 4, 5, T, lr, mode=mode, device='cuda')
gs = torch.zeros_like(g32)
for t0 in range(0, 32, 4):
    small.run(X[t0:t0 + 4], Y[t0:t0 + 4], theta); torch.cuda.synchronize()
    gs += small.grad.double()
    if rows32 is not None:
        r = small.bar[0].double()
        for t in range(4):
            d = ((r[t] - rows32[t0 + t]).norm() / rows32[t0 + t].norm()).item()
            if d > 1e-5:
                print('task', t0 + t, 'per-task gradient differs: rel', d)
offs, P = spec.param_offsets()
names = []
for l in range(spec.layers):
    names += ['bn%d.g' % l, 'bn%d.b' % l, 'conv%d.w' % l, 'conv%d.b' % l]
names += ['lin.w', 'lin.b']
print('img path', small.img, 'lr', lr, 'T', T, mode, 'total rel', ((gs - g32).norm() / g32.norm()).item())
for i, n in enumerate(names):
    a, b = offs[i], (offs[i + 1] if i + 1 < len(offs) else P)
    d = (gs[a:b] - g32[a:b]).norm().item(); nn = g32[a:b].norm().item()
    print('%-8s |g| %.4e  |diff| %.3e  rel %.2e' % (n, nn, d, d / max(nn, 1e-30)))
