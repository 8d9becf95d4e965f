"""Central differences of the mean adapted query loss vs <meta-gradient, v> at config-2 size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

lr, steps = float(sys.argv[1]), int(sys.argv[2])
spec = pspec.miniimagenet_spec(5)
theta = pspec.init_flat_params(spec, seed=42).cuda()
X, Y = make_tasks(32, 5, 5, (3, 84, 84), seed=0)
X, Y = X.cuda(), Y.cuda()
e = eng.MamlEngine(spec, 32, 5, steps, lr, device='cuda'); e.run(X, Y, theta); torch.cuda.synchronize()
g = e.grad.clone().double() / 32; del e
fo = eng.MamlEngine(spec, 32, 5, steps, lr, mode='first', device='cuda'); fo.run(X, Y, theta); torch.cuda.synchronize()
g1 = fo.grad.clone().double() / 32; del fo
ev = eng.MamlEngine(spec, 32, 5, steps, lr, mode='eval', device='cuda')


def L(th):
    ev.run(X, Y, th.float()); torch.cuda.synchronize()
    return ev.loss.double().mean().item()


print('|g| %.4f |g1| %.4f rel(g1,g) %.3e' % (g.norm(), g1.norm(), ((g1 - g).norm() / g.norm()).item()))
torch.manual_seed(0)
offs, P = spec.param_offsets()
for trial in range(4):
    if trial == 0:
        v = g / g.norm()
    elif trial == 3:                      # only the first block's parameters
        v = torch.zeros_like(g); v[:offs[4]] = torch.randn(offs[4], device='cuda', dtype=torch.float64); v /= v.norm()
    else:
        v = torch.randn_like(g); v /= v.norm()
    an, an1 = torch.dot(g, v).item(), torch.dot(g1, v).item()
    fds = []
    for h in (5e-4, 1e-3, 2e-3, 4e-3, 8e-3):
        fds.append((L(theta.double() + h * v) - L(theta.double() - h * v)) / (2 * h))
    print('trial %d  <g,v> %.5f  <g1,v> %.5f  fd(h=5e-4..8e-3) %s' % (trial, an, an1, ' '.join('%.5f' % f for f in fds)))
