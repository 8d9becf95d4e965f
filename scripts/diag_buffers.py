"""Where do a 32-task and a 4-task launch program first differ for the same tasks?  (first-order mode, query pass)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

spec = pspec.miniimagenet_spec(5)
theta = pspec.init_flat_params(spec, seed=42).cuda()
X, Y = make_tasks(32, 5, 5, (3, 84, 84), seed=0)
X, Y = X.cuda(), Y.cuda()
T, lr = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 0.001
big = eng.MamlEngine(spec, 32, 5, T, lr, mode='first', device='cuda')
big.run(X, Y, theta); torch.cuda.synchronize()
small = eng.MamlEngine(spec, 4, 5, T, lr, mode='first', device='cuda')
small.run(X[:4], Y[:4], theta); torch.cuda.synchronize()


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


for l in range(4):
    if l > 0:
        print('L%d  tZ %.2e' % (l, rel(small.tZ[l], big.tZ[l][:4])), end='  ')
    else:
        print('L0  zsel %.2e' % rel(small.tZ[0][1]['zsel'], big.tZ[0][1]['zsel'][:4]), end='  ')
    print('tP %.2e  tGP %.2e  tMI %.2e  tBR %.2e' % (rel(small.tP[l], big.tP[l][:4]), rel(small.tGP[l], big.tGP[l][:4]),
                                                     rel(small.tMI[l], big.tMI[l][:4]), rel(small.tBR[l], big.tBR[l][:4])))
offs, P = spec.param_offsets()
names = []
for l in range(4):
    names += ['bn%d.g' % l, 'bn%d.b' % l, 'conv%d.w' % l, 'conv%d.b' % l]
names += ['lin.w', 'lin.b']
bs, bb = small.bar[0], big.bar[0][:4]
for i, n in enumerate(names):
    a, b = offs[i], (offs[i + 1] if i + 1 < len(offs) else P)
    print('%-8s per-task grads rel %.2e' % (n, rel(bs[:, a:b], bb[:, a:b])))
print('theta_1 rel', rel(small.theta_steps[0], big.theta_steps[0][:4]))

# the summed gradient of tasks 0..3: the engine's task-ordered fp32 sum vs a double sum of the per-task rows
for e, name in ((small, '4-task engine'),):
    print(name, 'grad vs double sum of its per-task rows: rel', rel(e.grad, e.bar[0].double().sum(0)))
print('32-task engine grad vs double sum of its rows: rel', rel(big.grad, big.bar[0].double().sum(0)))
print('per-task row norms (first 4):', [round(float(v), 3) for v in big.bar[0][:4].norm(dim=1)], ' sum norm', float(big.grad.norm()))
