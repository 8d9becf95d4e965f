"""Experiment: one launch program over T tasks vs K concurrent programs over T/K tasks each (separate streams, one
CUDA graph): does overlapping the chains hide the per-kernel fixed latency at small task counts?
  python scripts/exp_split.py --tasks 4 --splits 2"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exploring_meta_b200 import engine as eng, spec as pspec
from exploring_meta_b200.synthetic import make_tasks

ap = argparse.ArgumentParser()
ap.add_argument('--tasks', type=int, default=4)
ap.add_argument('--splits', type=int, default=2)
a = ap.parse_args()
spec = pspec.miniimagenet_spec(5)
theta = pspec.init_flat_params(spec).cuda()
X, Y = make_tasks(a.tasks, 5, 5, (3, 84, 84), seed=0)


def timed(graph, reps=20):
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for k in (1, a.splits):
    per = a.tasks // k
    engines = []
    for i in range(k):
        e = eng.MamlEngine(spec, per, 5, 5, 0.5, device='cuda')
        e.x.copy_(X[i * per:(i + 1) * per]); e.y.copy_(Y[i * per:(i + 1) * per]); e.theta.copy_(theta)
        e.prog.replay(torch.cuda.current_stream().cuda_stream)
        engines.append(e)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(k)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main)
        for e, st in zip(engines, streams):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                e.replay()
            done = torch.cuda.Event(); done.record(st)
            main.wait_event(done)
    print('%d tasks as %d program(s) of %d: %.3f ms per step' % (a.tasks, k, per, timed(g)), flush=True)
    del engines, g
