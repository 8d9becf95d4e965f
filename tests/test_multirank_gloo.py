"""Meta-batch sharding over ranks: world_size 2 on the gloo backend, each rank running the product's
trainer on the CPU emulator of the C ABI.  Rank r adapts tasks [r*B/2, (r+1)*B/2); ONE sum-allreduce of the flat
[meta-grad ; loss sum ; correct count] buffer; the replicated Adam step must leave every rank with the parameters a
single process computes for the whole meta-batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _install_emulator():
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import cabi_emulator
    from exploring_meta_b200 import _lib, engine
    _lib._lib = cabi_emulator.EmulatedLib()
    engine._require_cuda = lambda device: None


def _one_step(kind, tasks, lo, hi, steps=2):
    from exploring_meta_b200 import spec as pspec
    from exploring_meta_b200.synthetic import make_tasks
    from exploring_meta_b200.trainer import AnilTrainer, MamlTrainer
    torch.manual_seed(0)
    if kind == 'maml':
        spec = pspec.NetSpec(1, 14, 14, 8, 3, 3, False, 'mean')
        X, Y = make_tasks(tasks, 3, 1, (1, 14, 14), seed=21)
        tr = MamlTrainer(spec, hi - lo, 1, 1, 0.3, 0.01, device='cpu', use_graph=False)
        tr.theta.copy_(torch.randn(tr.engine.P) * 0.3)
        theta = tr.theta
    else:
        spec = pspec.NetSpec(3, 12, 12, 8, 3, 2, True, 'none')
        X, Y = make_tasks(tasks, 3, 2, (3, 12, 12), seed=22)
        tr = AnilTrainer(spec, hi - lo, 2, 2, 0.3, 0.01, device='cpu', use_graph=False)
        tr.theta_all.copy_(torch.randn(tr.theta_all.numel()) * 0.2)
        theta = tr.theta_all
    for _ in range(steps):
        tr.meta_step(X[lo:hi], Y[lo:hi])
    loss, acc = tr.metrics()
    if kind == 'maml':       # BatchNorm running statistics: the sharded run composes the ranks' EMA contributions
        stats = torch.cat([torch.cat(tr.running_mean), torch.cat(tr.running_var),
                           torch.tensor([float(tr.num_batches_tracked)])])
    else:
        stats = torch.zeros(1)
    return theta.clone(), float(loss), float(acc), stats


def _worker(rank, world, port, kind, tasks, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    _install_emulator()
    dist.init_process_group('gloo', rank=rank, world_size=world)
    per = tasks // world
    theta, loss, acc, stats = _one_step(kind, tasks, rank * per, (rank + 1) * per)
    out[rank] = (theta, loss, acc, stats)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('kind', ['maml', 'anil'])
def test_two_ranks_equal_one_process(kind):
    tasks = 4
    ctx = mp.get_context('spawn')
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, tasks, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=240)
            assert p.exitcode == 0
        results = dict(out)
    _install_emulator()
    from exploring_meta_b200 import _lib, engine
    saved = (_lib._lib, engine._require_cuda)
    try:
        ref_theta, ref_loss, ref_acc, ref_stats = _one_step(kind, tasks, 0, tasks)
    finally:
        _lib._lib = None
        import importlib
        importlib.reload(engine)
    (t0, l0, a0, s0), (t1, l1, a1, s1) = results[0], results[1]
    # running statistics after two sharded iterations == the single-process sequence of per-call EMA updates
    assert torch.equal(s0, s1)
    assert torch.allclose(s0, ref_stats, rtol=1e-5, atol=1e-6)
    assert torch.equal(t0, t1)                                    # replicated Adam: bit-identical across ranks
    assert (t0 - ref_theta).abs().max() <= 1e-6 * ref_theta.abs().max()   # fp32 reassociation of the task sum only
    assert l0 == pytest.approx(ref_loss, rel=1e-5) and l1 == pytest.approx(ref_loss, rel=1e-5)
    assert a0 == pytest.approx(ref_acc, abs=1e-6)
