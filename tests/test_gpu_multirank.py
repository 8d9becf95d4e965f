"""Hardware twin of test_multirank_gloo.py: 2 ranks on 2 GPUs of one box (one process per GPU), the meta-batch sharded
over the ranks, shard sums combined inside the Adam kernel over NVLink peer memory (csrc/comm.cu; also through NCCL for
comparison).  After two meta-iterations every rank must hold bit-identical parameters, equal to what ONE GPU computes
for the whole meta-batch up to fp32 reassociation of the task sum (~1e-6 rel).
Needs >= 2 GPUs: skipped on a single-GPU box (run with ``gpurun --gpus 2``)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _case(kind):
    from exploring_meta_b200 import spec as pspec
    if kind == 'omni':          # BASELINE config 1: 32 tasks
        return pspec.omniglot_spec(5), 32, 1, 1, 0.5, (1, 28, 28), 31
    return pspec.miniimagenet_spec(5), 8, 1, 2, 0.01, (3, 84, 84), 32     # config-2 network, 8 tasks, 2 inner steps


def _train(kind, lo, hi, device, iterations=2):
    from exploring_meta_b200 import spec as pspec
    from exploring_meta_b200.synthetic import make_tasks
    from exploring_meta_b200.trainer import MamlTrainer
    spec, tasks, shots, steps, lr, shape, seed = _case(kind)
    tr = MamlTrainer(spec, hi - lo, shots, steps, lr, 0.003, device=device, use_graph=True)
    tr.theta.copy_(pspec.init_flat_params(spec, seed=42))
    for it in range(iterations):
        X, Y = make_tasks(tasks, spec.ways, shots, shape, seed=seed + it)
        tr.meta_step(X[lo:hi].to(device), Y[lo:hi].to(device))
    torch.cuda.synchronize(device)
    if tr.comm is not None:
        tr.comm.check()
    loss, acc = tr.metrics()
    stats = torch.cat([torch.cat(tr.running_mean), torch.cat(tr.running_var)]).cpu()
    return tr.theta.cpu().clone(), float(loss), float(acc), stats


def _worker(rank, world, port, kind, transport, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      XM_COMM=transport)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    tasks = _case(kind)[1]
    per = tasks // world
    out[rank] = _train(kind, rank * per, (rank + 1) * per, dev)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
@pytest.mark.parametrize('transport', ['p2p', 'nccl'])
@pytest.mark.parametrize('kind', ['omni', 'min'])
def test_two_gpus_equal_one_gpu(kind, transport):
    ctx = mp.get_context('spawn')
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, transport, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=300)
            assert p.exitcode == 0
        results = dict(out)
    ref_theta, ref_loss, ref_acc, ref_stats = _train(kind, 0, _case(kind)[1], torch.device('cuda', 0))
    (t0, l0, a0, s0), (t1, l1, a1, s1) = results[0], results[1]
    assert torch.equal(t0, t1), 'replicas diverged'                      # replicated Adam on identical sums
    assert torch.equal(s0, s1)
    d = float((t0 - ref_theta).abs().max() / ref_theta.abs().max())
    print('%s/%s: max |theta_2gpu - theta_1gpu| / max|theta| = %.2e' % (kind, transport, d))
    assert d <= 1e-6
    assert torch.allclose(s0, ref_stats, rtol=1e-5, atol=1e-6)
    assert l0 == pytest.approx(ref_loss, rel=1e-5) and l1 == pytest.approx(ref_loss, rel=1e-5)
    assert a0 == pytest.approx(ref_acc, abs=1e-6)
