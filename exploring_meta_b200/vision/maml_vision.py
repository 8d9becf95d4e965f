#!/usr/bin/env python3
"""MAML on vision tasks -- the reference's ``vision/maml_vision.py`` driver (same ``params``, same flags, same
``MamlVision(Experiment)`` class, same metric keys and checkpoint files) with the per-task Python loop of
``run()`` (:102-124) replaced by task-batched launches:

  * the ``meta_batch_size`` validation tasks of an iteration run as one 'eval' launch program on the pre-step
    parameters (the reference interleaves them with the train tasks; they never influence the update);
  * the ``meta_batch_size`` train tasks run as one second-order launch program, followed by the (optional) NCCL
    allreduce, ``grad / meta_batch_size`` and Adam -- ``MamlTrainer.meta_step``;
  * BatchNorm running statistics receive the same sequence of per-call updates the reference's shared buffers
    see (task-major, train call block then validation call block per task);
  * losses / accuracies are read back once per iteration instead of four ``.item()`` syncs per task.

Data: learn2learn's Omniglot / Mini-ImageNet task datasets need a download and are outside the hot path;
``exploring_meta_b200.synthetic.get_tasks`` supplies seeded synthetic task streams with the ``sample()``
interface.  Pass real ``(train, valid, test)`` task objects to ``MamlVision(tasks=...)`` to train on data.
Under ``torchrun`` the meta-batch is sharded over the ranks (``meta_batch_size`` must divide evenly).
"""
import argparse
import os
import random

import numpy as np
import torch
import torch.distributed as dist

from exploring_meta_b200 import _lib
from exploring_meta_b200.core_functions.maml import MAML
from exploring_meta_b200.core_functions.vision import evaluate
from exploring_meta_b200.core_functions.vision_models import MiniImagenetCNN, OmniglotCNN, net_spec_of
from exploring_meta_b200.engine import BN_MOMENTUM, MamlEngine, _p
from exploring_meta_b200.synthetic import get_tasks
from exploring_meta_b200.trainer import MamlTrainer
from exploring_meta_b200.utils.experiment import Experiment

params = {
    "ways": 5,
    "shots": 1,
    "outer_lr": 0.003,
    "inner_lr": 0.5,
    "adapt_steps": 1,
    "meta_batch_size": 32,
    "num_iterations": 10000,  # 10k for Mini-ImageNet, 5k for Omniglot
    "save_every": 1000,
    "seed": 42,
}

dataset = "min"  # omni or min (omniglot / Mini ImageNet)
omni_cnn = True
cuda = True
wandb = False


def pick_device(seed):
    """cuda:LOCAL_RANK.  There is no CPU path: a CPU device is returned only when a test has swapped the C ABI for
    its CPU emulator (tests/cabi_emulator.py), which is the one case ``engine._require_cuda`` lets through."""
    from exploring_meta_b200 import engine
    if cuda and torch.cuda.device_count():
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        print(f'Running with CUDA and device {torch.cuda.get_device_name(local)}')
        torch.cuda.manual_seed(seed)
        return torch.device('cuda', local)
    engine._require_cuda(torch.device('cpu'))            # raises XmetaError outside the emulator
    return torch.device('cpu')


def sample_stack(tasks, n, device):
    """``n`` x ``tasks.sample()`` stacked into ``x [n, rows, C, H, W]``, ``y [n, rows]`` on ``device``."""
    if hasattr(tasks, 'sample_batch'):
        x, y = tasks.sample_batch(n)
    else:
        got = [tasks.sample() for _ in range(n)]
        x, y = torch.stack([g[0] for g in got]), torch.stack([g[1] for g in got])
    return x.to(device, non_blocking=True), y.to(device, non_blocking=True)


def compose_bn_side_effects(model, engines, device):
    """Applies the per-call BN statistics of the given engines to the model's running buffers in the reference's
    order: for every task, the calls of engines[0] (train), then engines[1] (validation), ...
    (vision/maml_vision.py:104-122: train fast_adapt, then validation fast_adapt, per task)."""
    lib = _lib.load()
    bns = [blk.normalize for blk in model.base.children()]
    stacked = torch.stack([e.call_stats for e in engines])            # [E, T+1, L, B, 2, C]
    E, T1, L, B, _two, C = stacked.shape
    seq = stacked.permute(2, 3, 0, 1, 4, 5).contiguous()              # [L, B, E, T+1, 2, C]
    stream = torch.cuda.current_stream(device).cuda_stream if device.type == 'cuda' else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    calls = B * E * T1
    if world == 1:
        for l, bn in enumerate(bns):
            _lib.check(lib.xm_bn_ema(_p(bn.running_mean), _p(bn.running_var), _p(seq[l]), calls, 2 * C, 1, 0, C,
                                     BN_MOMENTUM, stream), 'xm_bn_ema')
            bn.num_batches_tracked += calls
        return
    # Sharded meta-batch: the EMA over forward calls is linear, so this rank contributes the EMA of its own calls
    # started from zero, damped by (1-m)^(calls of the ranks after it); one small allreduce, then every rank applies
    # r <- (1-m)^N r + sum (same closed form as trainer.MamlTrainer).
    part = torch.zeros(len(bns), 2, C, dtype=torch.float32, device=device)
    for l in range(len(bns)):
        _lib.check(lib.xm_bn_ema(_p(part[l, 0]), _p(part[l, 1]), _p(seq[l]), calls, 2 * C, 1, 0, C,
                                 BN_MOMENTUM, stream), 'xm_bn_ema')
    part.mul_((1.0 - BN_MOMENTUM) ** (calls * (world - 1 - dist.get_rank())))
    dist.all_reduce(part)
    decay = (1.0 - BN_MOMENTUM) ** (calls * world)
    for l, bn in enumerate(bns):
        bn.running_mean.mul_(decay).add_(part[l, 0])
        bn.running_var.mul_(decay).add_(part[l, 1])
        bn.num_batches_tracked += calls * world


class MamlVision(Experiment):

    def __init__(self, tasks=None, run=True):
        super(MamlVision, self).__init__(f"maml_{params['ways']}w{params['shots']}s",
                                         dataset, params, path="results/", use_wandb=wandb)
        random.seed(self.params['seed'])
        np.random.seed(self.params['seed'])
        torch.manual_seed(self.params['seed'])
        device = pick_device(self.params['seed'])
        if 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) > 1 and not dist.is_initialized():
            dist.init_process_group('nccl' if device.type == 'cuda' else 'gloo', **({'device_id': device} if device.type == 'cuda' else {}))
        rank = dist.get_rank() if dist.is_initialized() else 0

        if dataset == "omni":
            model = OmniglotCNN(self.params['ways'])
            self.params['model_type'] = 'omni_CNN'
            input_shape = (1, 28, 28)
        elif dataset == "min":
            model = MiniImagenetCNN(self.params['ways'])
            input_shape = (3, 84, 84)
        else:
            print("Dataset not supported")
            raise SystemExit(2)
        if tasks is None:
            # The reference's learn2learn dataset builders (utils/data_pre.py:16-112) need a download and are not
            # re-implemented: without caller-supplied task objects the run is on SYNTHETIC tasks, and says so.
            print('WARNING: no task datasets given -- training on synthetic %s-shaped tasks '
                  '(exploring_meta_b200.synthetic.get_tasks); pass tasks=(train, valid, test) objects with .sample() '
                  'for real data' % dataset, flush=True)
            self.logger['data'] = self.params['data'] = 'synthetic'
            tasks = get_tasks(dataset, self.params['ways'], self.params['shots'], seed=self.params['seed'] + 7919 * rank)
        else:
            self.logger['data'] = self.params['data'] = 'caller-supplied task datasets'
        self.model = model
        if run:
            self.run(tasks[0], tasks[1], tasks[2], model, input_shape, device)

    def run(self, train_tasks, valid_tasks, test_tasks, model, input_shape, device):
        P = self.params
        model.to(device)
        maml = MAML(model, lr=P['inner_lr'], first_order=False)
        loss = torch.nn.CrossEntropyLoss(reduction='mean')
        self.log_model(maml, device, input_shape=input_shape)

        world = dist.get_world_size() if dist.is_initialized() else 1
        if P['meta_batch_size'] % world:
            raise ValueError('meta_batch_size must be divisible by the number of ranks')
        B = P['meta_batch_size'] // world
        spec = net_spec_of(model)
        trainer = MamlTrainer(spec, B, P['shots'], P['adapt_steps'], P['inner_lr'], P['outer_lr'],
                              first_order=False, device=device)
        trainer.load_parameters(model.parameters())
        valid = MamlEngine(spec, B, P['shots'], P['adapt_steps'], P['inner_lr'], mode='eval', device=device)
        valid.theta = trainer.theta                       # validation reads the live master parameters
        valid.rebuild()

        def write_back():
            with torch.no_grad():
                o = 0
                for p in model.parameters():
                    p.copy_(trainer.theta[o:o + p.numel()].view_as(p))
                    o += p.numel()

        iteration = 0
        import time
        t0 = time.time()
        try:
            for iteration in range(P['num_iterations']):
                xv, yv = sample_stack(valid_tasks, B, device)
                valid.run(xv, yv)                                          # on theta_i, before the outer step
                xt, yt = sample_stack(train_tasks, B, device)
                trainer.meta_step(xt, yt, track_running_stats=False)       # second-order meta-gradient + Adam
                compose_bn_side_effects(model, [trainer.engine, valid], device)
                S = P['shots'] * P['ways']
                vals = torch.stack([trainer.engine.loss.mean(), trainer.engine.correct.float().mean() / S,
                                    valid.loss.mean(), valid.correct.float().mean() / S])
                if world > 1:
                    dist.all_reduce(vals)
                    vals /= world
                tl, ta, vl, va = vals.tolist()                             # the iteration's only host sync
                metrics = {'train_loss': tl, 'train_acc': ta, 'valid_loss': vl, 'valid_acc': va}
                self.log_metrics(metrics)
                if iteration % P['save_every'] == 0:
                    write_back()
                    self.save_model_checkpoint(model, str(iteration))
        except KeyboardInterrupt:
            print('\nManually stopped training! Start evaluation & saving...\n')
            self.logger['manually_stopped'] = True
            self.params['num_iterations'] = iteration

        write_back()
        self.save_model(model)
        self.logger['elapsed_time'] = str(round(time.time() - t0, 2)) + ' sec'
        test_acc = evaluate(self.params, test_tasks, maml, loss, device)
        if dist.is_initialized() and dist.get_world_size() > 1:        # every rank evaluated its own test tasks
            acc = torch.tensor([test_acc], dtype=torch.float64, device=device)
            dist.all_reduce(acc)
            test_acc = float(acc.item()) / dist.get_world_size()
        self.logger['test_acc'] = test_acc
        self.log_metrics({'test_acc': self.logger['test_acc']})
        self.save_logs_to_file()


def main(argv=None):
    global dataset
    parser = argparse.ArgumentParser(description='MAML on Vision')
    parser.add_argument('--dataset', type=str, default=dataset, help='Pick a dataset')
    parser.add_argument('--ways', type=int, default=params['ways'], help='N-ways (classes)')
    parser.add_argument('--shots', type=int, default=params['shots'], help='K-shots (samples per class)')
    parser.add_argument('--outer_lr', type=float, default=params['outer_lr'], help='Outer lr')
    parser.add_argument('--inner_lr', type=float, default=params['inner_lr'], help='Inner lr')
    parser.add_argument('--adapt_steps', type=int, default=params['adapt_steps'], help='Adaptation steps in inner loop')
    parser.add_argument('--meta_batch_size', type=int, default=params['meta_batch_size'], help='Batch size')
    parser.add_argument('--num_iterations', type=int, default=params['num_iterations'], help='Number of epochs')
    parser.add_argument('--save_every', type=int, default=params['save_every'], help='Interval to save model')
    parser.add_argument('--seed', type=int, default=params['seed'], help='Seed')
    args = parser.parse_args(argv)
    dataset = args.dataset
    for key in ('ways', 'shots', 'outer_lr', 'inner_lr', 'adapt_steps', 'meta_batch_size', 'num_iterations',
                'save_every', 'seed'):
        params[key] = getattr(args, key)
    return MamlVision()


if __name__ == '__main__':
    main()
