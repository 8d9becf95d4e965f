#!/usr/bin/env python3
"""ANIL on vision tasks -- the reference's ``vision/anil_vision.py`` driver (same ``params``, flags, ``AnilVision``
class, metric keys and checkpoint names ``features`` / ``head``) with the per-task loop of ``run()`` (:109-134)
replaced by task-batched launches: per iteration one 'eval' program for the validation tasks, one training program
(body forward once per task over all 2S rows, head-only adaptation inside a single kernel per task, first-order body
backward), the optional allreduce, ``grad / meta_batch_size`` and ONE Adam over body + head parameters
(``AnilTrainer.meta_step``).

Reference quirk kept out: ``fc_neurons`` is derived from the dataset at run time here (the reference freezes it at
import, ``anil_vision.py:40-43``, so its ``--dataset`` flag cannot switch it).  Mini-ImageNet body = ``ConvBase`` default
``hidden=64`` (features 1600), Omniglot body ``hidden=32`` (features 128), as in the reference (:86-89).
"""
import argparse
import os
import random
import time

import numpy as np
import torch
import torch.distributed as dist

from exploring_meta_b200 import _lib
from exploring_meta_b200.core_functions.maml import MAML
from exploring_meta_b200.core_functions.vision import evaluate
from exploring_meta_b200.core_functions.vision_models import ConvBase
from exploring_meta_b200.engine import BN_MOMENTUM, AnilEngine, _p
from exploring_meta_b200.spec import anil_body_spec
from exploring_meta_b200.synthetic import get_tasks
from exploring_meta_b200.trainer import AnilTrainer
from exploring_meta_b200.utils.experiment import Experiment
from exploring_meta_b200.vision.maml_vision import pick_device, sample_stack

params = {
    "ways": 5,
    "shots": 1,
    "outer_lr": 0.003,
    "inner_lr": 0.5,
    "adapt_steps": 1,
    "meta_batch_size": 32,
    "num_iterations": 10000,
    "save_every": 1000,
    "seed": 42,
}

dataset = "min"  # omni or min (omniglot / Mini ImageNet)
cuda = True
wandb = False


class Lambda(torch.nn.Module):
    def __init__(self, fn):
        super(Lambda, self).__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


class AnilVision(Experiment):

    def __init__(self, tasks=None, run=True):
        super(AnilVision, self).__init__(f"anil_{params['ways']}w{params['shots']}s",
                                         dataset, params, path="results/", use_wandb=wandb)
        random.seed(self.params['seed'])
        np.random.seed(self.params['seed'])
        torch.manual_seed(self.params['seed'])
        device = pick_device(self.params['seed'])
        if 'WORLD_SIZE' in os.environ and int(os.environ['WORLD_SIZE']) > 1 and not dist.is_initialized():
            dist.init_process_group('nccl' if device.type == 'cuda' else 'gloo', **({'device_id': device} if device.type == 'cuda' else {}))
        rank = dist.get_rank() if dist.is_initialized() else 0
        if dataset == "omni":
            input_shape = (1, 28, 28)
        elif dataset == "min":
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
        if run:
            self.run(tasks[0], tasks[1], tasks[2], input_shape, device)

    def run(self, train_tasks, valid_tasks, test_tasks, input_shape, device):
        P = self.params
        fc_neurons = 128 if dataset == "omni" else 1600
        if dataset == "omni":
            body = ConvBase(output_size=64, hidden=32, channels=1, max_pool=False)
        else:
            body = ConvBase(output_size=64, channels=3, max_pool=True)
        features = torch.nn.Sequential(body, Lambda(lambda x: x.reshape(-1, fc_neurons)))
        features.to(device)
        head = torch.nn.Linear(fc_neurons, P['ways'])
        head = MAML(head, lr=P['inner_lr'])
        head.to(device)
        self.features, self.head = features, head
        loss = torch.nn.CrossEntropyLoss(reduction='mean')
        self.log_model(features, device, input_shape=input_shape, name='features')
        self.log_model(head, device, input_shape=(P['ways'], fc_neurons), name='head')

        world = dist.get_world_size() if dist.is_initialized() else 1
        if P['meta_batch_size'] % world:
            raise ValueError('meta_batch_size must be divisible by the number of ranks')
        B = P['meta_batch_size'] // world
        spec = anil_body_spec(dataset, P['ways'])
        trainer = AnilTrainer(spec, B, P['shots'], P['adapt_steps'], P['inner_lr'], P['outer_lr'], device=device)
        trainer.load_parameters(features.parameters(), head.parameters())
        valid = AnilEngine(spec, B, P['shots'], P['adapt_steps'], P['inner_lr'], device=device, mode='eval')
        valid.theta, valid.head = trainer.engine.theta, trainer.engine.head
        valid.rebuild()
        bns = [blk.normalize for blk in body.children()]
        lib = _lib.load()

        def write_back():
            with torch.no_grad():
                o = 0
                for p in list(features.parameters()) + list(head.parameters()):
                    p.copy_(trainer.theta_all[o:o + p.numel()].view_as(p))
                    o += p.numel()

        def bn_side_effects():
            # per task: the train body forward, then the validation body forward (anil_vision.py:116-129)
            seq = torch.stack([trainer.engine.call_stats, valid.call_stats]).permute(1, 2, 0, 3, 4).contiguous()
            stream = torch.cuda.current_stream(device).cuda_stream if device.type == 'cuda' else 0
            C = spec.hidden
            world = dist.get_world_size() if dist.is_initialized() else 1
            if world == 1:
                for l, bn in enumerate(bns):                              # seq: [L, B, 2 engines, 2, C]
                    _lib.check(lib.xm_bn_ema(_p(bn.running_mean), _p(bn.running_var), _p(seq[l]), 2 * B, 2 * C, 1, 0, C,
                                             BN_MOMENTUM, stream), 'xm_bn_ema')
                    bn.num_batches_tracked += 2 * B
                return
            # sharded meta-batch: EMA partials of this rank's calls, damped by the calls of the later ranks, summed
            # over ranks; then r <- (1-m)^N r + sum  (same closed form as maml_vision.compose_bn_side_effects)
            part = torch.zeros(len(bns), 2, C, dtype=torch.float32, device=device)
            for l in range(len(bns)):
                _lib.check(lib.xm_bn_ema(_p(part[l, 0]), _p(part[l, 1]), _p(seq[l]), 2 * B, 2 * C, 1, 0, C,
                                         BN_MOMENTUM, stream), 'xm_bn_ema')
            part.mul_((1.0 - BN_MOMENTUM) ** (2 * B * (world - 1 - dist.get_rank())))
            dist.all_reduce(part)
            decay = (1.0 - BN_MOMENTUM) ** (2 * B * world)
            for l, bn in enumerate(bns):
                bn.running_mean.mul_(decay).add_(part[l, 0])
                bn.running_var.mul_(decay).add_(part[l, 1])
                bn.num_batches_tracked += 2 * B * world

        iteration = 0
        t0 = time.time()
        try:
            for iteration in range(P['num_iterations']):
                xv, yv = sample_stack(valid_tasks, B, device)
                valid.run(xv, yv)
                xt, yt = sample_stack(train_tasks, B, device)
                trainer.meta_step(xt, yt)
                bn_side_effects()
                S = P['shots'] * P['ways']
                vals = torch.stack([trainer.engine.loss.mean(), trainer.engine.correct.float().mean() / S,
                                    valid.loss.mean(), valid.correct.float().mean() / S])
                if world > 1:
                    dist.all_reduce(vals)
                    vals /= world
                tl, ta, vl, va = vals.tolist()
                self.log_metrics({'train_loss': tl, 'train_acc': ta, 'valid_loss': vl, 'valid_acc': va})
                if iteration % P['save_every'] == 0:
                    write_back()
                    self.save_model_checkpoint(features, 'features_' + str(iteration + 1))
                    self.save_model_checkpoint(head, 'head_' + str(iteration + 1))
        except KeyboardInterrupt:
            print('\nManually stopped training! Start evaluation & saving...\n')
            self.logger['manually_stopped'] = True
            self.params['num_iterations'] = iteration

        write_back()
        self.save_model(features, name='features')
        self.save_model(head, name='head')
        self.logger['elapsed_time'] = str(round(time.time() - t0, 2)) + ' sec'
        test_acc = evaluate(self.params, test_tasks, head, loss, device, features=features)
        if dist.is_initialized() and dist.get_world_size() > 1:        # every rank evaluated its own test tasks
            acc = torch.tensor([test_acc], dtype=torch.float64, device=device)
            dist.all_reduce(acc)
            test_acc = float(acc.item()) / dist.get_world_size()
        self.logger['test_acc'] = test_acc
        self.log_metrics({'test_acc': self.logger['test_acc']})
        self.save_logs_to_file()


def main(argv=None):
    global dataset
    parser = argparse.ArgumentParser(description='ANIL on Vision')
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
    return AnilVision()


if __name__ == '__main__':
    main()
