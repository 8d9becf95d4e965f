#!/usr/bin/env python
"""bench.py -- meta-train tasks/sec of the MAML / ANIL hot path.  Default = BASELINE.json configs[1]:
MAML Mini-ImageNet 5-way 5-shot, 4-conv 32-filter CNN, 5 inner steps, second-order, GLOBAL meta-batch 32.

  python bench.py [--gpus N --steps K --warmup W]          our arm (N>1: launched under torchrun)
  python bench.py --impl reference ...                     the reference's CPU path (oracle port) on host cores
  python bench.py --config 4 ...                           MAML Omniglot 20-way 5-shot, global meta-batch 256
                                                           (configs 1 and 3 likewise: parity cases, not the headline)

Scaling is STRONG by default: the named global meta-batch is sharded over the N ranks (32 -> 32/N tasks per GPU), as
north_star states the target; ``--scaling weak`` keeps the named batch PER GPU.  With N > 1 the strong run also times
the weak variant and reports it under "weak_scaling".

One "step" = one meta-iteration over one synthetic meta-batch: adapt T steps on the support rows of every task, query
loss, second-order meta-gradient, sum over ranks (NVLink peer memory, inside the Adam kernel), grad/B, Adam, BN
running-statistics update.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUTER_LR = 0.003
# BASELINE.json configs (1-based).  flops: SURVEY App. C (spec.flops_per_train_task)
CONFIGS = {
    1: dict(algo='maml', kind='omni', ways=5, shots=1, steps=1, inner_lr=0.5, batch=32,
            metric='meta-train tasks/sec (MAML Omniglot 5w1s)',
            workload='MAML Omniglot 5-way 1-shot, 4-conv 64-filter CNN, 1 inner step, second-order, meta-batch 32 '
                     '(synthetic 28x28x1)'),
    2: dict(algo='maml', kind='min', ways=5, shots=5, steps=5, inner_lr=0.5, batch=32,
            metric='meta-train tasks/sec (MAML MiniImageNet 5w5s)',
            workload='MAML Mini-ImageNet 5-way 5-shot, 4-conv 32-filter CNN, 5 inner steps, second-order, '
                     'meta-batch 32 (synthetic 84x84x3)'),
    3: dict(algo='anil', kind='min', ways=5, shots=5, steps=1, inner_lr=0.5, batch=32,
            metric='meta-train tasks/sec (ANIL MiniImageNet 5w5s)',
            workload='ANIL Mini-ImageNet 5-way 5-shot, 64-filter body forward once + head-only adaptation, '
                     'meta-batch 32 (synthetic 84x84x3)'),
    4: dict(algo='maml', kind='omni', ways=20, shots=5, steps=1, inner_lr=0.5, batch=256,
            metric='meta-train tasks/sec (MAML Omniglot 20w5s)',
            workload='MAML Omniglot 20-way 5-shot, 4-conv 64-filter CNN, 1 inner step, second-order, meta-batch 256 '
                     '(synthetic 28x28x1)'),
}
IN_SHAPE = {'omni': (1, 28, 28), 'min': (3, 84, 84)}
# BASELINE.json configs[4]: MAML-TRPO policy MLP on synthetic Particles2D-style replays
RL = dict(tasks=40, episodes=20, horizon=100, inner_lr=0.001, gamma=0.99, tau=1.0, value_reg=2, max_kl=0.01,
          ls_max_steps=15, backtrack_factor=0.5, outer_lr=1.0,
          metric='meta-train tasks/sec (MAML-TRPO policy MLP)',
          workload='MAML-TRPO 2x100 tanh policy MLP, 40 tasks x (support + query replay of 20 episodes x 100 steps), '
                   'one first-order adaptation per task + one meta_optimize_trpo (second-order gradient, 10 CG '
                   'iterations of Fisher-vector products, line search) per step; synthetic Particles2D-style replays')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=2, choices=sorted(CONFIGS) + [5])
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
    ap.add_argument('--tasks', type=int, default=0, help='tasks per GPU (default: from --config and --scaling)')
    ap.add_argument('--inner-steps', type=int, default=0, help='override the config\'s inner steps')
    ap.add_argument('--fast-tf32', action='store_true', help='single-pass TF32 contractions (not parity-grade)')
    ap.add_argument('--precision', type=int, default=-1, help='xm_set_precision value (default: library default)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-breakdown', action='store_true')
    ap.add_argument('--no-weak', action='store_true', help='skip the extra weak-scaling leg of a strong N>1 run')
    return ap.parse_args()


def config_dict(cfg, args, world, tasks_per_gpu, scaling):
    """The ``config`` object of the JSON line -- identical for our arm and the reference arm."""
    T = args.inner_steps or cfg['steps']
    return {'workload': cfg['workload'], 'baseline_config': args.config, 'global_meta_batch': tasks_per_gpu * world,
            'inner_steps': T, 'inner_lr': cfg['inner_lr'], 'ways': cfg['ways'], 'shots': cfg['shots'],
            'order': 'second' if cfg['algo'] == 'maml' else 'anil (second-order head adaptation)'}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'tensor_tflops': d['bf16_tflops_sustained'], 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tensor_tflops': 1400.0, 'source': 'fallback'}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, tasks, reps, inner_steps, threads=None):
    """The reference's CPU path (oracle port: reference model + fast_adapt semantics + learn2learn restatement, fp32,
    torch intra-op threads = all host cores unless given), train tasks only incl. the second-order backward, grad/B
    and the Adam step; returns tasks/s over ``reps`` timed repetitions of a ``tasks``-task sample after a 1-task
    warm-up."""
    import torch
    from oracle import maml_oracle as mo
    from exploring_meta_b200.synthetic import make_tasks
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    ways, shots = cfg['ways'], cfg['shots']
    head = None
    if cfg['algo'] == 'anil':
        ospec = mo.NetSpec(3, 84, 84, 64, ways, 4, True, 'flatten')
        params, head = mo.init_anil_params(ospec, seed=42)
    else:
        ospec = mo.omniglot_spec(ways) if cfg['kind'] == 'omni' else mo.miniimagenet_spec(ways)
        params = mo.init_params(ospec, seed=42)
    X, Y = make_tasks(tasks, ways, shots, IN_SHAPE[cfg['kind']], seed=0)
    mo.meta_iteration(params, X[:1], Y[:1], ospec, inner_steps, cfg['inner_lr'], anil_head=head)     # warm-up
    state = mo.new_adam_state(params)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = mo.meta_iteration(params, X, Y, ospec, inner_steps, cfg['inner_lr'], anil_head=head)
        mo.adam_step(params, [g / tasks for g in out['grad']], state, lr=OUTER_LR)
        times.append(time.perf_counter() - t0)
    return tasks / (sum(times) / len(times)), cores, times


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    T = args.inner_steps or cfg['steps']
    world = int(os.environ.get('WORLD_SIZE', '1'))
    sample_tasks = cfg['batch'] if args.config == 1 else 4         # cfg 1 whole batch, else 4 tasks
    rate, cores, times = cpu_reference_rate(cfg, sample_tasks, max(1, args.steps), T)
    ms = 1000.0 * sum(times) / len(times)
    sample = ('%d-task sample of the %d-task meta-batch per step (1-task warm-up), fp32, %d torch threads, oracle port '
              'of the reference path' % (sample_tasks, cfg['batch'], cores))
    per = cfg['batch'] // world if args.scaling == 'strong' else cfg['batch']
    line = {
        'impl': 'reference', 'metric': cfg['metric'], 'value': rate, 'unit': 'tasks/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': config_dict(cfg, args, world, per, args.scaling),
        'cpu_baseline': {'value': rate, 'unit': 'tasks/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'tasks/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def family(name, cargs):
    """Kernel family of one C-ABI call: the image layer (cin <= 4) runs on different kernels (CUDA-core FFMA2,
    HBM-bound) than the 32-channel layers (tcgen05), so they are reported separately."""
    if name in ('xm_conv', 'xm_wgrad') and cargs.g.cin <= 4:
        return name + ':image_layer'
    return name


def kernel_breakdown(engine, reps=2):
    """Per-entry-point device time of one launch program, CUDA events on the launching stream around
    every C-ABI call (un-captured replay).  Returns {name: {'ms': per-step ms, 'calls': n}}."""
    import ctypes
    import torch
    from exploring_meta_b200 import _lib
    stream = torch.cuda.current_stream()
    calls = engine.prog.calls
    totals = {}
    calls = [(fn, cargs, family(name, cargs)) for fn, cargs, name in calls]
    for rep in range(reps + 1):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in calls]
        for (fn, cargs, name), (e0, e1) in zip(calls, evs):
            e0.record(stream)
            code = fn(*cargs, stream.cuda_stream) if isinstance(cargs, tuple) else fn(ctypes.byref(cargs), stream.cuda_stream)
            _lib.check(code, name.split(':')[0])
            e1.record(stream)
        torch.cuda.synchronize()
        if rep == 0:
            continue
        for idx, ((fn, cargs, name), (e0, e1)) in enumerate(zip(calls, evs)):
            d = totals.setdefault(name, {'ms': 0.0, 'calls': 0, 'per_call': {}})
            ms = e0.elapsed_time(e1)
            d['ms'] += ms / reps
            d['calls'] += 1 if rep == 1 else 0
            d['per_call'][idx] = d['per_call'].get(idx, 0.0) + ms / reps
    return totals


def program_work(engine):
    """Algorithmic work per entry point of one launch program: conv/wgrad FLOPs (2/MAC)
    and compulsory bytes of the streaming BN kernels (each tensor touched once)."""
    from exploring_meta_b200 import _lib
    work = {}
    for idx, (fn, a, name) in enumerate(engine.prog.calls):
        fam = family(name, a)
        w = work.setdefault(fam, {'flops': 0.0, 'bytes': 0.0, 'per_call': {}})
        if name in ('xm_conv', 'xm_wgrad'):
            g = a.g
            pairs = 2 if (getattr(a, 'src2', None) or getattr(a, 'x2', None)) else 1
            fl = 2.0 * g.tasks * g.n * g.hz * g.wz * 9 * g.cin * g.cout * pairs
            w['flops'] += fl
            w['per_call'][idx] = fl
            if fam.endswith(':image_layer'):
                # HBM-bound: the z-sized tensor (written by the conv / read by the wgrad) + the images
                by = 4.0 * g.tasks * g.n * (g.hz * g.wz * g.cout + g.hin * g.win * g.cin)
                if name == 'xm_conv' and a.stat_mode == 2:
                    by += 4.0 * g.tasks * g.n * g.hz * g.wz * g.cout          # aux (z) read by the tangent pass
                w['bytes'] += by
        elif name.startswith('xm_img'):
            g = a.g
            K = 9 * g.cin
            x = 4.0 * g.tasks * g.n * g.cin * g.hin * g.win
            pe = 1.0 * g.tasks * g.n * g.hp * g.wp * g.cout            # pooled elements
            dense = 2.0 * g.tasks * g.n * g.hz * g.wz * K * g.cout     # the conv itself (CUDA-core FFMA2)
            sparse = 2.0 * pe * K                                      # one 3x3xcin patch per pooling winner
            fl, by = {'xm_img_gram': (2.0 * g.tasks * g.n * g.hz * g.wz * (K * (K + 1) / 2 + K), x),
                      'xm_img_fwd': (dense, x + 9 * pe),               # writes p, zsel (fp32), sel (u8)
                      'xm_img_dual_fwd': (dense, x + 13 * pe),         # reads zsel, sel; writes pdot, zdsel
                      'xm_img_bwd': (sparse, x + 9 * pe),              # reads gp, zsel, sel
                      'xm_img_dual_bwd': (sparse, x + 17 * pe)}[name]  # reads gp, gpdot, zsel, zdsel, sel
            w['flops'] += fl
            w['bytes'] += by
            w['per_call'][idx] = by
        elif name.startswith('xm_bn'):
            g = a.g
            z = 4.0 * g.tasks * g.n * g.hz * g.wz * g.cout
            p = 4.0 * g.tasks * g.n * g.hp * g.wp * g.cout
            by = {'xm_bn_fwd': z + p, 'xm_bn_bwd': 2 * z + p, 'xm_bn_dual_fwd': 2 * z + p,
                  'xm_bn_dual_bwd': 4 * z + 2 * p}[name]
            w['bytes'] += by
            w['per_call'][idx] = by
    return work


def build_trainer(cfg, tasks, T, dev, fast_tf32=False):
    """Trainer + flat initial parameters of one BASELINE config at ``tasks`` tasks per GPU."""
    import torch
    from exploring_meta_b200 import spec as pspec
    from exploring_meta_b200.trainer import AnilTrainer, MamlTrainer
    ways, shots = cfg['ways'], cfg['shots']
    if cfg['algo'] == 'anil':
        spec = pspec.anil_body_spec(cfg['kind'], ways)
        tr = AnilTrainer(spec, tasks, shots, T, cfg['inner_lr'], OUTER_LR, device=dev, use_graph=True)
        body = pspec.init_flat_params(spec, seed=42)              # leaves the RNG after the blocks, like the reference
        head = torch.nn.Linear(tr.engine.D, ways)
        tr.theta_all.copy_(torch.cat([body, head.weight.detach().reshape(-1), head.bias.detach().reshape(-1)]))
    else:
        spec = pspec.omniglot_spec(ways) if cfg['kind'] == 'omni' else pspec.miniimagenet_spec(ways)
        tr = MamlTrainer(spec, tasks, shots, T, cfg['inner_lr'], OUTER_LR, device=dev, use_graph=True)
        tr.theta.copy_(pspec.init_flat_params(spec, seed=42))
    return spec, tr


def flops_per_task(cfg, spec, T):
    if cfg['algo'] == 'anil':      # body forward + first-order backward over 2S rows (27.93 GFLOP)
        m = [hz * wz * spec.hidden * cin * 9 for (cin, _h, _w, hz, wz, _hp, _wp) in spec.block_dims()]
        rows = 2 * cfg['ways'] * cfg['shots']
        return 2.0 * rows * (2 * m[0] + 3 * sum(m[1:]))
    return float(spec.flops_per_train_task(cfg['shots'], T))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from exploring_meta_b200 import _lib
    from exploring_meta_b200.synthetic import make_tasks

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: exploring_meta_b200 has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG=VERSION / WARN) to stdout when the communicator is created: point
        # fd 1 at stderr until then so that stdout carries exactly the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    lib = _lib.load()
    if args.fast_tf32:
        lib.xm_set_precision(0)
    elif args.precision >= 0:
        lib.xm_set_precision(args.precision)

    cfg = CONFIGS[args.config]
    T = args.inner_steps or cfg['steps']
    ways, shots, shape = cfg['ways'], cfg['shots'], IN_SHAPE[cfg['kind']]
    if args.tasks:
        per = args.tasks
    elif args.scaling == 'strong':
        if cfg['batch'] % world:
            raise SystemExit('global meta-batch %d is not divisible by %d ranks' % (cfg['batch'], world))
        per = cfg['batch'] // world
    else:
        per = cfg['batch']
    global_tasks = per * world

    def host_batches(per_gpu):
        """Two synthetic meta-batches in pinned host memory.  Strong scaling: rank r's shard [r*per, (r+1)*per) of
        the SAME seeded global batch a single GPU would process; weak: one batch per rank."""
        out = []
        for b in range(2):
            if args.scaling == 'strong' and per_gpu == per and not args.tasks:
                X, Y = make_tasks(global_tasks, ways, shots, shape, seed=b)
                X, Y = X[rank * per:(rank + 1) * per].contiguous(), Y[rank * per:(rank + 1) * per].contiguous()
            else:
                X, Y = make_tasks(per_gpu, ways, shots, shape, seed=rank * 1000 + b)
            out.append((X.pin_memory(), Y.pin_memory()))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_resident(tr, steps, warmup, clocks=False):
        """W warm-up + K timed device-resident steps (whole step = one CUDA graph), max over ranks."""
        for _ in range(warmup):
            tr.step_resident()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0 and clocks:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            tr.step_resident()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), (sampler.stop() if (rank == 0 and clocks) else None)

    mem0 = torch.cuda.memory_allocated(dev)
    spec, tr = build_trainer(cfg, per, T, dev)
    working_set_mb = (torch.cuda.memory_allocated(dev) - mem0) / 1e6
    host = host_batches(per)
    e = tr.engine
    e.x.copy_(host[0][0]); e.y.copy_(host[0][1])
    l0 = int(lib.xm_launch_count())
    tr.step_eager()                                                 # un-captured step: counts OUR kernels per step
    torch.cuda.synchronize()
    kernels_per_step = int(lib.xm_launch_count()) - l0
    warmup = max(args.warmup, 3)

    def throttled(c):
        bad = {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'} & set(c.get('reasons') or [])
        stuck = (c.get('sm_mhz') and c.get('sm_max_mhz') and not c.get('reasons')
                 and c['sm_mhz'] < 0.75 * c['sm_max_mhz'])          # clocks well below max with no reason: leftover lock
        return bool(bad or stuck)

    total_ms, clocks = time_resident(tr, args.steps, warmup, clocks=True)
    # a measurement taken under a thermal / hardware slowdown (or a leftover clock lock) is rejected and taken once more
    redo = torch.tensor([1 if (rank == 0 and throttled(clocks)) else 0], device=dev)
    if world > 1:
        dist.all_reduce(redo, op=dist.ReduceOp.MAX)
    if int(redo.item()):
        first = clocks
        total_ms, clocks = time_resident(tr, args.steps, 1, clocks=True)
        if rank == 0:
            clocks['remeasured_after'] = first

    # ---- end to end through the public API: every step copies ITS pinned host batch to the device (stage(), a side
    # stream: the copy of batch i+1 overlaps the adaptation of batch i) and reads its loss / accuracy back ---------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tr.stage(*host[0])
    for i in range(2):
        tr.stage(*host[(i + 1) % 2])
        loss, correct = tr.meta_step()
        loss.cpu()
    barrier()
    ev0.record()
    for i in range(args.steps):
        tr.stage(*host[(i + 1) % 2])              # one host -> device batch copy per step, inside the timed region
        loss, correct = tr.meta_step()            # consumes the batch staged for this step
        loss_h, correct_h = loss.cpu(), correct.cpu()
    ev1.record()
    barrier()
    ms2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2.item())
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8
    d2h = per * 8
    if tr.comm is not None:
        tr.comm.check()

    # ---- weak-scaling leg of a strong N > 1 run (kept as an extra key) ------------------------------------------
    weak = None
    if world > 1 and args.scaling == 'strong' and not args.tasks and not args.no_weak:
        del tr, e
        torch.cuda.empty_cache()
        _spec, trw = build_trainer(cfg, cfg['batch'], T, dev)
        Xw, Yw = make_tasks(cfg['batch'], ways, shots, shape, seed=rank * 1000)
        trw.engine.x.copy_(Xw); trw.engine.y.copy_(Yw)
        wms, _ = time_resident(trw, args.steps, warmup)
        weak = {'tasks_per_gpu': cfg['batch'], 'global_meta_batch': cfg['batch'] * world,
                'value': cfg['batch'] * world * args.steps / (wms / 1000.0), 'unit': 'tasks/s',
                'ms_per_step': wms / args.steps}
        tr = trw
        e = trw.engine

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = global_tasks * args.steps / (total_ms / 1000.0)
    e2e = global_tasks * args.steps / (e2e_ms / 1000.0)
    peaks = measured_peaks()
    fpt = flops_per_task(cfg, spec, T)
    in_mb = h2d / 1e6
    cd = config_dict(cfg, args, world, per, args.scaling)
    run_info = {'tasks_per_gpu': per, 'working_set_mb_per_gpu': round(working_set_mb, 1),
               'parallelism': ('tasks sharded over %d GPU(s); the %d-float [meta-grad ; loss ; correct ; BN partials] '
                               'buffers are summed over NVLink peer memory inside the Adam kernel (no NCCL call)'
                               % (world, tr.flat.numel())) if world > 1 else 'single GPU',
               'l2_policy': 'no explicit flush: every step streams its inputs (%.1f MB/GPU) and the saved activations of '
                            'all inner steps (static buffers: %.0f MB/GPU, each written and re-read once per step) '
                            'through the 126 MB L2' % (in_mb, working_set_mb),
               'cuda_graph': 'whole step (launch program + shard sums + allreduce + Adam + BN EMA) is one graph'}
    line = {
        'metric': cfg['metric'], 'value': value, 'unit': 'tasks/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f32 (tensor-core contractions: %s)' % ('1xTF32' if args.fast_tf32 else '3xTF32 error-compensated, fp32 accumulate'),
        'data': 'synthetic',
        'config': cd, 'run': run_info,
        'e2e': {'value': e2e, 'unit': 'tasks/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_ms / args.steps},
        'gpu_launches': kernels_per_step * args.steps,
        'clocks': clocks,
        'algorithmic_tflops': value * fpt / 1e12,
        'flops_per_task': fpt,
    }
    if weak is not None:
        line['weak_scaling'] = weak

    if world == 1 and not args.no_kernel_breakdown:
        times = kernel_breakdown(e)
        work = program_work(e)
        step_ms = sum(v['ms'] for v in times.values())
        fams = {}
        for name, tv in sorted(times.items(), key=lambda kv: -kv[1]['ms']):
            w = work.get(name, {'flops': 0.0, 'bytes': 0.0})
            fam = {'ms_per_step': round(tv['ms'], 3), 'calls': tv['calls'], 'share': round(tv['ms'] / step_ms, 4)}
            if w['flops']:
                fam['tflops'] = round(w['flops'] / tv['ms'] / 1e9, 2)
            if w['bytes']:
                fam['gbs'] = round(w['bytes'] / tv['ms'] / 1e6, 1)
            fams[name] = fam
        line['kernels'] = fams
        def hbm_bound(name):
            return (name.startswith('xm_bn') or name.endswith(':image_layer')
                    or name in ('xm_img_bwd', 'xm_img_dual_bwd'))

        # exact-fp32 CUDA-core kernels (image-block forward conv, Gram matrix): bound by the FFMA / DFMA pipes
        fp32_peak = 148 * 128 * 2 * (clocks.get('sm_max_mhz') or 1965.0) * 1e6 / 1e12

        for name, fam in fams.items():
            if hbm_bound(name) and 'gbs' in fam:
                fam['roofline_frac'] = round(fam['gbs'] / peaks['hbm_gbs'], 4)
            elif name in ('xm_img_fwd', 'xm_img_dual_fwd') and 'tflops' in fam:
                fam['bound'] = 'fp32 CUDA cores (nominal %.1f TFLOP/s)' % fp32_peak
                fam['roofline_frac'] = round(fam['tflops'] / fp32_peak, 4)
            elif name == 'xm_img_gram':
                fam['bound'] = 'fp64 CUDA cores'
            elif 'tflops' in fam:
                fam['roofline_frac'] = round(fam['tflops'] / peaks['tensor_tflops'], 4)
        traffic = {}
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')     # written by scripts/ncu_digest.py
        if os.path.isfile(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
        top = max(times.items(), key=lambda kv: kv[1]['ms'])[0]
        tv, w = times[top], work.get(top, {'flops': 0.0, 'bytes': 0.0})
        trf = traffic.get(top, {}).get('dram_bytes_per_launch') if args.config == 2 else None
        if hbm_bound(top):
            achieved = w['bytes'] / tv['ms'] / 1e6
            line['roofline'] = {'kernel': top, 'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'],
                                'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'], 'traffic': trf,
                                'traffic_source': traffic.get('_meta'),
                                'peak_source': peaks['source'] + ' copy bandwidth (MEASURED_PEAKS.json)',
                                'algorithmic_bytes_per_launch': w['bytes'] / tv['calls'],
                                'launches_per_step': tv['calls'], 'avg_launch_ms': tv['ms'] / tv['calls']}
        else:
            achieved = w['flops'] / tv['ms'] / 1e9
            line['roofline'] = {'kernel': top, 'bound': 'tensor', 'achieved': achieved, 'peak': peaks['tensor_tflops'],
                                'unit': 'TFLOP/s', 'frac': achieved / peaks['tensor_tflops'], 'traffic': trf,
                                'traffic_source': traffic.get('_meta'),
                                'peak_source': peaks['source'] + ' bf16 dense sustained (MEASURED_PEAKS.json); the '
                                'kernel computes fp32-grade products as 3 TF32 tcgen05.mma each (TF32 dense = 1/2 of the bf16 '
                                'rate), so its own ceiling is 1/6 of this peak (DESIGN.md section 4)',
                                'algorithmic_flops_per_launch': w['flops'] / tv['calls'],
                                'launches_per_step': tv['calls'], 'avg_launch_ms': tv['ms'] / tv['calls']}
    if world == 1 and not args.no_cpu_baseline:
        n_sample = cfg['batch'] if args.config == 1 else 4
        rate, cores, times = cpu_reference_rate(cfg, n_sample, 3 if args.config == 1 else 2, T)
        rate1, _c, _t = cpu_reference_rate(cfg, 1, 1, T, threads=1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'tasks/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d-task sample x %d repetitions after a 1-task warm-up, fp32, %d torch '
                                          'threads (oracle port of the reference path, BASELINE.md 3.4)'
                                          % (n_sample, len(times), cores),
                                'one_thread': {'value': rate1, 'unit': 'tasks/s', 'cores': 1,
                                               'sample': '1 task x 1 repetition after a 1-task warm-up'}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def rl_cpu_rate(tasks, reps, threads=None):
    """Oracle port of rl/maml_trpo.py:101-134 (first-order adaptation per task + meta_optimize_trpo) on ``tasks``
    tasks, fp32, all host cores; tasks/s."""
    import torch
    from oracle import rl_oracle as ro
    from exploring_meta_b200.synthetic import make_replays
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = {k: RL[k] for k in ('inner_lr', 'tau', 'gamma', 'value_reg', 'max_kl', 'ls_max_steps', 'backtrack_factor', 'outer_lr')}
    theta = ro.init_policy(seed=42)
    data = make_replays(tasks, RL['episodes'], RL['horizon'], seed=0)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        old = [[x.detach() for x in ro.trpo_update([p.clone().requires_grad_() for p in theta], sup, cfg['inner_lr'],
                                                   cfg['tau'], cfg['gamma'], cfg['value_reg'], first_order=True)]
               for sup, _q in data]
        ro.meta_optimize_trpo(theta, [[s_, q_] for s_, q_ in data], old, cfg)
        times.append(time.perf_counter() - t0)
    return tasks / (sum(times) / len(times)), cores, times


def rl_config(args):
    return {'workload': RL['workload'], 'baseline_config': 5, 'global_meta_batch': RL['tasks'],
            'transitions_per_replay': RL['episodes'] * RL['horizon'], 'inner_lr': RL['inner_lr'], 'max_kl': RL['max_kl']}


def run_rl_reference(args):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    sample = 4
    rate, cores, times = rl_cpu_rate(sample, max(1, min(args.steps, 3)))
    desc = '%d-task sample of the 40-task meta-batch per step, fp32, %d torch threads, oracle port' % (sample, cores)
    print(json.dumps({'impl': 'reference', 'metric': RL['metric'], 'value': rate, 'unit': 'tasks/s', 'n_gpus': args.gpus,
                      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * sum(times) / len(times),
                      'higher_is_better': True, 'scaling': 'replicas only', 'vs_baseline': None, 'dtype': 'f32',
                      'data': 'synthetic', 'config': rl_config(args),
                      'cpu_baseline': {'value': rate, 'unit': 'tasks/s', 'cores': cores, 'kind': 'port', 'sample': desc},
                      'e2e': {'value': rate, 'unit': 'tasks/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                      'gpu_launches': 0}), flush=True)


def run_rl(args):
    """Config 5 on one GPU (the path does not shard in the reference's configuration: N > 1 = replicas only)."""
    import torch
    from exploring_meta_b200 import _lib
    from exploring_meta_b200.rl_engine import TrpoEngine
    from exploring_meta_b200.synthetic import make_replays
    from oracle import rl_oracle as ro
    if int(os.environ.get('RANK', '0')) != 0:
        return
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    lib = _lib.load()
    B, n = RL['tasks'], RL['episodes'] * RL['horizon']
    e = TrpoEngine(B, n, 2, 2, (100, 100), 'tanh', RL['inner_lr'], RL['gamma'], RL['tau'], RL['value_reg'], device=dev)
    theta = torch.cat([p.reshape(-1) for p in ro.init_policy(seed=42)]).to(dev)
    host = []
    for b in range(2):
        data = make_replays(B, RL['episodes'], RL['horizon'], seed=b)
        host.append({k: [torch.stack([t[j][k].reshape(n, -1) for t in data]).pin_memory() for j in range(2)]
                     for k in ('states', 'actions', 'rewards', 'dones', 'next_states')})
    fields = {'states': e.states, 'actions': e.actions, 'rewards': e.rewards, 'dones': e.dones, 'next_states': e.next_states}

    def stage(batch):
        for k, dst in fields.items():
            for j in range(2):
                dst[j].copy_(batch[k][j].view_as(dst[j]), non_blocking=True)

    def step(th):
        e.prepare()                                           # advantages of all 80 replays
        old = e.adapt(th).clone()                             # fast_adapt_trpo(first_order=True) for every task
        e.set_old_policies(old)
        new, diag = e.meta_optimize(th, RL['max_kl'], RL['ls_max_steps'], RL['backtrack_factor'], RL['outer_lr'])
        return new, diag

    stage(host[0])
    l0 = int(lib.xm_launch_count())
    _new, diag = step(theta)
    torch.cuda.synchronize()
    launches = int(lib.xm_launch_count()) - l0
    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step(theta)
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(theta)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    ev0.record()
    for i in range(args.steps):
        stage(host[i % 2])
        new, d_ = step(theta)
        new_h = new.cpu()
    ev1.record()
    torch.cuda.synchronize()
    ms2 = ev0.elapsed_time(ev1) / args.steps
    h2d = sum(t.numel() * 4 for k in host[0] for t in host[0][k])
    line = {'metric': RL['metric'], 'value': B / (ms / 1e3), 'unit': 'tasks/s', 'n_gpus': 1, 'steps': args.steps,
            'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'replicas only',
            'vs_baseline': None, 'dtype': 'f32 (exact fp32 CUDA-core products; advantages in f64)', 'data': 'synthetic',
            'config': rl_config(args),
            'run': {'line_search_step': diag['ls_step'], 'l2_policy': 'latency-bound: %d launches of <= 1.7 GFLOP per step; '
                    'working set (replays 5 MB, per-task vectors 7 MB) is L2-resident by nature of the workload' % launches},
            'e2e': {'value': B / (ms2 / 1e3), 'unit': 'tasks/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': e.P * 4,
                    'ms_per_step': ms2},
            'gpu_launches': launches * args.steps, 'clocks': clocks}
    if not args.no_cpu_baseline:
        rate, cores, times = rl_cpu_rate(4, 1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'tasks/s', 'cores': cores, 'kind': 'port',
                                'sample': '4-task sample x 1 repetition, fp32, %d torch threads (oracle port of '
                                          'rl/maml_trpo.py:101-134)' % cores}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    a = parse()
    if a.config == 5:
        (run_rl_reference if a.impl == 'reference' else run_rl)(a)
    elif a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
