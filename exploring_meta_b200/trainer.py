"""Task-batched meta-training step: the body of the reference's outer loop as one call.

``MamlTrainer.meta_step(x, y)`` replaces ``vision/maml_vision.py:95-141`` minus the validation pass:
``opt.zero_grad()``; for every task ``maml.clone()`` -> ``fast_adapt`` -> ``eval_loss.backward()``;
``p.grad *= 1/meta_batch_size``; ``Adam.step()``.  The per-task Python loop becomes one replay of a
captured launch program over the rank's shard of tasks; with ``world_size > 1`` the shard gradients
are combined by ONE sum-allreduce of the flat fp32 buffer [meta-grad ; loss sum ; correct count]
(NCCL over NVLink) before the identical, replicated Adam step.  Nothing here synchronises the host.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from .engine import AnilEngine, MamlEngine, _p

ADAM_BETAS = (0.9, 0.999)
ADAM_EPS = 1e-8


class _TrainerBase:
    def _init_common(self, device, outer_lr, total_params, use_graph, extra=0):
        self.device = torch.device(device)
        self.lib = _lib.load()
        self.outer_lr = float(outer_lr)
        self.iteration = 0
        self.use_graph = bool(use_graph)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # flat fp32 buffer that is summed over ranks: [grad (total_params) ; loss sum ; correct count ; BN EMA partials
        # (extra)].  ``flat`` is this rank's contribution, ``red`` the sum (what Adam and metrics() read).
        self.flat = torch.zeros(total_params + 2 + extra, dtype=torch.float32, device=self.device)
        self.red = torch.zeros_like(self.flat)
        self.m = torch.zeros(total_params, dtype=torch.float32, device=self.device)
        self.v = torch.zeros(total_params, dtype=torch.float32, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)    # Adam step count, advanced on device
        self._graphs = {}
        # transport of the shard sums: 'p2p' = inside the Adam kernel over NVLink peer memory (csrc/comm.cu), 'dist' =
        # torch.distributed all_reduce (what the gloo / CPU-emulator tests use; XM_COMM=nccl selects it on GPUs)
        self.comm = None
        if self.world > 1 and self.device.type == 'cuda' and os.environ.get('XM_COMM', 'p2p') == 'p2p':
            from .comm import PeerComm
            self.comm = PeerComm(self.flat.numel(), self.device)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == 'cuda' else 0

    # ---- input pipeline: host batch i+1 is copied to the device while batch i is being adapted -----------------
    def stage(self, x_host, y_host):
        """Starts the asynchronous host->device copy of a (pinned) batch into a staging slot on a side stream.
        ``meta_step()`` without arguments then consumes the oldest staged batch.  Two slots: at most two batches
        may be staged ahead of their ``meta_step``."""
        e = self.engine
        if getattr(self, '_stage', None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._stage = [(torch.empty_like(e.x), torch.empty_like(e.y)) for _ in range(2)]
            self._staged = [torch.cuda.Event() for _ in range(2)]      # H2D into slot complete
            self._consumed = [torch.cuda.Event() for _ in range(2)]    # slot copied into the engine inputs
            for ev in self._consumed:
                ev.record(torch.cuda.current_stream(self.device))
            self._head = self._tail = 0
        slot = self._tail % 2
        self._tail += 1
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._consumed[slot])
            self._stage[slot][0].copy_(x_host.view_as(e.x), non_blocking=True)
            self._stage[slot][1].copy_(y_host, non_blocking=True)
            self._staged[slot].record(self._copy_stream)

    def _take_staged(self):
        e = self.engine
        assert getattr(self, '_stage', None) is not None and self._head < self._tail, 'no staged batch'
        slot = self._head % 2
        self._head += 1
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._staged[slot])
        e.x.copy_(self._stage[slot][0], non_blocking=True)
        e.y.copy_(self._stage[slot][1], non_blocking=True)
        self._consumed[slot].record(cur)

    def _reduce_and_step(self, theta_all, global_tasks):
        """Shard sums -> (sum over ranks) -> Adam, all on the current stream and graph-capturable: the Adam step
        number lives in device memory (``step_dev``)."""
        e, n = self.engine, theta_all.numel()
        stream = self._stream()
        _lib.check(self.lib.xm_finish_shard(_p(e.loss), e.correct.data_ptr(), self.tasks, _p(self.flat, n),
                                            self.step_dev.data_ptr(), stream), 'xm_finish_shard')
        local = self.flat
        if self.world > 1 and self.comm is None:
            self.red.copy_(self.flat)
            dist.all_reduce(self.red, op=dist.ReduceOp.SUM)
            local = self.red
        a = _lib.XmAdamArgs()
        a.theta, a.m, a.v, a.n_params = _p(theta_all), _p(self.m), _p(self.v), n
        a.local, a.reduced, a.n_total = _p(local), _p(self.red), self.flat.numel()
        a.grad_scale, a.lr = 1.0 / global_tasks, self.outer_lr
        a.beta1, a.beta2, a.eps = ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS
        a.step = self.step_dev.data_ptr()
        self._adam_args = a                       # keep the block alive
        _lib.check(self.lib.xm_allreduce_adam(self.comm.ptr if self.comm is not None else None, ctypes.byref(a),
                                              stream), 'xm_allreduce_adam')

    def _run(self, key, body):
        """Runs ``body()`` (engine program + outer step) -- through a CUDA graph captured on first use when enabled."""
        if not (self.use_graph and self.device.type == 'cuda'):
            return body()
        g = self._graphs.get(key)
        if g is None:
            e = self.engine
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                e.prog.replay(side.cuda_stream)                # warm-up (idempotent): loads every kernel of the program
                self._warm_outer(side.cuda_stream)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = g
        g.replay()

    def step_resident(self):
        """One meta-iteration on the batch already in ``engine.x`` / ``engine.y`` (no input copy)."""
        self._run('resident', self._step_body)
        self._count_step()

    def step_eager(self):
        """Same, launching kernel by kernel (no graph): used to count launches and by per-kernel timing."""
        self._step_body()
        self._count_step()

    def _count_step(self):
        self.iteration += 1

    def _warm_outer(self, stream):
        """Loads the outer-step kernels on throw-away buffers (a warm-up of the real ones would apply an update)."""
        t = torch.zeros(8, dtype=torch.float32, device=self.device)
        c = torch.zeros(1, dtype=torch.int32, device=self.device)
        st = torch.ones(1, dtype=torch.int32, device=self.device)
        _lib.check(self.lib.xm_finish_shard(_p(t), c.data_ptr(), 1, _p(t, 4), st.data_ptr(), stream), 'xm_finish_shard')
        a = _lib.XmAdamArgs()
        a.theta, a.m, a.v, a.n_params = _p(t), _p(t, 2), _p(t, 4), 2
        a.local, a.reduced, a.n_total = _p(t, 6), _p(t, 6), 2
        a.grad_scale, a.lr, a.beta1, a.beta2, a.eps, a.step = 1.0, 0.0, 0.9, 0.999, 1e-8, st.data_ptr()
        _lib.check(self.lib.xm_allreduce_adam(None, ctypes.byref(a), stream), 'xm_allreduce_adam')
        torch.cuda.synchronize(self.device)


class MamlTrainer(_TrainerBase):
    """MAML meta-training over a shard of ``tasks`` tasks per rank.

    ``theta`` [P] holds the master parameters in ``module.parameters()`` order; it is updated in place
    by every ``meta_step``.  ``metrics()`` returns (mean query loss, mean query accuracy) over the
    global meta-batch of the last step as device scalars."""

    def __init__(self, spec, tasks, shots, steps, inner_lr, outer_lr=0.003, first_order=False,
                 device='cuda', use_graph=True):
        self.engine = MamlEngine(spec, tasks, shots, steps, inner_lr,
                                 mode='first' if first_order else 'second', device=device)
        self.spec, self.tasks = spec, int(tasks)
        # sharded runs carry the BatchNorm running-statistics side effect in the same allreduce: per layer the rank's
        # composed EMA contribution {mean[C], var[C]}
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._init_common(device, outer_lr, self.engine.P, use_graph,
                          extra=2 * spec.layers * spec.hidden if world > 1 else 0)
        self.theta = self.engine.theta
        self.running_mean = [torch.zeros(spec.hidden, device=self.device) for _ in range(spec.layers)]
        self.running_var = [torch.ones(spec.hidden, device=self.device) for _ in range(spec.layers)]
        self.num_batches_tracked = 0
        # the engine accumulates straight into the all-reduce buffer
        self.engine.grad = self.flat[:self.engine.P]
        self.engine.rebuild()

    def load_parameters(self, params):
        """params: iterable of tensors in ``parameters()`` order (e.g. a reference ``MiniImagenetCNN``)."""
        flat = torch.cat([p.detach().reshape(-1).float() for p in params])
        assert flat.numel() == self.engine.P
        self.theta.copy_(flat)

    def meta_step(self, x=None, y=None, track_running_stats=True):
        """One meta-iteration.  ``x, y``: the shard's tasks (host or device tensors), or None to consume the batch
        queued by ``stage()``."""
        e = self.engine
        if x is None:
            self._take_staged()
        else:
            e.x.copy_(x, non_blocking=True)
            e.y.copy_(y, non_blocking=True)
        key = 'resident' if track_running_stats else 'no_stats'
        self._run(key, lambda: self._step_body(track_running_stats))
        self._count_step(track_running_stats)
        return e.loss, e.correct

    def _count_step(self, track_running_stats=True):
        self.iteration += 1
        if track_running_stats:
            self.num_batches_tracked += self.tasks * (self.engine.steps + 1) * self.world

    def _step_body(self, track_running_stats=True):
        e = self.engine
        e.replay()
        if track_running_stats and self.world == 1:
            e.update_running_stats(self.running_mean, self.running_var)
        elif track_running_stats:
            self._stage_running_stats()
        self._reduce_and_step(self.theta, self.tasks * self.world)
        if track_running_stats and self.world > 1:
            self._apply_running_stats()

    # The reference's shared BN buffers see one EMA update r <- (1-m) r + m s per forward call, task after task
    # (vision/maml_vision.py:102-112 through learn2learn's clones).  The recurrence is linear: over the N calls of the
    # global meta-batch r_N = (1-m)^N r_0 + sum_k m (1-m)^(N-k) s_k, and the sum splits by rank: rank q contributes the
    # EMA of ITS calls started from zero, damped by (1-m)^(calls of the ranks after it).  Those contributions ride in
    # the meta-gradient allreduce; every rank then applies the same closed form.
    def _stage_running_stats(self):
        e, L, C = self.engine, self.spec.layers, self.spec.hidden
        rank = dist.get_rank()
        calls = self.tasks * (e.steps + 1)                     # forward calls of one rank per layer
        tail = self.flat[e.P + 2:].view(L, 2, C)
        tail.zero_()
        e.update_running_stats([tail[l, 0] for l in range(L)], [tail[l, 1] for l in range(L)])
        from .engine import BN_MOMENTUM
        tail.mul_((1.0 - BN_MOMENTUM) ** (calls * (self.world - 1 - rank)))

    def _apply_running_stats(self):
        e, L, C = self.engine, self.spec.layers, self.spec.hidden
        from .engine import BN_MOMENTUM
        calls = self.tasks * (e.steps + 1) * self.world
        decay = (1.0 - BN_MOMENTUM) ** calls
        tail = self.red[e.P + 2:].view(L, 2, C)
        for l in range(L):
            self.running_mean[l].mul_(decay).add_(tail[l, 0])
            self.running_var[l].mul_(decay).add_(tail[l, 1])

    def metrics(self):
        P, n = self.engine.P, self.tasks * self.world
        return self.red[P] / n, self.red[P + 1] / (n * self.engine.S)


class AnilTrainer(_TrainerBase):
    """ANIL meta-training (``vision/anil_vision.py:109-151``): body + head trained by one Adam over
    ``list(features.parameters()) + list(head.parameters())`` (:98-99)."""

    def __init__(self, spec, tasks, shots, steps, inner_lr, outer_lr=0.003, first_order=False,
                 device='cuda', use_graph=True):
        self.engine = AnilEngine(spec, tasks, shots, steps, inner_lr, first_order=first_order, device=device)
        self.spec, self.tasks = spec, int(tasks)
        e = self.engine
        self._init_common(device, outer_lr, e.P + e.PH, use_graph)
        # body and head parameters live in one flat vector so that one Adam launch updates both
        self.theta_all = torch.zeros(e.P + e.PH, dtype=torch.float32, device=self.device)
        e.theta = self.theta_all[:e.P]
        e.head = self.theta_all[e.P:]
        e.grad = self.flat[:e.P]
        e.head_grad = self.flat[e.P:e.P + e.PH]
        e.rebuild()

    def load_parameters(self, body_params, head_params):
        flat = torch.cat([p.detach().reshape(-1).float() for p in list(body_params) + list(head_params)])
        assert flat.numel() == self.theta_all.numel()
        self.theta_all.copy_(flat)

    def meta_step(self, x=None, y=None):
        e = self.engine
        if x is None:
            self._take_staged()
        else:
            e.x.copy_(x, non_blocking=True)
            e.y.copy_(y, non_blocking=True)
        self._run('resident', self._step_body)
        self._count_step()
        return e.loss, e.correct

    def _step_body(self):
        self.engine.replay()
        self._reduce_and_step(self.theta_all, self.tasks * self.world)

    def metrics(self):
        n, g = self.engine.P + self.engine.PH, self.tasks * self.world
        return self.red[n] / g, self.red[n + 1] / (g * self.engine.S)
