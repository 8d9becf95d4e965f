"""Task-batched MAML / ANIL meta-gradient engine: host-side launch programs over libxmeta's kernels.

One ``MamlEngine.run()`` replaces the body of the reference's per-task Python loop
(``vision/maml_vision.py:102-114``: ``maml.clone()`` -> ``fast_adapt`` -> ``eval_loss.backward()``) for a
whole shard of tasks at once; every task carries its own fast weights.  The second-order term is
computed forward-over-reverse: with theta_{t+1} = theta_t - lr * g(theta_t) the outer cotangent obeys
``bar_t = bar_{t+1} - lr * H(theta_t) bar_{t+1}`` and, the Hessian being symmetric, ``H v`` is the tangent of
the gradient computation in direction v.  Each inner step therefore costs one tangent ("dual")
forward+backward sweep that re-uses the activations saved by the primal sweep -- same algorithmic
FLOPs as the reference's reverse-over-reverse graph, a fraction of its memory
traffic and ~100x fewer launches.

The engine is a *static launch program*: all buffers are allocated once, every kernel argument
block is built once, and ``run()`` only replays the list -- so the whole iteration can be captured
into a CUDA graph (``capture()``).
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import (XM_CONV_DGRAD, XM_CONV_FWD, XM_STAT_NONE, XM_STAT_SUM_AUX, XM_STAT_SUM_SQ,
                   XmAnilHeadArgs, XmBlockGeom, XmBnArgs, XmConvArgs, XmHeadArgs, XmImgArgs, XmWgradArgs)

BN_EPS = 1e-5          # torch.nn.BatchNorm2d default, vision_models.py:168-174
BN_MOMENTUM = 0.1
# XM_NO_IMG_BLOCK=1 routes the first ConvBlock through the generic conv / bn / wgrad kernels (A/B comparison)
_NO_IMG_BLOCK = os.environ.get('XM_NO_IMG_BLOCK', '') not in ('', '0')
# XM_SIDE_STREAM=0 keeps every call on one stream (A/B comparison of the parallel weight-gradient branch)
_SIDE_STREAM = os.environ.get('XM_SIDE_STREAM', '1') not in ('', '0')


def _require_cuda(device):
    if device.type != 'cuda':
        raise _lib.XmetaError('exploring_meta_b200 runs on CUDA devices only (got %s); there is no CPU path' % device)


def _p(t, off=0):
    """Device address of element ``off`` of a float32 tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr() + off * t.element_size()


class _Program:
    """A recorded list of C-ABI calls: (function, argument block, name, lane).

    Lane 1 calls may run concurrently with the main lane between a ``fork()`` and the next ``join()``: the weight
    gradients of a block are off the critical path of the backward chain (bn_bwd -> dgrad -> bn_bwd -> ...), so at
    small task counts, where a kernel does not fill the GPU, they overlap with it on a second stream (captured into
    the same CUDA graph as a parallel branch).  ``replay(stream)`` without ``side`` runs everything in order on one
    stream."""
    FORK, JOIN = 'fork', 'join'

    def __init__(self, lib, conv_ws=None):
        self.lib = lib
        self.calls = []
        self.ops = []                    # calls interleaved with fork / join markers
        self.conv_ws = conv_ws          # workspace handed to every xm_conv call (stride-2 layers on the tcgen05 path)
        self._forked = False

    def emit(self, name, args, lane=0):
        if name == 'xm_conv' and self.conv_ws is not None:
            args.workspace, args.workspace_bytes = _p(self.conv_ws), self.conv_ws.numel() * 4
        call = (getattr(self.lib, name), args, name)
        self.calls.append(call)
        if lane == 1:
            self.ops.append((self.FORK,))       # side lane waits for everything issued on the main lane so far
            self._forked = True
        self.ops.append(call + (lane,))

    def emit_raw(self, name, *argv):
        call = (getattr(self.lib, name), argv, name)
        self.calls.append(call)
        self.ops.append(call + (0,))

    def join(self):
        """Main lane waits for the side lane (no-op if nothing was forked since the last join)."""
        if self._forked:
            self.ops.append((self.JOIN,))
            self._forked = False

    @staticmethod
    def _call(fn, args, name, stream):
        code = fn(*args, stream) if isinstance(args, tuple) else fn(ctypes.byref(args), stream)
        if code != 0:
            _lib.check(code, name)

    def replay(self, stream, main=None, side=None):
        """``stream``: raw handle of the main stream.  With ``main`` / ``side`` (torch.cuda.Stream objects, ``main``
        wrapping ``stream``) lane-1 calls go to ``side`` with event dependencies."""
        if side is None:
            for fn, args, name in self.calls:
                self._call(fn, args, name, stream)
            return
        for op in self.ops:
            if op[0] == self.FORK:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
            elif op[0] == self.JOIN:
                ev = torch.cuda.Event()
                ev.record(side)
                main.wait_event(ev)
            else:
                fn, args, name, lane = op
                self._call(fn, args, name, side.cuda_stream if lane == 1 else stream)


class _EngineBase:
    def __init__(self, spec, tasks, device):
        self.spec = spec
        self.tasks = int(tasks)
        self.device = torch.device(device)
        _require_cuda(self.device)
        self.lib = _lib.load()
        self.dims = spec.block_dims()
        self.C = spec.hidden
        self.offs, self.P = spec.param_offsets()
        self._graph = None

    # ---- allocation helpers -------------------------------------------------------------------
    def _f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    def _f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.device)

    def geom(self, l, n):
        cin, hin, win, hz, wz, hp, wp = self.dims[l]
        return XmBlockGeom(self.tasks, n, cin, self.C, hin, win, hz, wz, hp, wp,
                           1 if self.spec.pool else 2, 1 if self.spec.pool else 0)

    def zshape(self, l, n):
        _, _, _, hz, wz, _, _ = self.dims[l]
        return (self.tasks, n, hz, wz, self.C)

    def pshape(self, l, n):
        _, _, _, _, _, hp, wp = self.dims[l]
        return (self.tasks, n, hp, wp, self.C)

    def _alloc_common(self, nmax):
        L = self.spec.layers
        self.sums = self._f64(self.tasks, 2, self.C)
        self.dsums = self._f64(self.tasks, 2, self.C)
        g0 = self.geom(0, nmax)
        nbytes = max(int(self.lib.xm_bn_scratch_bytes(ctypes.byref(self.geom(l, nmax)))) for l in range(L))
        self.bn_scratch = torch.empty(max(nbytes, 8) // 8, dtype=torch.float64, device=self.device)
        wbytes = max(int(self.lib.xm_wgrad_scratch_bytes(ctypes.byref(self.geom(l, nmax)))) for l in range(L))
        self.wg_partial = torch.empty(max(wbytes, 4) // 4, dtype=torch.float32, device=self.device)
        cbytes = max(int(self.lib.xm_conv_workspace_bytes(ctypes.byref(self.geom(l, nmax)))) for l in range(L))
        self.conv_ws = torch.empty(max(cbytes, 4) // 4, dtype=torch.float32, device=self.device) if cbytes > 0 else None
        del g0
        # fused image block (xm_img_*): first ConvBlock without ever writing its pre-BN map (csrc/img_block.cu)
        g = self.geom(0, nmax)
        self.img = bool(self.lib.xm_img_supported(ctypes.byref(g))) and not _NO_IMG_BLOCK
        if self.img:
            K = 9 * g.cin
            self.img_scratch = self._f64(int(self.lib.xm_img_scratch_bytes(ctypes.byref(g))) // 8)
            self._gram_words = self.tasks * (K * K + K)

    def _img_state(self, n):
        """Side buffers of one image-block call: winner values / positions (pooled variant only: the stride-2
        variant recomputes everything from the image) and the sparse sums of the backward."""
        st = {'ssum': self._f64(self.tasks, self.C, 9 * self.spec.in_c + 3), 'zsel': None, 'sel': None}
        if self.spec.pool:
            st['zsel'] = self._f32(*self.pshape(0, n))
            st['sel'] = torch.empty(self.pshape(0, n), dtype=torch.uint8, device=self.device)
        return st

    def _img_args(self, n, img, gram, theta, tstride):
        o = self.offs
        a = XmImgArgs()
        a.g, a.eps = self.geom(0, n), BN_EPS
        a.row0, a.row_step, a.rows_per_task = img
        a.x, a.gram = _p(self.x), _p(gram)
        a.w, a.w_task_stride = _p(theta, o[2]), tstride
        a.gamma, a.beta, a.gb_task_stride = _p(theta, o[0]), _p(theta, o[1]), tstride
        a.scratch = _p(self.img_scratch)
        return a

    def _emit_gram(self, prog, n, img, gram, lane=0):
        a = XmImgArgs()
        a.g, a.eps = self.geom(0, n), BN_EPS
        a.row0, a.row_step, a.rows_per_task = img
        a.x, a.gram = _p(self.x), _p(gram)
        prog.emit('xm_img_gram', a, lane=lane)

    def _emit_img_fwd(self, prog, n, img, gram, theta, tstride, st, Pout, MI, call_stats):
        a = self._img_args(n, img, gram, theta, tstride)
        a.mean_invstd, a.call_stats, a.p = _p(MI), _p(call_stats), _p(Pout)
        a.zsel, a.sel = _p(st['zsel']), _p(st['sel'])
        prog.emit('xm_img_fwd', a)

    def _emit_img_bwd(self, prog, n, img, gram, theta, tstride, st, GP, MI, BR, out, out_stride, base, base_stride, scale):
        o = self.offs
        a = self._img_args(n, img, gram, theta, tstride)
        a.mean_invstd, a.bwd_red, a.gp = _p(MI), _p(BR), _p(GP)
        a.zsel, a.sel, a.ssum = _p(st['zsel']), _p(st['sel']), _p(st['ssum'])
        a.out_gamma, a.out_beta, a.out_w, a.out_b = _p(out, o[0]), _p(out, o[1]), _p(out, o[2]), _p(out, o[3])
        a.out_task_stride = out_stride
        a.base_gamma, a.base_beta = _p(base, o[0]), _p(base, o[1])
        a.base_w, a.base_b = _p(base, o[2]), _p(base, o[3])
        a.base_task_stride, a.scale = base_stride, scale
        prog.emit('xm_img_bwd', a)

    # ---- emitters shared by the MAML and ANIL programs --------------------------------------
    def _emit_block_fwd(self, prog, l, n, src, img, theta, tstride, Z, Pout, MI, call_stats):
        """conv -> batch statistics -> normalise + ReLU + pool   (ConvBlock.forward, vision_models.py:188-193)."""
        o = self.offs
        if l == 0 and self.img:
            gram, st = Z
            return self._emit_img_fwd(prog, n, img, gram, theta, tstride, st, Pout, MI, call_stats)
        a = XmConvArgs()
        a.g = self.geom(l, n)
        a.mode = XM_CONV_FWD
        a.stat_mode = XM_STAT_SUM_SQ
        if l == 0:
            a.src_nchw, a.row0, a.row_step, a.rows_per_task = 1, img[0], img[1], img[2]
            a.src1 = _p(self.x)
        else:
            a.src1 = _p(src)
        a.w1, a.w1_task_stride = _p(theta, o[4 * l + 2]), tstride
        a.out, a.stats = _p(Z), _p(self.sums)
        prog.emit('xm_conv', a)
        b = XmBnArgs()
        b.g, b.eps = self.geom(l, n), BN_EPS
        b.z, b.sums = _p(Z), _p(self.sums)
        b.gamma, b.beta, b.gb_task_stride = _p(theta, o[4 * l]), _p(theta, o[4 * l + 1]), tstride
        b.mean_invstd, b.call_stats, b.p = _p(MI), _p(call_stats), _p(Pout)
        b.scratch = _p(self.bn_scratch)
        prog.emit('xm_bn_fwd', b)

    def _emit_block_bwd(self, prog, l, n, xin, img, theta, tstride, Z, GP, MI, BR, GZ, GPprev,
                        out, out_stride, base, base_stride, scale):
        """BN/ReLU/pool backward -> dgrad -> wgrad, parameter gradients through the axpy epilogue."""
        o = self.offs
        if l == 0 and self.img:
            gram, st = Z
            return self._emit_img_bwd(prog, n, img, gram, theta, tstride, st, GP, MI, BR, out, out_stride,
                                      base, base_stride, scale)
        b = XmBnArgs()
        b.g, b.eps = self.geom(l, n), BN_EPS
        b.z, b.gp, b.mean_invstd, b.bwd_red, b.gz = _p(Z), _p(GP), _p(MI), _p(BR), _p(GZ)
        b.gamma, b.beta, b.gb_task_stride = _p(theta, o[4 * l]), _p(theta, o[4 * l + 1]), tstride
        b.out_gamma, b.out_beta, b.out_task_stride = _p(out, o[4 * l]), _p(out, o[4 * l + 1]), out_stride
        b.base_gamma, b.base_beta = _p(base, o[4 * l]), _p(base, o[4 * l + 1])
        b.base_task_stride, b.scale = base_stride, scale
        b.scratch = _p(self.bn_scratch)
        prog.emit('xm_bn_bwd', b)
        if l > 0:
            d = XmConvArgs()
            d.g, d.mode, d.stat_mode = self.geom(l, n), XM_CONV_DGRAD, XM_STAT_NONE
            d.src1, d.w1, d.w1_task_stride = _p(GZ), _p(theta, o[4 * l + 2]), tstride
            d.out = _p(GPprev)
            prog.emit('xm_conv', d)
        w = XmWgradArgs()
        w.g = self.geom(l, n)
        if l == 0:
            w.src_nchw, w.row0, w.row_step, w.rows_per_task = 1, img[0], img[1], img[2]
            w.x1 = _p(self.x)
        else:
            w.x1 = _p(xin)
        w.g1 = _p(GZ)
        w.out_w, w.out_b, w.out_task_stride = _p(out, o[4 * l + 2]), _p(out, o[4 * l + 3]), out_stride
        w.base_w, w.base_b, w.base_task_stride = _p(base, o[4 * l + 2]), _p(base, o[4 * l + 3]), base_stride
        w.scale = scale
        w.partial, w.partial_bytes = _p(self.wg_partial), self.wg_partial.numel() * 4
        prog.emit('xm_wgrad', w, lane=1)      # off the critical path: parallel branch (one scratch buffer: all on lane 1)

    # ---- execution ----------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == 'cuda' else 0

    def replay(self, parallel=True):
        """Issues the program on the current stream; with ``parallel`` the side-lane calls (weight gradients) go to a
        second stream that forks from / joins into it (inside a capture this becomes a parallel graph branch)."""
        if not (parallel and _SIDE_STREAM and self.device.type == 'cuda'):
            return self.prog.replay(self._stream())
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)
        self.prog.replay(main.cuda_stream, main, self._side)

    def launch(self):
        """Replays the program on the current stream (asynchronous)."""
        if self._graph is not None:
            self._graph.replay()
        else:
            self.replay()

    def capture(self):
        """Captures the launch program into a CUDA graph; subsequent ``launch()`` calls replay it."""
        if self._graph is not None:
            return
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.prog.replay(side.cuda_stream)          # warm-up outside capture
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.replay()
        self._graph = graph

    @property
    def launches_per_run(self):
        return len(self.prog.calls)

    def rebuild(self):
        """Re-records the launch program (after a caller re-pointed ``theta`` / ``grad`` at its own
        buffers) and drops any captured graph."""
        self.prog = _Program(self.lib, self.conv_ws)
        self._graph = None
        self._build()


class MamlEngine(_EngineBase):
    """Meta-gradient of ``tasks`` MAML tasks (second-order by default), all buffers static.

    Inputs are written by the caller into ``x`` [tasks, 2S, C, H, W] (support = even rows, query = odd
    rows, as ``prepare_batch`` splits them), ``y`` [tasks, 2S] int64 and ``theta`` [P] (master
    parameters, ``parameters()`` order).  After ``launch()``:
      ``grad``       [P]      sum over tasks of d(query loss)/d(theta)      (NOT yet scaled by 1/B)
      ``loss``       [tasks]  query loss after adaptation (``fast_adapt``'s valid_loss)
      ``correct``    [tasks]  int32 argmax==label count (``accuracy`` * Q)
      ``theta_steps``[steps, tasks, P]  adapted fast weights theta_1..theta_T
      ``call_stats`` [steps+1, layers, tasks, 2, C]  BN batch mean / unbiased var of every forward call
    ``mode``: 'second' (reference default), 'first' (first_order=True), 'eval' (adapt + query metrics
    only: the validation / meta-test pass of maml_vision.py:117-124 and core_functions/vision.py:26-42).
    """

    def __init__(self, spec, tasks, shots, steps, inner_lr, mode='second', device='cuda'):
        super().__init__(spec, tasks, device)
        assert mode in ('second', 'first', 'eval')
        assert spec.head in ('flatten', 'mean')
        self.shots, self.steps, self.lr, self.mode = int(shots), int(steps), float(inner_lr), mode
        self.S = spec.ways * self.shots
        self.rows = 2 * self.S
        B, T, L, C, P = self.tasks, self.steps, spec.layers, self.C, self.P
        S = self.S
        self.x = self._f32(B, self.rows, spec.in_c, spec.in_h, spec.in_w)
        self.y = torch.zeros((B, self.rows), dtype=torch.int64, device=self.device)
        self.theta = self._f32(P)
        self.grad = torch.zeros(P, dtype=torch.float32, device=self.device)
        self.loss = torch.zeros(B, dtype=torch.float32, device=self.device)
        self.correct = torch.zeros(B, dtype=torch.int32, device=self.device)
        self.theta_steps = self._f32(max(T, 1), B, P)
        self.call_stats = torch.zeros((T + 1, L, B, 2, C), dtype=torch.float32, device=self.device)
        self._alloc_common(S)
        keep = T if mode == 'second' else 1           # per-step activations are only re-read by the dual sweep
        l0 = 1 if self.img else 0                     # the fused image block keeps (gram, winners) instead of z
        if self.img:
            self.gram_sup = self._f64(self._gram_words)
            self.gram_qry = self._f64(self._gram_words)
        self.Z = [[(self.gram_sup, self._img_state(S)) if l < l0 else self._f32(*self.zshape(l, S))
                   for l in range(L)] for _ in range(keep)]
        self.Pa = [[self._f32(*self.pshape(l, S)) for l in range(L)] for _ in range(keep)]
        self.GP = [[self._f32(*self.pshape(l, S)) for l in range(L)] for _ in range(keep)]
        self.MI = [[self._f32(B, 2, C) for l in range(L)] for _ in range(keep)]
        self.BR = [[self._f32(B, 2, C) for l in range(L)] for _ in range(keep)]
        # temporaries: query pass (phase 2) and dual sweep (phase 3) share them
        self.tZ = [(self.gram_qry, self._img_state(S)) if l < l0 else self._f32(*self.zshape(l, S)) for l in range(L)]
        self.tZD = self._f32(*self.pshape(0, S)) if (self.img and spec.pool and mode == 'second') else None   # zdot at the winners
        self.tP = [self._f32(*self.pshape(l, S)) for l in range(L)]
        self.tGP = [self._f32(*self.pshape(l, S)) for l in range(L)]
        self.tMI = [self._f32(B, 2, C) for l in range(L)]
        self.tBR = [self._f32(B, 2, C) for l in range(L)]
        self.tDR = [self._f32(B, 2, C) for l in range(L)]
        # pre-BN cotangents, one buffer per layer: the weight gradient of layer l (parallel branch) reads GZ[l] while the
        # main chain is already producing GZ[l-1]
        self.GZ = [None if l < l0 else self._f32(self.tZ[l].numel()) for l in range(L)]
        self.GZdot = [None if (l < l0 or mode != 'second') else self._f32(self.tZ[l].numel()) for l in range(L)]
        self.bar = [self._f32(B, P), self._f32(B, P)] if mode != 'eval' else None
        self.prog = _Program(self.lib, self.conv_ws)
        self._build()

    # theta_t as (tensor, task stride): theta_0 is the shared master vector
    def _theta(self, t):
        if t == 0:
            return self.theta, 0
        return self.theta_steps[t - 1], self.P

    def _emit_head(self, prog, n, feat, theta, tstride, label_row0, GPout, out, out_stride, base,
                   base_stride, scale, loss=None, correct=None, dual=None):
        o = self.offs
        hp, wp = self.spec.out_hw()
        h = XmHeadArgs()
        h.tasks, h.n, h.ways, h.c, h.hw = self.tasks, n, self.spec.ways, self.C, hp * wp
        h.mode = 0 if self.spec.head == 'flatten' else 1
        h.feat = _p(feat)
        h.labels, h.label_row0, h.label_row_step, h.labels_per_task = _p(self.y), label_row0, 2, self.rows
        L4 = 4 * self.spec.layers
        h.w, h.b, h.wb_task_stride = _p(theta, o[L4]), _p(theta, o[L4 + 1]), tstride
        h.loss, h.correct = _p(loss), _p(correct)
        if dual is None:
            h.dual = 0
            h.g_feat = _p(GPout)
        else:
            featdot, v = dual
            h.dual = 1
            h.feat_dot = _p(featdot)
            h.w_dot, h.b_dot, h.wbdot_task_stride = _p(v, o[L4]), _p(v, o[L4 + 1]), self.P
            h.g_feat_dot = _p(GPout)
        if out is not None:
            h.out_w, h.out_b, h.out_task_stride = _p(out, o[L4]), _p(out, o[L4 + 1]), out_stride
            h.base_w, h.base_b = _p(base, o[L4]), _p(base, o[L4 + 1])
            h.base_task_stride, h.scale = base_stride, scale
        prog.emit('xm_head', h)

    def _build(self):
        prog, L, T, S, P = self.prog, self.spec.layers, self.steps, self.S, self.P
        sup, qry = (0, 2, self.rows), (1, 2, self.rows)
        if self.img:                                   # Gram matrices of the support / query images, once
            self._emit_gram(prog, S, sup, self.gram_sup)
            # the query rows' Gram matrix is first read by the query pass (phase 2): parallel branch, joined with the
            # first inner step's weight gradients
            self._emit_gram(prog, S, qry, self.gram_qry, lane=1 if T > 0 else 0)
        # ---- phase 1: T inner steps on the support rows (core_functions/vision.py:9-13) ----------
        for t in range(T):
            k = t if self.mode == 'second' else 0
            th, ts = self._theta(t)
            nxt = self.theta_steps[t]
            if t > 0:
                prog.join()                            # theta_t complete (weight gradients of step t-1 run on the side lane)
            for l in range(L):
                self._emit_block_fwd(prog, l, S, self.Pa[k][l - 1] if l else None, sup, th, ts,
                                     self.Z[k][l], self.Pa[k][l], self.MI[k][l], self.call_stats[t, l])
            self._emit_head(prog, S, self.Pa[k][L - 1], th, ts, 0, self.GP[k][L - 1],
                            nxt, P, th, ts, -self.lr)
            for l in reversed(range(L)):
                self._emit_block_bwd(prog, l, S, self.Pa[k][l - 1] if l else None, sup, th, ts,
                                     self.Z[k][l], self.GP[k][l], self.MI[k][l], self.BR[k][l], self.GZ[l],
                                     self.GP[k][l - 1] if l else None, nxt, P, th, ts, -self.lr)
        # ---- phase 2: query loss / accuracy at theta_T (vision.py:15-17) and its gradient -----
        th, ts = self._theta(T)
        prog.join()
        for l in range(L):
            self._emit_block_fwd(prog, l, S, self.tP[l - 1] if l else None, qry, th, ts,
                                 self.tZ[l], self.tP[l], self.tMI[l], self.call_stats[T, l])
        if self.mode == 'eval':
            self._emit_head(prog, S, self.tP[L - 1], th, ts, 1, None, None, 0, None, 0, 0.0,
                            loss=self.loss, correct=self.correct)
            return
        bar = self.bar[0]
        self._emit_head(prog, S, self.tP[L - 1], th, ts, 1, self.tGP[L - 1], bar, P, None, 0, 1.0,
                        loss=self.loss, correct=self.correct)
        for l in reversed(range(L)):
            self._emit_block_bwd(prog, l, S, self.tP[l - 1] if l else None, qry, th, ts,
                                 self.tZ[l], self.tGP[l], self.tMI[l], self.tBR[l], self.GZ[l],
                                 self.tGP[l - 1] if l else None, bar, P, None, 0, 1.0)
        # ---- phase 3: bar_t = bar_{t+1} - lr * H(theta_t) bar_{t+1}, t = T-1 .. 0 (second order) ----
        cur = 0
        if self.mode == 'second':
            for t in reversed(range(T)):
                self._emit_dual_step(prog, t, self.bar[cur], self.bar[1 - cur])
                cur = 1 - cur
        # ---- sum over tasks in task order = accumulation into the master .grad (maml_vision.py:112) --
        prog.join()
        prog.emit_raw('xm_accumulate_tasks', _p(self.bar[cur]), P, self.tasks, P, _p(self.grad), 0)

    def _emit_dual_step(self, prog, t, v, out):
        """Tangent sweep of inner step t in direction v (= cotangent of theta_{t+1}); writes
        out = v - lr * d/d(eps) grad(theta_t + eps v)."""
        L, S, P, o = self.spec.layers, self.S, self.P, self.offs
        sup = (0, 2, self.rows)
        th, ts = self._theta(t)
        Z, Pa, GP, MI, BR = self.Z[t], self.Pa[t], self.GP[t], self.MI[t], self.BR[t]
        prog.join()                                    # v complete; the previous sweep's readers of tP / GZ are done
        for l in range(L):
            if l == 0 and self.img:          # fused image block: zdot only at the winners, statistics from the Gram matrix
                gram, st = Z[0]
                a = self._img_args(S, sup, gram, th, ts)
                a.w_dot, a.wdot_task_stride = _p(v, o[2]), P
                a.gamma_dot, a.beta_dot, a.gbdot_task_stride = _p(v, o[0]), _p(v, o[1]), P
                a.mean_invstd, a.dual_red = _p(MI[0]), _p(self.tDR[0])
                a.zsel, a.sel, a.pdot, a.zdsel = _p(st['zsel']), _p(st['sel']), _p(self.tP[0]), _p(self.tZD)
                prog.emit('xm_img_dual_fwd', a)
                continue
            a = XmConvArgs()
            a.g, a.mode, a.stat_mode = self.geom(l, S), XM_CONV_FWD, XM_STAT_SUM_AUX
            if l == 0:                       # images carry no tangent: zdot = conv(x, Wdot)
                a.src_nchw, a.row0, a.row_step, a.rows_per_task = 1, sup[0], sup[1], sup[2]
                a.src1 = _p(self.x)
            else:
                a.src1 = _p(Pa[l - 1])
                a.src2, a.w2, a.w2_task_stride = _p(self.tP[l - 1]), _p(th, o[4 * l + 2]), ts
            a.w1, a.w1_task_stride = _p(v, o[4 * l + 2]), P
            a.out, a.aux, a.stats = _p(self.tZ[l]), _p(Z[l]), _p(self.dsums)
            prog.emit('xm_conv', a)
            b = XmBnArgs()
            b.g, b.eps = self.geom(l, S), BN_EPS
            b.z, b.zdot, b.dsums, b.mean_invstd = _p(Z[l]), _p(self.tZ[l]), _p(self.dsums), _p(MI[l])
            b.gamma, b.beta, b.gb_task_stride = _p(th, o[4 * l]), _p(th, o[4 * l + 1]), ts
            b.gamma_dot, b.beta_dot, b.gbdot_task_stride = _p(v, o[4 * l]), _p(v, o[4 * l + 1]), P
            b.pdot, b.dual_red = _p(self.tP[l]), _p(self.tDR[l])
            b.scratch = _p(self.bn_scratch)
            prog.emit('xm_bn_dual_fwd', b)
        self._emit_head(prog, S, Pa[L - 1], th, ts, 0, self.tGP[L - 1], out, P, v, P, -self.lr,
                        dual=(self.tP[L - 1], v))
        for l in reversed(range(L)):
            if l == 0 and self.img:
                gram, st = Z[0]
                a = self._img_args(S, sup, gram, th, ts)
                a.w_dot, a.wdot_task_stride = _p(v, o[2]), P
                a.gamma_dot, a.beta_dot, a.gbdot_task_stride = _p(v, o[0]), _p(v, o[1]), P
                a.mean_invstd, a.bwd_red, a.dual_red = _p(MI[0]), _p(BR[0]), _p(self.tDR[0])
                a.gp, a.gpdot = _p(GP[0]), _p(self.tGP[0])
                a.zsel, a.sel, a.zdsel, a.ssum = _p(st['zsel']), _p(st['sel']), _p(self.tZD), _p(st['ssum'])
                a.out_gamma, a.out_beta, a.out_w, a.out_b = _p(out, o[0]), _p(out, o[1]), _p(out, o[2]), _p(out, o[3])
                a.out_task_stride = P
                a.base_gamma, a.base_beta, a.base_w, a.base_b = _p(v, o[0]), _p(v, o[1]), _p(v, o[2]), _p(v, o[3])
                a.base_task_stride, a.scale = P, -self.lr
                prog.emit('xm_img_dual_bwd', a)
                continue
            b = XmBnArgs()
            b.g, b.eps = self.geom(l, S), BN_EPS
            b.z, b.zdot, b.gp, b.gpdot = _p(Z[l]), _p(self.tZ[l]), _p(GP[l]), _p(self.tGP[l])
            b.mean_invstd, b.bwd_red, b.dual_red = _p(MI[l]), _p(BR[l]), _p(self.tDR[l])
            b.gamma, b.beta, b.gb_task_stride = _p(th, o[4 * l]), _p(th, o[4 * l + 1]), ts
            b.gamma_dot, b.beta_dot, b.gbdot_task_stride = _p(v, o[4 * l]), _p(v, o[4 * l + 1]), P
            # gz is re-derived only where a consumer needs it: dgrad / wgrad pair 2 of layers with a tangent input
            b.gz, b.gzdot = (_p(self.GZ[l]) if l > 0 else None), _p(self.GZdot[l])
            b.out_gamma, b.out_beta, b.out_task_stride = _p(out, o[4 * l]), _p(out, o[4 * l + 1]), P
            b.base_gamma, b.base_beta, b.base_task_stride = _p(v, o[4 * l]), _p(v, o[4 * l + 1]), P
            b.scale = -self.lr
            b.scratch = _p(self.bn_scratch)
            prog.emit('xm_bn_dual_bwd', b)
            if l > 0:                        # gxdot = dgrad(gzdot, W) + dgrad(gz, Wdot)
                d = XmConvArgs()
                d.g, d.mode, d.stat_mode = self.geom(l, S), XM_CONV_DGRAD, XM_STAT_NONE
                d.src1, d.w1, d.w1_task_stride = _p(self.GZdot[l]), _p(th, o[4 * l + 2]), ts
                d.src2, d.w2, d.w2_task_stride = _p(self.GZ[l]), _p(v, o[4 * l + 2]), P
                d.out = _p(self.tGP[l - 1])
                prog.emit('xm_conv', d)
            w = XmWgradArgs()                # gWdot = wgrad(x, gzdot) + wgrad(xdot, gz)
            w.g = self.geom(l, S)
            if l == 0:
                w.src_nchw, w.row0, w.row_step, w.rows_per_task = 1, sup[0], sup[1], sup[2]
                w.x1 = _p(self.x)
            else:
                w.x1 = _p(Pa[l - 1])
                w.x2, w.g2 = _p(self.tP[l - 1]), _p(self.GZ[l])
            w.g1 = _p(self.GZdot[l])
            w.out_w, w.out_b, w.out_task_stride = _p(out, o[4 * l + 2]), _p(out, o[4 * l + 3]), P
            w.base_w, w.base_b, w.base_task_stride = _p(v, o[4 * l + 2]), _p(v, o[4 * l + 3]), P
            w.scale = -self.lr
            w.partial, w.partial_bytes = _p(self.wg_partial), self.wg_partial.numel() * 4
            prog.emit('xm_wgrad', w, lane=1)

    # ---- convenience ----------------------------------------------------------------------------
    def run(self, x=None, y=None, theta=None):
        """Copies the given inputs into the static buffers (any may be None = already in place) and
        launches.  Returns ``(grad, loss, correct)`` views of the static output buffers."""
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        if theta is not None:
            self.theta.copy_(theta, non_blocking=True)
        self.launch()
        return self.grad, self.loss, self.correct

    def update_running_stats(self, running_mean, running_var):
        """Applies the BN running-statistics side effect of this shard's forward calls to the given
        per-layer buffers, in the reference's order: task-major, then call (T support calls, 1 query)."""
        T, L, B, C = self.steps, self.spec.layers, self.tasks, self.C
        cs = self.call_stats
        for l in range(L):
            _lib.check(self.lib.xm_bn_ema(_p(running_mean[l]), _p(running_var[l]), _p(cs[0, l]),
                                          B, 2 * C, T + 1, cs.stride(0), C, BN_MOMENTUM, self._stream()),
                       'xm_bn_ema')
        return B * (T + 1)


class AnilEngine(_EngineBase):
    """ANIL (``vision/anil_vision.py:109-122``): body forward once per task over all 2S rows with the
    shared body weights (one BN batch per task, ``utils/data_pre.py:118-119``), head-only adaptation
    in one kernel, then a first-order body backward.  Outputs: ``grad`` [P_body] and ``head_grad``
    [ways*D + ways] (sums over tasks, unscaled), ``loss``, ``correct``, ``call_stats`` [layers, tasks, 2, C]."""

    def __init__(self, spec, tasks, shots, steps, inner_lr, first_order=False, device='cuda', mode='train'):
        super().__init__(spec, tasks, device)
        assert spec.head == 'none' and mode in ('train', 'eval')
        self.shots, self.steps, self.lr = int(shots), int(steps), float(inner_lr)
        self.first_order = bool(first_order)
        self.mode = mode              # 'eval': body forward + head adaptation + query metrics only (validation / test)
        self.S = spec.ways * self.shots
        self.rows = 2 * self.S
        B, L, C, P, R = self.tasks, spec.layers, self.C, self.P, self.rows
        hp, wp = spec.out_hw()
        self.D = C * hp * wp
        self.PH = spec.ways * self.D + spec.ways
        self.x = self._f32(B, R, spec.in_c, spec.in_h, spec.in_w)
        self.y = torch.zeros((B, R), dtype=torch.int64, device=self.device)
        self.theta = self._f32(P)
        self.head = self._f32(self.PH)
        self.grad = torch.zeros(P, dtype=torch.float32, device=self.device)
        self.head_grad = torch.zeros(self.PH, dtype=torch.float32, device=self.device)
        self.loss = torch.zeros(B, dtype=torch.float32, device=self.device)
        self.correct = torch.zeros(B, dtype=torch.int32, device=self.device)
        self.call_stats = torch.zeros((L, B, 2, C), dtype=torch.float32, device=self.device)
        self._alloc_common(R)
        l0 = 1 if self.img else 0
        if self.img:
            self.gram = self._f64(self._gram_words)
        self.Z = [(self.gram, self._img_state(R)) if l < l0 else self._f32(*self.zshape(l, R)) for l in range(L)]
        self.Pa = [self._f32(*self.pshape(l, R)) for l in range(L)]
        self.GP = [self._f32(*self.pshape(l, R)) for l in range(L)]
        self.MI = [self._f32(B, 2, C) for l in range(L)]
        self.BR = [self._f32(B, 2, C) for l in range(L)]
        self.GZ = [None if l < l0 else self._f32(self.Z[l].numel()) for l in range(L)]
        self.task_grad = self._f32(B, P)
        self.task_head_grad = self._f32(B, self.PH)
        self.prog = _Program(self.lib, self.conv_ws)
        self._build()

    def _build(self):
        prog, L, B, P, R = self.prog, self.spec.layers, self.tasks, self.P, self.rows
        rows = (0, 1, R)
        if self.img:
            self._emit_gram(prog, R, rows, self.gram)
        for l in range(L):
            self._emit_block_fwd(prog, l, R, self.Pa[l - 1] if l else None, rows, self.theta, 0,
                                 self.Z[l], self.Pa[l], self.MI[l], self.call_stats[l])
        hp, wp = self.spec.out_hw()
        h = XmAnilHeadArgs()
        h.tasks, h.rows, h.ways, h.c, h.hw, h.mode = B, R, self.spec.ways, self.C, hp * wp, 0
        h.steps, h.first_order, h.lr = self.steps, int(self.first_order), self.lr
        h.feat, h.labels = _p(self.Pa[L - 1]), _p(self.y)
        h.w, h.b = _p(self.head), _p(self.head, self.spec.ways * self.D)
        h.loss, h.correct, h.g_feat = _p(self.loss), _p(self.correct), _p(self.GP[L - 1])
        h.g_w, h.g_b = _p(self.task_head_grad), _p(self.task_head_grad, self.spec.ways * self.D)
        h.g_task_stride = self.PH
        nbytes = int(self.lib.xm_anil_head_scratch_bytes(ctypes.byref(h)))
        self.head_scratch = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=self.device)
        h.scratch, h.scratch_bytes = _p(self.head_scratch), self.head_scratch.numel() * 4
        prog.emit('xm_anil_head', h)
        if self.mode == 'eval':
            return
        for l in reversed(range(L)):
            self._emit_block_bwd(prog, l, R, self.Pa[l - 1] if l else None, rows, self.theta, 0,
                                 self.Z[l], self.GP[l], self.MI[l], self.BR[l], self.GZ[l],
                                 self.GP[l - 1] if l else None, self.task_grad, P, None, 0, 1.0)
        prog.join()
        prog.emit_raw('xm_accumulate_tasks', _p(self.task_grad), P, B, P, _p(self.grad), 0)
        prog.emit_raw('xm_accumulate_tasks', _p(self.task_head_grad), self.PH, B, self.PH, _p(self.head_grad), 0)

    def run(self, x=None, y=None, theta=None, head=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        if theta is not None:
            self.theta.copy_(theta, non_blocking=True)
        if head is not None:
            self.head.copy_(head, non_blocking=True)
        self.launch()
        return self.grad, self.head_grad, self.loss, self.correct

    def update_running_stats(self, running_mean, running_var):
        L, B, C = self.spec.layers, self.tasks, self.C
        for l in range(L):
            _lib.check(self.lib.xm_bn_ema(_p(running_mean[l]), _p(running_var[l]), _p(self.call_stats[l]),
                                          B, 2 * C, 1, 0, C, BN_MOMENTUM, self._stream()), 'xm_bn_ema')
        return B
