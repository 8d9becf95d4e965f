"""Eager (one call = one BN batch) autograd view of libxmeta's block kernels.

``conv_block(x, gamma, beta, weight, bias, ...)`` is ``ConvBlock.forward`` of the reference
(``core_functions/vision_models.py:188-193``: conv3x3 -> BatchNorm2d with per-call batch statistics -> ReLU ->
MaxPool2d(2,2) or nothing) as ONE differentiable torch op whose forward, backward and double-backward all
run on the same sm_100a kernels the task-batched engine uses (with ``tasks = 1``):

  forward          xm_conv(FWD, SUM_SQ) -> xm_bn_fwd
  backward         xm_bn_bwd -> xm_conv(DGRAD) -> xm_wgrad                 (``torch.autograd.grad`` / ``.backward()``)
  double-backward  xm_conv(FWD, two pairs, SUM_AUX) -> xm_bn_dual_fwd -> xm_bn_dual_bwd -> xm_conv(DGRAD, two
                   pairs) -> xm_wgrad(two pairs)                           (``create_graph=True`` then ``.backward()``)

The double-backward needs no reverse-over-reverse kernels: the block backward is G(xi; g) = J(xi)^T g with
xi = (x, gamma, beta, W), i.e. the gradient of the scalar phi(xi) = <g, f(xi)> for fixed g.  Its vector-Jacobian
product with cotangent u is (d^2 phi / d xi^2) u -- a symmetric Hessian -- which equals the TANGENT of the backward
computation in direction u, and the cotangent reaching g is J(xi) u, the tangent of the forward.  Both come out
of one "dual" sweep (SURVEY App. F gives the closed forms the kernels implement).  Third derivatives are not
provided.

This is what lets ``MAML.clone()/adapt()`` (exploring_meta_b200/maml.py) keep learn2learn's exact autograd
semantics -- ``torch.autograd.grad(loss, fast_weights, create_graph=True)`` -- on top of the CUDA kernels.
Activations are NHWC in memory and are handed to torch as NCHW-shaped channels_last views.
"""
import ctypes

import torch

from . import _lib, engine as _engine
from ._lib import (XM_CONV_DGRAD, XM_CONV_FWD, XM_STAT_NONE, XM_STAT_SUM_AUX, XM_STAT_SUM_SQ, XmBlockGeom,
                   XmBnArgs, XmConvArgs, XmWgradArgs)

BN_EPS = 1e-5


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream if device.type == 'cuda' else 0


def _call(name, args, device):
    lib = _lib.load()
    code = getattr(lib, name)(ctypes.byref(args), _stream(device))
    if code != 0:
        _lib.check(code, name)


def _geom(n, cin, cout, hin, win, stride, pool):
    hz, wz = (hin + 2 - 3) // stride + 1, (win + 2 - 3) // stride + 1
    hp, wp = (hz // 2, wz // 2) if pool else (hz, wz)
    return XmBlockGeom(1, n, cin, cout, hin, win, hz, wz, hp, wp, stride, 1 if pool else 0)


def _nhwc(t):
    """NCHW-shaped tensor -> contiguous [N, H, W, C] fp32 storage (no copy when already channels_last)."""
    return t.permute(0, 2, 3, 1).contiguous()


def _as_nchw(t_nhwc):
    return t_nhwc.permute(0, 3, 1, 2)


def _src(args, x, prefix_nchw=True):
    """Points a conv/wgrad argument block at the block input: user images stay NCHW (gathered by the kernel,
    like the engine's first layer), everything else is NHWC.  Returns the tensor that must stay alive."""
    if x.is_contiguous() and x.size(1) <= 4:
        args.src_nchw, args.row0, args.row_step, args.rows_per_task = 1, 0, 1, x.size(0)
        return x
    return _nhwc(x)


def _scratch(g, device):
    lib = _lib.load()
    nb = int(lib.xm_bn_scratch_bytes(ctypes.byref(g)))
    wb = int(lib.xm_wgrad_scratch_bytes(ctypes.byref(g)))
    return (torch.empty(max(nb, 8) // 8, dtype=torch.float64, device=device),
            torch.empty(max(wb, 4) // 4, dtype=torch.float32, device=device))


class _BlockForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, weight, bias, stride, pool):
        dev = x.device
        _engine._require_cuda(dev)
        n, cin, hin, win = x.shape
        cout = weight.size(0)
        g = _geom(n, cin, cout, hin, win, stride, pool)
        x_in, gamma_in, beta_in, weight_in = x, gamma, beta, weight        # graph-connected: saved for backward
        x, gamma, beta, weight = x.detach(), gamma.detach().contiguous(), beta.detach().contiguous(), \
            weight.detach().contiguous()
        z = torch.empty((n, g.hz, g.wz, cout), dtype=torch.float32, device=dev)
        p = torch.empty((n, g.hp, g.wp, cout), dtype=torch.float32, device=dev)
        sums = torch.empty((2, cout), dtype=torch.float64, device=dev)
        mi = torch.empty((2, cout), dtype=torch.float32, device=dev)
        stats = torch.empty((2, cout), dtype=torch.float32, device=dev)
        bn_scratch, _ = _scratch(g, dev)
        a = XmConvArgs()
        a.g, a.mode, a.stat_mode = g, XM_CONV_FWD, XM_STAT_SUM_SQ
        xs = _src(a, x)
        a.src1, a.w1, a.out, a.stats = _ptr(xs), _ptr(weight), _ptr(z), _ptr(sums)
        _call('xm_conv', a, dev)
        b = XmBnArgs()
        b.g, b.eps = g, BN_EPS
        b.z, b.sums, b.gamma, b.beta = _ptr(z), _ptr(sums), _ptr(gamma), _ptr(beta)
        b.mean_invstd, b.call_stats, b.p, b.scratch = _ptr(mi), _ptr(stats), _ptr(p), _ptr(bn_scratch)
        _call('xm_bn_fwd', b, dev)
        ctx.geom = (n, cin, cout, hin, win, stride, pool)
        ctx.save_for_backward(x_in, gamma_in, beta_in, weight_in, z, mi)
        ctx.mark_non_differentiable(stats)
        return _as_nchw(p), stats

    @staticmethod
    def backward(ctx, gp, _gstats):
        x, gamma, beta, weight, z, mi = ctx.saved_tensors
        need_x = ctx.needs_input_grad[0]
        gx, ggamma, gbeta, gw = _BlockBackward.apply(x, gamma, beta, weight, gp, z, mi, ctx.geom, need_x)
        gbias = torch.zeros_like(gamma) if ctx.needs_input_grad[4] else None      # cancelled by train-mode BN
        return (gx if need_x else None), ggamma, gbeta, gw, gbias, None, None


class _BlockBackward(torch.autograd.Function):
    """(x, gamma, beta, W, g_p) -> (g_x, g_gamma, g_beta, g_W); z / mean_invstd are the forward's saved data."""

    @staticmethod
    def forward(ctx, x, gamma, beta, weight, gp, z, mi, geom, need_x):
        dev = z.device
        n, cin, cout, hin, win, stride, pool = geom
        g = _geom(*geom)
        x_in, gamma_in, beta_in, weight_in, gp_in = x, gamma, beta, weight, gp
        gamma, beta, weight = gamma.detach().contiguous(), beta.detach().contiguous(), weight.detach().contiguous()
        probe = XmWgradArgs()
        xs = _src(probe, x.detach())
        x_is_image = bool(probe.src_nchw)
        gp = _nhwc(gp.detach().float())
        gz = torch.empty_like(z)
        bwd_red = torch.empty((2, cout), dtype=torch.float32, device=dev)
        ggamma = torch.empty(cout, dtype=torch.float32, device=dev)
        gbeta = torch.empty(cout, dtype=torch.float32, device=dev)
        gw = torch.empty_like(weight)
        bn_scratch, wg_partial = _scratch(g, dev)
        b = XmBnArgs()
        b.g, b.eps = g, BN_EPS
        b.z, b.gp, b.mean_invstd, b.bwd_red, b.gz = _ptr(z), _ptr(gp), _ptr(mi), _ptr(bwd_red), _ptr(gz)
        b.gamma, b.beta = _ptr(gamma), _ptr(beta)
        b.out_gamma, b.out_beta, b.scale, b.scratch = _ptr(ggamma), _ptr(gbeta), 1.0, _ptr(bn_scratch)
        _call('xm_bn_bwd', b, dev)
        gx = None
        if need_x:
            gx = torch.empty((n, hin, win, cin), dtype=torch.float32, device=dev)
            d = XmConvArgs()
            d.g, d.mode, d.stat_mode = g, XM_CONV_DGRAD, XM_STAT_NONE
            d.src1, d.w1, d.out = _ptr(gz), _ptr(weight), _ptr(gx)
            _call('xm_conv', d, dev)
        w = XmWgradArgs()
        w.g = g
        if x_is_image:
            w.src_nchw, w.row0, w.row_step, w.rows_per_task = 1, 0, 1, n
        w.x1, w.g1, w.out_w, w.scale = _ptr(xs), _ptr(gz), _ptr(gw), 1.0
        w.partial, w.partial_bytes = _ptr(wg_partial), wg_partial.numel() * 4
        _call('xm_wgrad', w, dev)
        ctx.geom, ctx.need_x = geom, need_x
        ctx.save_for_backward(x_in, gamma_in, beta_in, weight_in, gp_in, z, mi, bwd_red)
        if gx is None:
            gx = torch.zeros((), device=dev)
            ctx.mark_non_differentiable(gx)
        else:
            gx = _as_nchw(gx)
        return gx, ggamma, gbeta, gw

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, ux, ugamma, ubeta, uw):
        x, gamma, beta, weight, gp, z, mi, bwd_red = ctx.saved_tensors
        dev = z.device
        n, cin, cout, hin, win, stride, pool = ctx.geom
        g = _geom(*ctx.geom)
        gamma, beta, weight = gamma.contiguous(), beta.contiguous(), weight.contiguous()
        probe = XmWgradArgs()
        xs = _src(probe, x)
        x_is_image = bool(probe.src_nchw)
        gp = _nhwc(gp.float())
        have_ux = ctx.need_x and ux is not None
        if have_ux and x_is_image:            # src_nchw applies to both pairs of a call: a tangent input needs NHWC
            xs, x_is_image = _nhwc(x), False
        ugamma = torch.zeros_like(gamma) if ugamma is None else ugamma.contiguous().float()
        ubeta = torch.zeros_like(beta) if ubeta is None else ubeta.contiguous().float()
        uw = torch.zeros_like(weight) if uw is None else uw.contiguous().float()
        ux = _nhwc(ux.float()) if have_ux else None
        zdot = torch.empty_like(z)
        gz, gzdot = torch.empty_like(z), torch.empty_like(z)
        pdot = torch.empty_like(gp)
        dsums = torch.empty((2, cout), dtype=torch.float64, device=dev)
        dual_red = torch.empty((2, cout), dtype=torch.float32, device=dev)
        ggamma_dot = torch.empty(cout, dtype=torch.float32, device=dev)
        gbeta_dot = torch.empty(cout, dtype=torch.float32, device=dev)
        gw_dot = torch.empty_like(weight)
        bn_scratch, wg_partial = _scratch(g, dev)
        # zdot = conv(x, Wdot) + conv(xdot, W)
        a = XmConvArgs()
        a.g, a.mode, a.stat_mode = g, XM_CONV_FWD, XM_STAT_SUM_AUX
        if x_is_image:
            a.src_nchw, a.row0, a.row_step, a.rows_per_task = 1, 0, 1, n
        a.src1, a.w1 = _ptr(xs), _ptr(uw)
        if have_ux:
            a.src2, a.w2 = _ptr(ux), _ptr(weight)
        a.out, a.aux, a.stats = _ptr(zdot), _ptr(z), _ptr(dsums)
        _call('xm_conv', a, dev)
        b = XmBnArgs()
        b.g, b.eps = g, BN_EPS
        b.z, b.zdot, b.dsums, b.mean_invstd = _ptr(z), _ptr(zdot), _ptr(dsums), _ptr(mi)
        b.gamma, b.beta, b.gamma_dot, b.beta_dot = _ptr(gamma), _ptr(beta), _ptr(ugamma), _ptr(ubeta)
        b.pdot, b.dual_red, b.scratch = _ptr(pdot), _ptr(dual_red), _ptr(bn_scratch)
        _call('xm_bn_dual_fwd', b, dev)
        b2 = XmBnArgs()
        b2.g, b2.eps = g, BN_EPS
        b2.z, b2.zdot, b2.gp = _ptr(z), _ptr(zdot), _ptr(gp)            # gpdot = NULL: g_p is held fixed
        b2.mean_invstd, b2.bwd_red, b2.dual_red = _ptr(mi), _ptr(bwd_red), _ptr(dual_red)
        b2.gamma, b2.beta, b2.gamma_dot, b2.beta_dot = _ptr(gamma), _ptr(beta), _ptr(ugamma), _ptr(ubeta)
        b2.gz, b2.gzdot = _ptr(gz), _ptr(gzdot)
        b2.out_gamma, b2.out_beta, b2.scale, b2.scratch = _ptr(ggamma_dot), _ptr(gbeta_dot), 1.0, _ptr(bn_scratch)
        _call('xm_bn_dual_bwd', b2, dev)
        gx_dot = None
        if ctx.need_x:
            gx_dot = torch.empty((n, hin, win, cin), dtype=torch.float32, device=dev)
            d = XmConvArgs()
            d.g, d.mode, d.stat_mode = g, XM_CONV_DGRAD, XM_STAT_NONE
            d.src1, d.w1 = _ptr(gzdot), _ptr(weight)
            d.src2, d.w2 = _ptr(gz), _ptr(uw)
            d.out = _ptr(gx_dot)
            _call('xm_conv', d, dev)
            gx_dot = _as_nchw(gx_dot)
        w = XmWgradArgs()
        w.g = g
        if x_is_image:
            w.src_nchw, w.row0, w.row_step, w.rows_per_task = 1, 0, 1, n
        w.x1, w.g1 = _ptr(xs), _ptr(gzdot)
        if have_ux:
            w.x2, w.g2 = _ptr(ux), _ptr(gz)
        w.out_w, w.scale = _ptr(gw_dot), 1.0
        w.partial, w.partial_bytes = _ptr(wg_partial), wg_partial.numel() * 4
        _call('xm_wgrad', w, dev)
        if x_is_image and ctx.need_x:
            gx_dot = gx_dot.contiguous()
        return gx_dot, ggamma_dot, gbeta_dot, gw_dot, _as_nchw(pdot), None, None, None, None


def conv_block(x, gamma, beta, weight, bias=None, stride=1, pool=True):
    """``ConvBlock.forward`` on one BN batch.  x: [N, Cin, H, W]; returns ``(out [N, Cout, Hp, Wp], stats)``
    where ``stats[0]`` / ``stats[1]`` are the batch mean / UNBIASED variance of this call (what the running
    statistics are updated with).  ``bias`` takes part in autograd only to receive its (zero) gradient."""
    if bias is None:
        bias = torch.zeros_like(gamma)
    if x.dtype != torch.float32 or weight.dtype != torch.float32:
        raise _lib.XmetaError('exploring_meta_b200 kernels are fp32: got %s / %s' % (x.dtype, weight.dtype))
    return _BlockForward.apply(x, gamma, beta, weight, bias, int(stride), bool(pool))


def update_running_stats_(running_mean, running_var, stats, momentum=0.1):
    """BatchNorm2d's train-mode side effect for one call: r <- (1-m) r + m s (unbiased variance)."""
    dev = running_mean.device
    lib = _lib.load()
    c = running_mean.numel()
    _lib.check(lib.xm_bn_ema(_ptr(running_mean), _ptr(running_var), _ptr(stats), 1, 0, 1, 0, c,
                             float(momentum), _stream(dev)), 'xm_bn_ema')
