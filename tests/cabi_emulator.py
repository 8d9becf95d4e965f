"""CPU emulator of the libxmeta C ABI (TEST INFRASTRUCTURE -- never imported by the product).

Every function takes the same ctypes argument blocks as the real library, interprets the raw
addresses as CPU tensors and evaluates the kernel's *contract* with plain torch ops.  Uses:
  * ``-m "not gpu"`` tests run the product's host-side launch programs (exploring_meta_b200/engine.py)
    against this emulator and compare with the oracle -- this checks the forward-over-reverse
    algorithm, the buffer plumbing and the launch order without a GPU;
  * ``-m gpu`` kernel tests use the same functions as the per-kernel reference for the CUDA kernels.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

_NP = {torch.float32: np.float32, torch.float64: np.float64, torch.int64: np.int64, torch.int32: np.int32,
       torch.uint8: np.uint8}


def view(addr, shape, dtype=torch.float32):
    """A writable CPU tensor over raw memory at ``addr``."""
    n = 1
    for s in shape:
        n *= int(s)
    if n == 0:
        return torch.empty(shape, dtype=dtype)
    itemsize = torch.empty((), dtype=dtype).element_size()
    buf = (ctypes.c_char * (n * itemsize)).from_address(int(addr))
    arr = np.frombuffer(buf, dtype=_NP[dtype], count=n)
    return torch.from_numpy(arr).view(*shape)


def param(addr, stride, tasks, shape):
    """Per-task parameter tensor [tasks, *shape] gathered from ``addr + t*stride`` floats (a copy)."""
    n = 1
    for s in shape:
        n *= int(s)
    return torch.stack([view(addr + 4 * t * stride, (n,)).clone().view(*shape) for t in range(tasks)])


def write_param(addr, stride, tasks, value):
    for t in range(tasks):
        view(addr + 4 * t * stride, (value[t].numel(),)).copy_(value[t].reshape(-1))


def axpy_out(out, out_stride, base, base_stride, scale, tasks, grad):
    """out = (base ? base : 0) + scale * grad, per task."""
    if not out:
        return
    shape = grad.shape[1:]
    b = param(base, base_stride, tasks, shape) if base else torch.zeros_like(grad)
    write_param(out, out_stride, tasks, b + scale * grad)


def _nhwc_to_nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def _nchw_to_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _block_input(addr, g, nchw, row0, row_step, rows_per_task):
    """Block input of every task as NCHW [tasks, n, cin, hin, win]."""
    if nchw:
        full = view(addr, (g.tasks, rows_per_task, g.cin, g.hin, g.win))
        idx = torch.arange(g.n) * row_step + row0
        return full[:, idx].clone()
    x = view(addr, (g.tasks, g.n, g.hin, g.win, g.cin))
    return x.permute(0, 1, 4, 2, 3).contiguous()


def _sel_masks(a, z, mi, gamma, beta):
    """xhat, y and the 0/1 selection mask (pool arg-max AND y > 0) for one BN call, NHWC per task."""
    g = a.g
    mean, invstd = mi[:, 0], mi[:, 1]
    xhat = (z - mean[:, None, None, None, :]) * invstd[:, None, None, None, :]
    y = gamma[:, None, None, None, :] * xhat + beta[:, None, None, None, :]
    if not g.pool:
        return xhat, y, (y > 0).to(z.dtype)
    sel = torch.zeros_like(z)
    for t in range(g.tasks):
        act = F.relu(_nhwc_to_nchw(y[t]))
        pooled, idx = F.max_pool2d(act, 2, 2, return_indices=True)
        onehot = torch.zeros(act.shape[0], act.shape[1], g.hz * g.wz, dtype=z.dtype)
        onehot.scatter_(2, idx.flatten(2), (pooled.flatten(2) > 0).to(z.dtype))
        sel[t] = _nchw_to_nhwc(onehot.view(act.shape))
    return xhat, y, sel


def _up(gp, g):
    """Cotangent of the pooled output scattered back to every position of its window (NHWC)."""
    if not g.pool:
        return gp
    up = gp.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    return F.pad(up, (0, 0, 0, g.wz - 2 * g.wp, 0, g.hz - 2 * g.hp))


def _down(v, g):
    """Sum over each pooling window (exactly one selected element contributes)."""
    if not g.pool:
        return v
    v = v[:, :, :2 * g.hp, :2 * g.wp]
    return v.reshape(g.tasks, g.n, g.hp, 2, g.wp, 2, g.cout).sum(dim=(3, 5))


def _mean(v):
    return v.double().mean(dim=(1, 2, 3)).to(v.dtype)          # per (task, channel)


def _bc(s):
    return s[:, None, None, None, :]


class EmulatedLib:
    """Drop-in for the ctypes library object returned by ``exploring_meta_b200._lib.load()``."""

    def __init__(self):
        self.launches = 0
        self._err = b''

    @staticmethod
    def _args(ref):
        return ref._obj if hasattr(ref, '_obj') else ref

    # ------------------------------------------------------------------ misc
    def xm_version(self):
        return 100

    def xm_last_error(self):
        return self._err

    def xm_launch_count(self):
        return self.launches

    def xm_bn_scratch_bytes(self, ref):
        g = self._args(ref)
        return g.tasks * 4 * g.cout * 8

    def xm_conv_workspace_bytes(self, ref):
        return 0

    def xm_wgrad_scratch_bytes(self, ref):
        g = self._args(ref)
        return g.tasks * 9 * g.cin * g.cout * 4

    # ------------------------------------------------------------------ conv
    def xm_conv(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 1
        out = None
        for src, w, ws in ((a.src1, a.w1, a.w1_task_stride), (a.src2, a.w2, a.w2_task_stride)):
            if not src:
                continue
            W = param(w, ws, g.tasks, (g.cout, g.cin, 3, 3))
            if a.mode == 0:
                x = _block_input(src, g, a.src_nchw, a.row0, a.row_step, a.rows_per_task)
                r = torch.stack([F.conv2d(x[t], W[t], None, stride=g.stride, padding=1) for t in range(g.tasks)])
            else:
                gz = view(src, (g.tasks, g.n, g.hz, g.wz, g.cout)).permute(0, 1, 4, 2, 3).contiguous()
                r = torch.stack([torch.nn.grad.conv2d_input((g.n, g.cin, g.hin, g.win), W[t], gz[t],
                                                            stride=g.stride, padding=1) for t in range(g.tasks)])
            out = r if out is None else out + r
        out = out.permute(0, 1, 3, 4, 2).contiguous()          # NHWC
        view(a.out, out.shape).copy_(out)
        if a.stat_mode:
            st = view(a.stats, (g.tasks, 2, out.shape[-1]), torch.float64)
            st[:, 0] = out.double().sum(dim=(1, 2, 3))
            other = out if a.stat_mode == 1 else view(a.aux, out.shape)
            st[:, 1] = (out.double() * other.double()).sum(dim=(1, 2, 3))
        return 0

    def xm_wgrad(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 2
        total = None
        for xa, ga in ((a.x1, a.g1), (a.x2, a.g2)):
            if not xa:
                continue
            x = _block_input(xa, g, a.src_nchw, a.row0, a.row_step, a.rows_per_task)
            gz = view(ga, (g.tasks, g.n, g.hz, g.wz, g.cout)).permute(0, 1, 4, 2, 3).contiguous()
            r = torch.stack([torch.nn.grad.conv2d_weight(x[t], (g.cout, g.cin, 3, 3), gz[t],
                                                         stride=g.stride, padding=1) for t in range(g.tasks)])
            total = r if total is None else total + r
        axpy_out(a.out_w, a.out_task_stride, a.base_w, a.base_task_stride, a.scale, g.tasks, total)
        axpy_out(a.out_b, a.out_task_stride, a.base_b, a.base_task_stride, 0.0, g.tasks,
                 torch.zeros(g.tasks, g.cout))
        return 0

    # ------------------------------------------------------------------ BN + ReLU + pool
    def _gb(self, a):
        g = a.g
        return (param(a.gamma, a.gb_task_stride, g.tasks, (g.cout,)),
                param(a.beta, a.gb_task_stride, g.tasks, (g.cout,)))

    def xm_bn_fwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 1
        z = view(a.z, (g.tasks, g.n, g.hz, g.wz, g.cout))
        sums = view(a.sums, (g.tasks, 2, g.cout), torch.float64)
        cnt = g.n * g.hz * g.wz
        mean = sums[:, 0] / cnt
        var = (sums[:, 1] / cnt - mean * mean).clamp_min(0)
        mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
        mi[:, 0] = mean.float()
        mi[:, 1] = (1.0 / torch.sqrt(var + a.eps)).float()
        if a.call_stats:
            cs = view(a.call_stats, (g.tasks, 2, g.cout))
            cs[:, 0] = mean.float()
            cs[:, 1] = (var * (cnt / max(cnt - 1, 1))).float()
        gamma, beta = self._gb(a)
        xhat, y, sel = _sel_masks(a, z, mi, gamma, beta)
        act = F.relu(y)
        if g.pool:
            p = act[:, :, :2 * g.hp, :2 * g.wp].reshape(g.tasks, g.n, g.hp, 2, g.wp, 2, g.cout).amax(dim=(3, 5))
        else:
            p = act
        view(a.p, p.shape).copy_(p)
        return 0

    def xm_bn_bwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 2
        z = view(a.z, (g.tasks, g.n, g.hz, g.wz, g.cout))
        mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
        gamma, beta = self._gb(a)
        xhat, y, sel = _sel_masks(a, z, mi, gamma, beta)
        gbn = sel * _up(view(a.gp, (g.tasks, g.n, g.hp, g.wp, g.cout)), g)
        m1, m2 = _mean(gbn), _mean(gbn * xhat)
        br = view(a.bwd_red, (g.tasks, 2, g.cout))
        br[:, 0], br[:, 1] = m1, m2
        gz = _bc(gamma * mi[:, 1]) * (gbn - _bc(m1) - xhat * _bc(m2))
        view(a.gz, gz.shape).copy_(gz)
        cnt = g.n * g.hz * g.wz
        axpy_out(a.out_gamma, a.out_task_stride, a.base_gamma, a.base_task_stride, a.scale, g.tasks, m2 * cnt)
        axpy_out(a.out_beta, a.out_task_stride, a.base_beta, a.base_task_stride, a.scale, g.tasks, m1 * cnt)
        return 0

    def xm_bn_dual_fwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 1
        z = view(a.z, (g.tasks, g.n, g.hz, g.wz, g.cout))
        zd = view(a.zdot, z.shape)
        mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
        ds = view(a.dsums, (g.tasks, 2, g.cout), torch.float64)
        cnt = g.n * g.hz * g.wz
        mean, r = mi[:, 0].double(), mi[:, 1].double()
        d1 = ds[:, 0] / cnt
        d2 = r * (ds[:, 1] / cnt - mean * d1)
        dr = view(a.dual_red, (g.tasks, 2, g.cout))
        dr[:, 0], dr[:, 1] = d1.float(), d2.float()
        gamma, beta = self._gb(a)
        gd = param(a.gamma_dot, a.gbdot_task_stride, g.tasks, (g.cout,))
        bd = param(a.beta_dot, a.gbdot_task_stride, g.tasks, (g.cout,))
        xhat, y, sel = _sel_masks(a, z, mi, gamma, beta)
        xhd = _bc(mi[:, 1]) * (zd - _bc(dr[:, 0]) - xhat * _bc(dr[:, 1]))
        yd = _bc(gd) * xhat + _bc(gamma) * xhd + _bc(bd)
        pd = _down(sel * yd, g)
        view(a.pdot, pd.shape).copy_(pd)
        return 0

    def xm_bn_dual_bwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 2
        z = view(a.z, (g.tasks, g.n, g.hz, g.wz, g.cout))
        zd = view(a.zdot, z.shape)
        mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
        br = view(a.bwd_red, (g.tasks, 2, g.cout))
        dr = view(a.dual_red, (g.tasks, 2, g.cout))
        gamma, beta = self._gb(a)
        gd = param(a.gamma_dot, a.gbdot_task_stride, g.tasks, (g.cout,))
        xhat, y, sel = _sel_masks(a, z, mi, gamma, beta)
        gbn = sel * _up(view(a.gp, (g.tasks, g.n, g.hp, g.wp, g.cout)), g)
        if a.gpdot:
            gbnd = sel * _up(view(a.gpdot, (g.tasks, g.n, g.hp, g.wp, g.cout)), g)
        else:
            gbnd = torch.zeros_like(gbn)
        r, m1, m2, d1, d2 = mi[:, 1], br[:, 0], br[:, 1], dr[:, 0], dr[:, 1]
        e1, e2, e3 = _mean(gbnd), _mean(gbnd * xhat), _mean(gbn * zd)
        q = r * (e3 - d1 * m1 - d2 * m2)                 # <g * xhat_dot>
        rdot = -r * r * d2
        xhd = _bc(r) * (zd - _bc(d1) - xhat * _bc(d2))
        proj = gbn - _bc(m1) - xhat * _bc(m2)
        gz = _bc(gamma * r) * proj
        gzd = _bc(gd * r + gamma * rdot) * proj + _bc(gamma * r) * (gbnd - _bc(e1) - xhd * _bc(m2) - xhat * _bc(e2 + q))
        if a.gz:
            view(a.gz, gz.shape).copy_(gz)
        view(a.gzdot, gzd.shape).copy_(gzd)
        cnt = g.n * g.hz * g.wz
        axpy_out(a.out_gamma, a.out_task_stride, a.base_gamma, a.base_task_stride, a.scale, g.tasks, (e2 + q) * cnt)
        axpy_out(a.out_beta, a.out_task_stride, a.base_beta, a.base_task_stride, a.scale, g.tasks, e1 * cnt)
        return 0

    # ------------------------------------------------------------------ image block (fused first ConvBlock)
    # Contract evaluated the DENSE way (conv -> z -> BN -> pool and back); the CUDA kernels reach the same
    # numbers through the Gram-matrix closed forms of exploring_meta_b200/csrc/img_block.cu.
    @staticmethod
    def _img_ok(g):
        pooled = (1 <= g.cin <= 4 and g.stride == 1 and g.pool == 1 and g.hz % 2 == 0 and g.wz % 2 == 0
                  and g.cout % 32 == 0)
        flat = g.cin == 1 and g.stride == 2 and g.pool == 0 and g.cout in (32, 64)      # Omniglot image block
        return pooled or flat

    def xm_img_supported(self, ref):
        return 1 if self._img_ok(self._args(ref)) else 0

    def xm_img_gram_bytes(self, ref):
        g = self._args(ref)
        K = 9 * g.cin
        return g.tasks * (K * K + K) * 8

    def xm_img_scratch_bytes(self, ref):
        g = self._args(ref)
        return g.tasks * g.cout * (9 * g.cin + 3) * 8

    def _img_x(self, a):
        return _block_input(a.x, a.g, 1, a.row0, a.row_step, a.rows_per_task)       # [tasks, n, cin, H, W]

    def _img_conv(self, a, waddr, wstride):
        g = a.g
        x = self._img_x(a)
        W = param(waddr, wstride, g.tasks, (g.cout, g.cin, 3, 3))
        z = torch.stack([F.conv2d(x[t], W[t], None, stride=g.stride, padding=1) for t in range(g.tasks)])
        return z.permute(0, 1, 3, 4, 2).contiguous()                                # NHWC

    def _img_selmask(self, a, z=None):
        """0/1 mask over z positions decoded from the stored winner indices (no-pool variant: recomputed y > 0)."""
        g = a.g
        if not g.pool:
            mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
            gamma = param(a.gamma, a.gb_task_stride, g.tasks, (g.cout,))
            beta = param(a.beta, a.gb_task_stride, g.tasks, (g.cout,))
            y = _bc(gamma) * ((z - _bc(mi[:, 0])) * _bc(mi[:, 1])) + _bc(beta)
            return (y > 0).to(z.dtype)
        sel = view(a.sel, (g.tasks, g.n, g.hp, g.wp, g.cout), torch.uint8).long()
        m = torch.zeros(g.tasks, g.n, g.hp, 2, g.wp, 2, g.cout)
        for d in range(4):
            m[:, :, :, d >> 1, :, d & 1, :] = (sel == d).float()
        return m.reshape(g.tasks, g.n, g.hz, g.wz, g.cout)

    def xm_img_gram(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 1
        K = 9 * g.cin
        x = self._img_x(a).double()
        out = view(a.gram, (g.tasks, K * K + K), torch.float64)
        for t in range(g.tasks):
            cols = F.unfold(x[t], 3, padding=1, stride=g.stride)  # [n, K, positions], k = ci*9 + kh*3 + kw
            X = cols.permute(1, 0, 2).reshape(K, -1)
            out[t, :K * K] = (X @ X.t()).reshape(-1)
            out[t, K * K:] = X.sum(1)
        return 0

    def xm_img_fwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 1
        z = self._img_conv(a, a.w, a.w_task_stride)
        cnt = g.n * g.hz * g.wz
        mean = z.double().mean(dim=(1, 2, 3))
        var = (z.double() ** 2).mean(dim=(1, 2, 3)) - mean * mean
        var = var.clamp_min(0)
        mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
        mi[:, 0] = mean.float()
        mi[:, 1] = (1.0 / torch.sqrt(var + a.eps)).float()
        if a.call_stats:
            cs = view(a.call_stats, (g.tasks, 2, g.cout))
            cs[:, 0] = mean.float()
            cs[:, 1] = (var * (cnt / max(cnt - 1, 1))).float()
        gamma = param(a.gamma, a.gb_task_stride, g.tasks, (g.cout,))
        beta = param(a.beta, a.gb_task_stride, g.tasks, (g.cout,))
        xhat = (z - _bc(mi[:, 0])) * _bc(mi[:, 1])
        y = _bc(gamma) * xhat + _bc(beta)
        if not g.pool:
            view(a.p, y.shape).copy_(F.relu(y))
            return 0
        yw = y.reshape(g.tasks, g.n, g.hp, 2, g.wp, 2, g.cout).permute(0, 1, 2, 4, 6, 3, 5).reshape(
            g.tasks, g.n, g.hp, g.wp, g.cout, 4)
        zw = z.reshape(g.tasks, g.n, g.hp, 2, g.wp, 2, g.cout).permute(0, 1, 2, 4, 6, 3, 5).reshape(
            g.tasks, g.n, g.hp, g.wp, g.cout, 4)
        ymax, idx = yw.max(dim=-1)                               # first maximum in row-major window order
        # torch.max returns the first index among equal maxima on CPU
        on = ymax > 0
        view(a.p, ymax.shape).copy_(torch.where(on, ymax, torch.zeros_like(ymax)))
        view(a.zsel, ymax.shape).copy_(torch.where(on, zw.gather(-1, idx[..., None])[..., 0], torch.zeros_like(ymax)))
        view(a.sel, ymax.shape, torch.uint8).copy_(torch.where(on, idx, torch.full_like(idx, 255)).to(torch.uint8))
        return 0

    def _img_bwd_common(self, a):
        g = a.g
        z = self._img_conv(a, a.w, a.w_task_stride)
        mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
        gamma = param(a.gamma, a.gb_task_stride, g.tasks, (g.cout,))
        xhat = (z - _bc(mi[:, 0])) * _bc(mi[:, 1])
        sel = self._img_selmask(a, z)
        gbn = sel * _up(view(a.gp, (g.tasks, g.n, g.hp, g.wp, g.cout)), g)
        return z, mi, gamma, xhat, sel, gbn

    def _img_wgrad(self, a, gz):
        g = a.g
        x = self._img_x(a)
        gzn = gz.permute(0, 1, 4, 2, 3).contiguous()
        return torch.stack([torch.nn.grad.conv2d_weight(x[t], (g.cout, g.cin, 3, 3), gzn[t], stride=g.stride, padding=1)
                            for t in range(g.tasks)])

    def _img_outputs(self, a, gW, ggamma, gbeta):
        g = a.g
        axpy_out(a.out_w, a.out_task_stride, a.base_w, a.base_task_stride, a.scale, g.tasks, gW)
        axpy_out(a.out_b, a.out_task_stride, a.base_b, a.base_task_stride, 0.0, g.tasks, torch.zeros(g.tasks, g.cout))
        axpy_out(a.out_gamma, a.out_task_stride, a.base_gamma, a.base_task_stride, a.scale, g.tasks, ggamma)
        axpy_out(a.out_beta, a.out_task_stride, a.base_beta, a.base_task_stride, a.scale, g.tasks, gbeta)

    def xm_img_bwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 2
        z, mi, gamma, xhat, sel, gbn = self._img_bwd_common(a)
        m1, m2 = _mean(gbn), _mean(gbn * xhat)
        if a.bwd_red:
            br = view(a.bwd_red, (g.tasks, 2, g.cout))
            br[:, 0], br[:, 1] = m1, m2
        cnt = g.n * g.hz * g.wz
        if a.ssum:
            K = 9 * g.cin
            ss = view(a.ssum, (g.tasks, g.cout, K + 3), torch.float64)
            ss[:, :, :K] = self._img_wgrad(a, gbn).double().reshape(g.tasks, g.cout, K)
            ss[:, :, K] = gbn.double().sum(dim=(1, 2, 3))
            ss[:, :, K + 1] = (gbn.double() * xhat.double()).sum(dim=(1, 2, 3))
            ss[:, :, K + 2] = 0
        gz = _bc(gamma * mi[:, 1]) * (gbn - _bc(m1) - xhat * _bc(m2))
        self._img_outputs(a, self._img_wgrad(a, gz), m2 * cnt, m1 * cnt)
        return 0

    def xm_img_dual_fwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 1
        z = self._img_conv(a, a.w, a.w_task_stride)
        zd = self._img_conv(a, a.w_dot, a.wdot_task_stride)
        mi = view(a.mean_invstd, (g.tasks, 2, g.cout))
        mean, r = mi[:, 0].double(), mi[:, 1].double()
        d1 = zd.double().mean(dim=(1, 2, 3))
        d2 = r * ((zd.double() * z.double()).mean(dim=(1, 2, 3)) - mean * d1)
        dr = view(a.dual_red, (g.tasks, 2, g.cout))
        dr[:, 0], dr[:, 1] = d1.float(), d2.float()
        gamma = param(a.gamma, a.gb_task_stride, g.tasks, (g.cout,))
        gd = param(a.gamma_dot, a.gbdot_task_stride, g.tasks, (g.cout,))
        bd = param(a.beta_dot, a.gbdot_task_stride, g.tasks, (g.cout,))
        xhat = (z - _bc(mi[:, 0])) * _bc(mi[:, 1])
        xhd = _bc(mi[:, 1]) * (zd - _bc(dr[:, 0]) - xhat * _bc(dr[:, 1]))
        yd = _bc(gd) * xhat + _bc(gamma) * xhd + _bc(bd)
        sel = self._img_selmask(a, z)
        pd = _down(sel * yd, g)
        view(a.pdot, pd.shape).copy_(pd)
        if g.pool:
            view(a.zdsel, pd.shape).copy_(_down(sel * zd, g))
        return 0

    def xm_img_dual_bwd(self, ref, stream):
        a = self._args(ref)
        g = a.g
        self.launches += 2
        z, mi, gamma, xhat, sel, gbn = self._img_bwd_common(a)
        zd = self._img_conv(a, a.w_dot, a.wdot_task_stride)
        br = view(a.bwd_red, (g.tasks, 2, g.cout))
        dr = view(a.dual_red, (g.tasks, 2, g.cout))
        gd = param(a.gamma_dot, a.gbdot_task_stride, g.tasks, (g.cout,))
        if a.gpdot:
            gbnd = sel * _up(view(a.gpdot, (g.tasks, g.n, g.hp, g.wp, g.cout)), g)
        else:
            gbnd = torch.zeros_like(gbn)
        r, m1, m2, d1, d2 = mi[:, 1], br[:, 0], br[:, 1], dr[:, 0], dr[:, 1]
        e1, e2, e3 = _mean(gbnd), _mean(gbnd * xhat), _mean(gbn * zd)
        q = r * (e3 - d1 * m1 - d2 * m2)
        rdot = -r * r * d2
        xhd = _bc(r) * (zd - _bc(d1) - xhat * _bc(d2))
        proj = gbn - _bc(m1) - xhat * _bc(m2)
        gzd = _bc(gd * r + gamma * rdot) * proj + _bc(gamma * r) * (gbnd - _bc(e1) - xhd * _bc(m2) - xhat * _bc(e2 + q))
        cnt = g.n * g.hz * g.wz
        self._img_outputs(a, self._img_wgrad(a, gzd), (e2 + q) * cnt, e1 * cnt)
        return 0

    # ------------------------------------------------------------------ heads
    @staticmethod
    def _flatten_feat(f, mode):
        """[tasks, n, hw, c] -> X [tasks, n, D]: NCHW flatten order (c major) or spatial mean."""
        if mode == 0:
            return f.permute(0, 1, 3, 2).reshape(f.shape[0], f.shape[1], -1)
        return f.mean(dim=2)

    @staticmethod
    def _unflatten_grad(gx, mode, hw, c):
        if mode == 0:
            return gx.reshape(gx.shape[0], gx.shape[1], c, hw).permute(0, 1, 3, 2).contiguous()
        return (gx / hw)[:, :, None, :].expand(-1, -1, hw, -1).contiguous()

    def xm_head(self, ref, stream):
        a = self._args(ref)
        self.launches += 1
        B, n, ways, c, hw = a.tasks, a.n, a.ways, a.c, a.hw
        D = c * hw if a.mode == 0 else c
        X = self._flatten_feat(view(a.feat, (B, n, hw, c)), a.mode)
        W = param(a.w, a.wb_task_stride, B, (ways, D))
        b = param(a.b, a.wb_task_stride, B, (ways,))
        lab = view(a.labels, (B, a.labels_per_task), torch.int64)
        ys = lab[:, torch.arange(n) * a.label_row_step + a.label_row0]
        logits = torch.einsum('bnd,bwd->bnw', X, W) + b[:, None, :]
        if a.logits:
            view(a.logits, logits.shape).copy_(logits)
        logp = F.log_softmax(logits, dim=2)
        prob = logp.exp()
        onehot = F.one_hot(ys, ways).to(prob.dtype)
        if a.loss:
            view(a.loss, (B,)).copy_(-(logp * onehot).sum(2).mean(1))
        if a.correct:
            view(a.correct, (B,), torch.int32).copy_((logits.argmax(2) == ys).sum(1).int())
        gl = (prob - onehot) / n
        if not a.dual:
            gW, gb, gX = torch.einsum('bnw,bnd->bwd', gl, X), gl.sum(1), torch.einsum('bnw,bwd->bnd', gl, W)
            if a.g_feat:
                view(a.g_feat, (B, n, hw, c)).copy_(self._unflatten_grad(gX, a.mode, hw, c))
        else:
            Xd = self._flatten_feat(view(a.feat_dot, (B, n, hw, c)), a.mode) if a.feat_dot else torch.zeros_like(X)
            Wd = param(a.w_dot, a.wbdot_task_stride, B, (ways, D))
            bd = param(a.b_dot, a.wbdot_task_stride, B, (ways,))
            ld = torch.einsum('bnd,bwd->bnw', Xd, W) + torch.einsum('bnd,bwd->bnw', X, Wd) + bd[:, None, :]
            gld = prob * (ld - (prob * ld).sum(2, keepdim=True)) / n
            gW = torch.einsum('bnw,bnd->bwd', gld, X) + torch.einsum('bnw,bnd->bwd', gl, Xd)
            gb = gld.sum(1)
            gX = torch.einsum('bnw,bwd->bnd', gld, W) + torch.einsum('bnw,bwd->bnd', gl, Wd)
            if a.g_feat_dot:
                view(a.g_feat_dot, (B, n, hw, c)).copy_(self._unflatten_grad(gX, a.mode, hw, c))
        axpy_out(a.out_w, a.out_task_stride, a.base_w, a.base_task_stride, a.scale, B, gW)
        axpy_out(a.out_b, a.out_task_stride, a.base_b, a.base_task_stride, a.scale, B, gb)
        return 0

    def xm_anil_head_scratch_bytes(self, ref):
        a = self._args(ref)
        return a.tasks * (a.steps + 1) * (a.ways * (a.c * a.hw if a.mode == 0 else a.c) + a.ways) * 4

    def xm_anil_head(self, ref, stream):
        a = self._args(ref)
        self.launches += 1
        B, R, ways, c, hw = a.tasks, a.rows, a.ways, a.c, a.hw
        D = c * hw if a.mode == 0 else c
        X = self._flatten_feat(view(a.feat, (B, R, hw, c)), a.mode)
        lab = view(a.labels, (B, R), torch.int64)
        Fs, Fq, ys, yq = X[:, 0::2], X[:, 1::2], lab[:, 0::2], lab[:, 1::2]
        S, Q = Fs.shape[1], Fq.shape[1]
        W = view(a.w, (ways, D)).clone()[None].repeat(B, 1, 1)
        b = view(a.b, (ways,)).clone()[None].repeat(B, 1)
        Ys, Yq = F.one_hot(ys, ways).float(), F.one_hot(yq, ways).float()
        Ws, bs, gls, ps = [W], [b], [], []
        for _ in range(a.steps):
            p = F.softmax(torch.einsum('bnd,bwd->bnw', Fs, Ws[-1]) + bs[-1][:, None], dim=2)
            gl = (p - Ys) / S
            ps.append(p)
            gls.append(gl)
            Ws.append(Ws[-1] - a.lr * torch.einsum('bnw,bnd->bwd', gl, Fs))
            bs.append(bs[-1] - a.lr * gl.sum(1))
        lq = torch.einsum('bnd,bwd->bnw', Fq, Ws[-1]) + bs[-1][:, None]
        logp = F.log_softmax(lq, dim=2)
        view(a.loss, (B,)).copy_(-(logp * Yq).sum(2).mean(1))
        view(a.correct, (B,), torch.int32).copy_((lq.argmax(2) == yq).sum(1).int())
        glq = (logp.exp() - Yq) / Q
        Wb, bb = torch.einsum('bnw,bnd->bwd', glq, Fq), glq.sum(1)
        gFq = torch.einsum('bnw,bwd->bnd', glq, Ws[-1])
        gFs = torch.zeros_like(Fs)
        if not a.first_order:
            for t in reversed(range(a.steps)):
                uW, ub = -a.lr * Wb, -a.lr * bb
                cot = torch.einsum('bnd,bwd->bnw', Fs, uW) + ub[:, None]
                gFs = gFs + torch.einsum('bnw,bwd->bnd', gls[t], uW)
                dl = ps[t] * (cot - (ps[t] * cot).sum(2, keepdim=True)) / S
                Wb = Wb + torch.einsum('bnw,bnd->bwd', dl, Fs)
                bb = bb + dl.sum(1)
                gFs = gFs + torch.einsum('bnw,bwd->bnd', dl, Ws[t])
        gX = torch.zeros_like(X)
        gX[:, 0::2], gX[:, 1::2] = gFs, gFq
        view(a.g_feat, (B, R, hw, c)).copy_(self._unflatten_grad(gX, a.mode, hw, c))
        write_param(a.g_w, a.g_task_stride, B, Wb)
        write_param(a.g_b, a.g_task_stride, B, bb)
        return 0

    # ------------------------------------------------------------------ on-device task sampler
    def xm_sample_tasks(self, ref, stream):
        from oracle import task_sampler_oracle as tso
        a = self._args(ref)
        self.launches += 1
        cs = view(a.class_start, (a.num_classes + 1,), torch.int32).numpy()
        n_items = int(cs[-1])
        data = view(a.data, (n_items, a.channels, a.height, a.width), torch.uint8).numpy()
        seed = a.seed & ((1 << 64) - 1)
        x, y, items, classes = tso.sample_tasks(data, cs, a.tasks, a.ways, a.shots2, seed, a.first_task,
                                                rotate=bool(a.rotate), scale=a.scale, offset=a.offset)
        per = a.ways * a.shots2
        view(a.x, x.shape).copy_(torch.from_numpy(x))
        view(a.y, (a.tasks, per), torch.int64).copy_(torch.from_numpy(y))
        if a.items:
            view(a.items, (a.tasks, per), torch.int32).copy_(torch.from_numpy(items))
        if a.classes:
            view(a.classes, (a.tasks, a.ways), torch.int32).copy_(torch.from_numpy(classes))
        return 0

    # ------------------------------------------------------------------ outer-step helpers
    def xm_accumulate_tasks(self, src, task_stride, tasks, count, dst, accumulate, stream):
        self.launches += 1
        d = view(dst, (count,))
        acc = d.clone() if accumulate else torch.zeros(count)
        for t in range(tasks):
            acc = acc + view(src + 4 * t * task_stride, (count,))
        d.copy_(acc)
        return 0

    def xm_adam_step(self, theta, grad, m, v, count, grad_scale, lr, beta1, beta2, eps, step, stream):
        self.launches += 1
        th, g, mm, vv = (view(p, (count,)) for p in (theta, grad, m, v))
        gs = g * grad_scale
        mm.copy_(beta1 * mm + (1 - beta1) * gs)
        vv.copy_(beta2 * vv + (1 - beta2) * gs * gs)
        denom = vv.sqrt() / (1 - beta2 ** step) ** 0.5 + eps
        th.copy_(th - (lr / (1 - beta1 ** step)) * mm / denom)
        return 0

    def xm_finish_shard(self, loss, correct, tasks, out2, step, stream):
        self.launches += 1
        o = view(out2, (2,))
        l, c = view(loss, (tasks,)), view(correct, (tasks,), torch.int32)
        s = torch.zeros((), dtype=torch.float32)
        k = torch.zeros((), dtype=torch.float32)
        for t in range(tasks):
            s = s + l[t]
            k = k + c[t].float()
        o[0], o[1] = s, k
        if step:
            view(step, (1,), torch.int32).add_(1)
        return 0

    def xm_allreduce_adam(self, comm, ref, stream):
        """Single-rank contract only (comm NULL): reduced = local, then Adam on the first n_params entries."""
        assert not comm, 'the emulator has no peer-memory transport'
        a = self._args(ref)
        red = view(a.reduced, (a.n_total,))
        red.copy_(view(a.local, (a.n_total,)).clone())
        step = int(view(a.step, (1,), torch.int32)[0])
        return self.xm_adam_step(a.theta, a.reduced, a.m, a.v, a.n_params, a.grad_scale, a.lr, a.beta1, a.beta2,
                                 a.eps, step, stream)

    # ------------------------------------------------------------------ config 5: policy-MLP path
    def xm_rl_advantages(self, ref, stream):
        """Contract of csrc/rl.cu::rl_adv_kernel in float64 torch ops (cherry's algorithms)."""
        self.launches += 1
        a = self._args(ref)
        n, sd = a.n, a.state_dim
        for rp in range(a.replays):
            s = view(a.states + 4 * rp * n * sd, (n, sd)).double()
            ns = view(a.next_states + 4 * rp * n * sd, (n, sd)).double()
            r = view(a.rewards + 4 * rp * n, (n,)).double()
            d = view(a.dones + 4 * rp * n, (n,)).double()

            def scan(factor, x):
                out, R = torch.zeros_like(x), 0.0
                for t in reversed(range(n)):
                    R = x[t] + factor * (R * (1.0 - d[t]))
                    out[t] = R
                return out

            def feats(st):
                t = torch.arange(n, dtype=torch.float64) / 100.0
                return torch.cat([st, st ** 2, t[:, None], t[:, None] ** 2, t[:, None] ** 3,
                                  torch.ones(n, 1, dtype=torch.float64)], dim=1)
            ret = scan(a.gamma, r)
            f = feats(s)
            A = f.t() @ f + a.reg * torch.eye(f.size(1), dtype=torch.float64)
            coef = torch.linalg.solve(A, f.t() @ ret)
            v, nv = f @ coef, feats(ns) @ coef
            boot = v * (1 - d) + nv * d
            nxt = torch.cat([boot[1:], torch.zeros(1, dtype=torch.float64)])
            td = r + a.gamma * (1 - d) * nxt - boot
            adv = scan(a.tau * a.gamma, td)
            if a.advantages:
                view(a.advantages + 4 * rp * n, (n,)).copy_(adv.float())
            if n > 1:
                adv = (adv - adv.mean()) / (adv.std() + 1e-8)
            view(a.coef + 4 * rp * n, (n,)).copy_((a.coef_scale * adv).float())
            if a.returns:
                view(a.returns + 4 * rp * n, (n,)).copy_(ret.float())
        return 0

    def xm_rl_sweep_scratch_bytes(self, ref):
        return 64

    def xm_rl_sweep(self, ref, stream):
        """Contract of csrc/rl.cu::rl_sweep_kernel, evaluated with torch.autograd in float64 (double backward for the
        Hessian-vector product, jvp / vjp for the Fisher factor) -- independent of the kernel's hand-derived tangents."""
        import math
        self.launches += 1
        a = self._args(ref)
        n, IN, OUT, H1, H2 = a.n, a.in_dim, a.out_dim, a.h1, a.h2
        P = OUT + H1 * IN + H1 + H2 * H1 + H2 + OUT * H2 + OUT
        actf = torch.tanh if a.activation == 1 else torch.relu
        LOG_EPS = math.log(1e-6)

        def unpack(th):
            o, parts = 0, []
            for shape in ((OUT,), (H1, IN), (H1,), (H2, H1), (H2,), (OUT, H2), (OUT,)):
                k = 1
                for d in shape:
                    k *= d
                parts.append(th[o:o + k].view(*shape))
                o += k
            return parts

        def outputs(th, x):
            sg, W1, b1, W2, b2, W3, b3 = unpack(th)
            h = actf(x @ W1.t() + b1)
            h = actf(h @ W2.t() + b2)
            return h @ W3.t() + b3, torch.clamp(sg, min=LOG_EPS)

        for t in range(a.tasks):
            x = view(a.states + 4 * t * n * IN, (n, IN)).double()
            th0 = view(a.theta + 4 * t * a.theta_task_stride, (P,)).double().clone()
            act = view(a.actions + 4 * t * n * OUT, (n, OUT)).double() if a.actions else None
            coef = view(a.coef + 4 * t * n, (n,)).double() if a.coef else None
            thd = view(a.theta_dot + 4 * t * a.theta_dot_task_stride, (P,)).double().clone() if a.theta_dot else None
            body = slice(OUT, OUT + H1 * IN + H1 + H2 * H1 + H2)
            if thd is not None and (a.head_only & 2):
                thd[body] = 0.0
            lamo = view(a.logstd_old + 4 * t * OUT, (OUT,)).double() if a.logstd_old else None
            muo = view(a.mu_old + 4 * t * n * OUT, (n, OUT)).double() if a.mu_old else None

            def log_prob(mu, lam, actions):
                return (-(actions - mu) ** 2 / (2 * torch.exp(2 * lam)) - lam - 0.5 * math.log(2 * math.pi)).mean(dim=1)

            def loss_fn(th):
                mu, lam = outputs(th, x)
                if a.loss == 0:
                    return (coef * log_prob(mu, lam, act)).sum(), torch.zeros((), dtype=torch.float64)
                lp, lpo = log_prob(mu, lam, act), log_prob(muo, lamo, act)
                kl = (lamo - lam + (torch.exp(2 * lam) + (mu - muo) ** 2) / (2 * torch.exp(2 * lamo)) - 0.5).sum()
                ratio = torch.exp(lp - lpo)
                if a.clip > 0:          # ppo.policy_loss: -mean(min(r A, clamp(r) A)) with coef = -A / n
                    val = torch.max(coef * ratio, coef * ratio.clamp(1.0 - a.clip, 1.0 + a.clip))
                else:
                    val = coef * ratio
                return val.sum(), a.kl_scale * kl

            res, l, kl = None, None, None
            if a.loss == 2:                                         # Fisher factor J^T F J theta_dot
                (mu, lam), (mud, lamd) = torch.autograd.functional.jvp(lambda th: outputs(th, x), th0, thd)
                cot = (a.kl_scale * mud * torch.exp(-2 * lamo), a.kl_scale * 2.0 * n * lamd)
                _o, res = torch.autograd.functional.vjp(lambda th: outputs(th, x), th0, cot)
            else:
                th = th0.requires_grad_()
                l, kl = loss_fn(th)
                if a.what == 1:
                    res = torch.autograd.grad(l, th)[0]
                elif a.what == 2:
                    g = torch.autograd.grad(l, th, create_graph=True)[0]
                    res = torch.autograd.grad((g * thd).sum(), th)[0]
                if a.mu_out:
                    view(a.mu_out + 4 * t * n * OUT, (n, OUT)).copy_(outputs(th0.detach(), x)[0].float())
            if res is not None:
                if a.head_only & 1:
                    res = res.clone()
                    res[body] = 0.0
                base = view(a.base + 4 * t * a.base_task_stride, (P,)).double() if a.base else torch.zeros(P, dtype=torch.float64)
                view(a.out + 4 * t * a.out_task_stride, (P,)).copy_((base + a.scale * res).float())
            if a.task_loss and l is not None:
                view(a.task_loss + 4 * t, (1,))[0] = float(l.detach())
            if a.task_kl and kl is not None:
                view(a.task_kl + 4 * t, (1,))[0] = float(kl.detach())
        return 0

    def xm_bn_ema(self, rm, rv, stats, n_outer, outer_stride, n_inner, inner_stride, C, momentum, stream):
        self.launches += 1
        m, v = view(rm, (C,)), view(rv, (C,))
        for o in range(n_outer):
            for i in range(n_inner):
                s = view(stats + 4 * (o * outer_stride + i * inner_stride), (2, C))
                m.copy_((1 - momentum) * m + momentum * s[0])
                v.copy_((1 - momentum) * v + momentum * s[1])
        return 0
