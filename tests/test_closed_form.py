"""The closed forms behind the fused image block (exploring_meta_b200/csrc/img_block.cu), checked
in float64 against torch.autograd on the block the reference defines (core_functions/vision_models.py:188-193:
conv3x3 -> BatchNorm(train) -> ReLU -> MaxPool 2x2, and the stride-2 / no-pool variant of the Omniglot networks).

Independent of the kernels and of the C-ABI emulator (which evaluates the same contract densely): this pins the
DERIVATION -- statistics from the Gram matrix, weight gradient = gamma r (S - m1 sx - m2 XH), and its tangent."""
import pytest
import torch
import torch.nn.functional as F

EPS = 1e-5


@pytest.fixture(autouse=True)
def _float64_default():
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)


def _block(x, w, gamma, beta, stride, pool):
    z = F.conv2d(x, w, None, stride=stride, padding=1)
    y = F.batch_norm(z, None, None, gamma, beta, training=True, eps=EPS)
    a = F.relu(y)
    return F.max_pool2d(a, 2, 2) if pool else a


def _closed_form(x, w, gamma, beta, gp, wd, gd, bd, gpd, stride, pool):
    """Everything the image-block kernels compute, written with the Gram matrix and the winner gather only."""
    n, cin, H, W = x.shape
    cout, K = w.shape[0], 9 * cin
    cols = F.unfold(x, 3, padding=1, stride=stride)                 # [n, K, positions]
    X = cols.permute(0, 2, 1).reshape(-1, K)                        # im2col: [N, K], k = ci*9 + kh*3 + kw
    N = X.shape[0]
    G, sx = X.t() @ X, X.sum(0)                                     # xm_img_gram
    wv, wdv = w.reshape(cout, K), wd.reshape(cout, K)
    # forward statistics without a pass over the pixels
    mean = (wv @ sx) / N
    var = torch.einsum('ck,kl,cl->c', wv, G, wv) / N - mean ** 2
    r = 1.0 / torch.sqrt(var + EPS)
    d1 = (wdv @ sx) / N
    d2 = r * (torch.einsum('ck,kl,cl->c', wdv, G, wv) / N - mean * d1)
    # winners: per output element of the block, the position that wins the pool and passes the ReLU
    z = (X @ wv.t())                                                # [N, cout]  (only used to LOCATE winners / values there)
    hz, wz = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    y = gamma * (z - mean) * r + beta
    if pool:
        yi = y.reshape(n, hz, wz, cout).permute(0, 3, 1, 2)
        pooled, idx = F.max_pool2d(yi, 2, 2, return_indices=True)   # idx: flat position inside the (hz, wz) map
        alive = pooled > 0
        base = (torch.arange(n) * hz * wz).view(n, 1, 1, 1)
        win = (idx + base)                                          # row of X / z of every winner
    else:
        alive = (y > 0).reshape(n, hz, wz, cout).permute(0, 3, 1, 2)
        win = torch.arange(N).reshape(n, hz, wz, 1).permute(0, 3, 1, 2).expand(-1, cout, -1, -1)
    ch = torch.arange(cout).view(1, cout, 1, 1).expand_as(win)
    g = torch.where(alive, gp, torch.zeros_like(gp))                # cotangent at the winners
    gdot = torch.where(alive, gpd, torch.zeros_like(gpd))
    zsel, zdsel = z[win, ch], (X @ wdv.t())[win, ch]
    xhat = (zsel - mean.view(1, -1, 1, 1)) * r.view(1, -1, 1, 1)
    Xw = X[win]                                                     # [n, cout, hp, wp, K]: the winners' patches
    S = torch.einsum('nchw,nchwk->ck', g, Xw)
    Sd = torch.einsum('nchw,nchwk->ck', gdot, Xw)
    s1, s2 = g.sum((0, 2, 3)), (g * xhat).sum((0, 2, 3))
    m1, m2 = s1 / N, s2 / N
    XH = r[:, None] * (wv @ G - mean[:, None] * sx[None])
    dW = (gamma * r)[:, None] * (S - m1[:, None] * sx[None] - m2[:, None] * XH)
    out = {'dW': dW.reshape_as(w), 'dgamma': s2, 'dbeta': s1}
    # tangent of the block output (xm_img_dual_fwd)
    xhd = r.view(1, -1, 1, 1) * (zdsel - d1.view(1, -1, 1, 1) - xhat * d2.view(1, -1, 1, 1))
    pdot = gd.view(1, -1, 1, 1) * xhat + gamma.view(1, -1, 1, 1) * xhd + bd.view(1, -1, 1, 1)
    out['pdot'] = torch.where(alive, pdot, torch.zeros_like(pdot))
    # tangent of the backward (xm_img_dual_bwd)
    e1, e2, e3 = gdot.sum((0, 2, 3)) / N, (gdot * xhat).sum((0, 2, 3)) / N, (g * zdsel).sum((0, 2, 3)) / N
    q = r * (e3 - d1 * m1 - d2 * m2)
    m2dot = e2 + q
    coef, gr = gd * r - gamma * r * r * d2, gamma * r
    XHD = r[:, None] * (wdv @ G - d1[:, None] * sx[None] - d2[:, None] * XH)
    dWd = coef[:, None] * (S - m1[:, None] * sx[None] - m2[:, None] * XH) + \
        gr[:, None] * (Sd - e1[:, None] * sx[None] - m2[:, None] * XHD - m2dot[:, None] * XH)
    out.update({'dWdot': dWd.reshape_as(w), 'dgammadot': m2dot * N, 'dbetadot': e1 * N})
    return out


@pytest.mark.parametrize('stride,pool,cin,H,W', [(1, True, 3, 8, 12), (1, True, 1, 6, 6), (2, False, 1, 9, 9), (2, False, 1, 14, 14)])
def test_image_block_closed_forms_match_autograd(stride, pool, cin, H, W):
    torch.manual_seed(0)
    n, cout = 3, 5
    x = torch.randn(n, cin, H, W)
    w = (torch.randn(cout, cin, 3, 3) * 0.4).requires_grad_()
    gamma = (torch.rand(cout) + 0.2).requires_grad_()
    beta = (torch.randn(cout) * 0.3).requires_grad_()
    wd, gd, bd = torch.randn_like(w), torch.randn_like(gamma), torch.randn_like(beta)      # tangent direction
    p = _block(x, w, gamma, beta, stride, pool)
    gp, gpd = torch.randn_like(p), torch.randn_like(p)                                      # cotangent and its tangent
    cf = _closed_form(x, w.detach(), gamma.detach(), beta.detach(), gp, wd, gd, bd, gpd, stride, pool)

    # first order: VJP of the block
    dW, dgamma, dbeta = torch.autograd.grad(p, (w, gamma, beta), gp, create_graph=True)
    assert torch.allclose(cf['dW'], dW, rtol=1e-9, atol=1e-10)
    assert torch.allclose(cf['dgamma'], dgamma, rtol=1e-9, atol=1e-10)
    assert torch.allclose(cf['dbeta'], dbeta, rtol=1e-9, atol=1e-10)

    # tangent of the forward: JVP of the block in direction (wd, gd, bd)
    _, pdot = torch.autograd.functional.jvp(lambda a, b, c: _block(x, a, b, c, stride, pool),
                                            (w.detach(), gamma.detach(), beta.detach()), (wd, gd, bd))
    assert torch.allclose(cf['pdot'], pdot, rtol=1e-9, atol=1e-10)

    # tangent of the backward: d/d(eps) of the VJP at (theta + eps*dir, gp + eps*gpd) -- what second-order MAML needs
    def vjp(a, b, c, cot):
        a, b, c = a.requires_grad_(), b.requires_grad_(), c.requires_grad_()
        return torch.autograd.grad(_block(x, a, b, c, stride, pool), (a, b, c), cot, create_graph=True)
    _, tang = torch.autograd.functional.jvp(vjp, (w.detach().clone(), gamma.detach().clone(), beta.detach().clone(), gp),
                                            (wd, gd, bd, gpd))
    assert torch.allclose(cf['dWdot'], tang[0], rtol=1e-8, atol=1e-9)
    assert torch.allclose(cf['dgammadot'], tang[1], rtol=1e-8, atol=1e-9)
    assert torch.allclose(cf['dbetadot'], tang[2], rtol=1e-8, atol=1e-9)
