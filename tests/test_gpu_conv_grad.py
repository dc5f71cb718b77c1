"""-m gpu: generic convolution gradients (csrc/conv_generic.cu) through the C ABI against torch autograd on bf16-rounded
operands, for every layer geometry of LinkNet34 (lib/models/linknet.py:39-62)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

import snb_b200  # noqa: F401
from snb_b200 import _native as N
from snb_b200 import engine as E

pytestmark = pytest.mark.gpu


def bf(t):
    return t.to(torch.bfloat16).float()


def slab_from(x_nchw, extra=0):
    n, c, h, w = x_nchw.shape
    s = E.Slab(n, h, w, c + extra, "cuda")
    s.t.zero_()
    s.t[..., :c].copy_(x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16))
    return s


def nchw(slab, c):
    return slab.t[..., :c].float().permute(0, 3, 1, 2).contiguous()


def geom(n, bh, bw, bc, bcs, sh, sw, sc, scs, k, stride, pad):
    g = N.ConvGeom()
    g.n, g.big_h, g.big_w, g.big_c, g.big_cstride = n, bh, bw, bc, bcs
    g.small_h, g.small_w, g.small_c, g.small_cstride = sh, sw, sc, scs
    g.kh = g.kw = k
    g.stride, g.pad = stride, pad
    return g


def rel_l2(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


# (cin, cout, k, stride, pad, h, w): resnet34 blocks, shortcut, decoder 1x1, head convs, the im2col stem
CONVS = [(64, 64, 3, 1, 1, 16, 24), (64, 128, 3, 2, 1, 16, 24), (64, 128, 1, 2, 0, 16, 24), (256, 64, 1, 1, 0, 8, 8),
         (16, 64, 1, 1, 0, 32, 32), (32, 32, 3, 1, 0, 33, 33), (32, 1, 2, 1, 1, 31, 31), (160, 64, 1, 1, 0, 32, 32),
         (512, 512, 3, 1, 1, 4, 4), (3, 64, 7, 2, 3, 32, 32)]


@pytest.mark.parametrize("cin,cout,k,stride,pad,h,w", CONVS)
def test_conv2d_gradients(cuda, cin, cout, k, stride, pad, h, w):
    g = torch.Generator(device="cuda").manual_seed(cin * 31 + cout + k)
    n = 3
    x = bf(torch.randn((n, cin, h, w), device="cuda", generator=g)).requires_grad_(True)
    wt = (torch.randn((cout, cin, k, k), device="cuda", generator=g) * (2.0 / (cin * k * k)) ** 0.5).requires_grad_(True)
    y = F.conv2d(x, wt, None, stride=stride, padding=pad)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dy)
    oh, ow = y.shape[2:]
    xs, dys = slab_from(x.detach(), extra=8), slab_from(dy, extra=16)
    gm = geom(n, h, w, cin, xs.c, oh, ow, cout, dys.c, k, stride, pad)
    dx = E.Slab(n, h, w, cin + 8, "cuda")
    dx.t.fill_(3.0)
    st = N.stream_ptr()
    N.check(N.lib().snb_conv_generic_dgrad(ctypes.byref(gm), N.c_vp(dys.t.data_ptr()), N.ptr(wt.detach()), N.c_vp(0),
                                           N.c_vp(dx.t.data_ptr()), st))
    dw = torch.full_like(wt, 7.0)
    N.check(N.lib().snb_conv_generic_wgrad(ctypes.byref(gm), N.c_vp(xs.t.data_ptr()), N.c_vp(dys.t.data_ptr()), N.ptr(dw), st))
    torch.cuda.synchronize()
    # dgrad used the fp32 weights; torch used them too (bf16 only on activations / gradients); output rounded to bf16
    assert rel_l2(nchw(dx, cin), x.grad) < 6e-3
    assert torch.all(dx.t[..., cin:] == 3.0)
    assert rel_l2(dw, wt.grad) < 2e-3
    # forward of the same geometry (what a ConvTranspose2d dgrad runs)
    ys = E.Slab(n, oh, ow, cout + 16, "cuda")
    bias = torch.randn(cout, device="cuda", generator=g)
    N.check(N.lib().snb_conv_generic_fwd(ctypes.byref(gm), N.c_vp(xs.t.data_ptr()), N.ptr(wt.detach()), N.ptr(bias),
                                         N.c_vp(ys.t.data_ptr()), st))
    assert rel_l2(nchw(ys, cout), (y + bias.view(1, -1, 1, 1)).detach()) < 6e-3


# ConvTranspose2d(cin_t, cout_t, k, stride, pad): decoder deconv2 (k4 s2 p1) and finaldeconv1 (k3 s2 p0)
@pytest.mark.parametrize("cin_t,cout_t,k,stride,pad,h,w", [(128, 128, 4, 2, 1, 8, 8), (16, 16, 4, 2, 1, 16, 16), (64, 32, 3, 2, 0, 16, 20)])
def test_conv_transpose2d_gradients(cuda, cin_t, cout_t, k, stride, pad, h, w):
    g = torch.Generator(device="cuda").manual_seed(cin_t + cout_t + k)
    n = 2
    x = bf(torch.randn((n, cin_t, h, w), device="cuda", generator=g)).requires_grad_(True)
    wt = (torch.randn((cin_t, cout_t, k, k), device="cuda", generator=g) * (2.0 / (cin_t * k)) ** 0.5).requires_grad_(True)
    y = F.conv_transpose2d(x, wt, None, stride=stride, padding=pad)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dy)
    oh, ow = y.shape[2:]
    xs, dys = slab_from(x.detach()), slab_from(dy)
    # underlying conv: big = the transposed conv's OUTPUT (ci = cout_t), small = its INPUT (co = cin_t)
    gm = geom(n, oh, ow, cout_t, dys.c, h, w, cin_t, xs.c, k, stride, pad)
    st = N.stream_ptr()
    dx = E.Slab(n, h, w, cin_t, "cuda")
    N.check(N.lib().snb_conv_generic_fwd(ctypes.byref(gm), N.c_vp(dys.t.data_ptr()), N.ptr(wt.detach()), N.c_vp(0),
                                         N.c_vp(dx.t.data_ptr()), st))
    dw = torch.empty_like(wt)
    N.check(N.lib().snb_conv_generic_wgrad(ctypes.byref(gm), N.c_vp(dys.t.data_ptr()), N.c_vp(xs.t.data_ptr()), N.ptr(dw), st))
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx, cin_t), x.grad) < 6e-3
    assert rel_l2(dw, wt.grad) < 2e-3
    # its forward is the dgrad map of the underlying conv
    ys = E.Slab(n, oh, ow, cout_t, "cuda")
    N.check(N.lib().snb_conv_generic_dgrad(ctypes.byref(gm), N.c_vp(xs.t.data_ptr()), N.ptr(wt.detach()), N.c_vp(0),
                                           N.c_vp(ys.t.data_ptr()), st))
    assert rel_l2(nchw(ys, cout_t), y.detach()) < 6e-3


def test_geometry_errors(cuda):
    gm = geom(1, 16, 16, 8, 8, 9, 8, 8, 8, 3, 2, 1)     # small_h should be 8
    t = torch.zeros(1, device="cuda")
    with pytest.raises(AssertionError):
        N.check(N.lib().snb_conv_generic_fwd(ctypes.byref(gm), N.ptr(t), N.ptr(t), N.c_vp(0), N.ptr(t), N.stream_ptr()))


@pytest.mark.parametrize("c,abn,slope,res,after", [(64, False, 0.0, True, False), (128, True, 0.01, True, True),
                                                  (32, True, 0.01, False, False), (256, False, -1.0, False, False)])
def test_bn_train_backward_nhwc(cuda, c, abn, slope, res, after):
    """snb_bn_backward_nhwc against torch autograd through batch_norm(training) + activation + residual."""
    g = torch.Generator(device="cuda").manual_seed(c + 5)
    n, h, w = 3, 10, 12
    x = bf(torch.randn((n, c, h, w), device="cuda", generator=g) * 1.5 + 0.3).requires_grad_(True)
    r = bf(torch.randn((n, c, h, w), device="cuda", generator=g)).requires_grad_(True)
    wgt = ((torch.rand(c, device="cuda", generator=g) + 0.5) *
           torch.where(torch.rand(c, device="cuda", generator=g) < 0.3, -1.0, 1.0)).requires_grad_(True)
    bias = (torch.randn(c, device="cuda", generator=g) * 0.2).requires_grad_(True)
    gamma = wgt.abs() + 1e-5 if abn else wgt
    y = F.batch_norm(x, None, None, gamma, bias, training=True, eps=1e-5)
    if res and not after:
        y = y + r
    if slope >= 0:
        y = F.leaky_relu(y, slope)
    if res and after:
        y = y + r
    dout = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dout)
    # forward on the device (provides scale / shift / mean / var), then the backward kernel
    xs, rs, gs = slab_from(x.detach()), slab_from(r.detach()), slab_from(dout)
    out = E.Slab(n, h, w, c, "cuda")
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    op = E.BnTrainOp(xs.view(), out.view(), (wgt.detach(), bias.detach(), rm, rv, 1e-5, 0.1), abn, slope,
                     rs.view() if res else None, after)
    st = N.stream_ptr()
    op(st)
    dx, dres = E.Slab(n, h, w, c, "cuda"), E.Slab(n, h, w, c, "cuda")
    dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    work = torch.zeros(3 * c + 2, dtype=torch.float64, device="cuda")
    rb = res and not after
    N.check(N.lib().snb_bn_backward_nhwc(
        N.c_vp(xs.t.data_ptr()), c, N.c_vp(gs.t.data_ptr()), c, n * h * w, c, N.ptr(op.scale), N.ptr(op.shift), N.ptr(op.mean),
        N.ptr(op.var), N.ptr(wgt.detach()), 1 if abn else 0, 1e-5, slope, N.c_vp(rs.t.data_ptr() if rb else 0), c,
        N.c_vp(dx.t.data_ptr()), c, N.c_vp(dres.t.data_ptr() if rb else 0), c, N.ptr(dgamma), N.ptr(dbeta), N.ptr(work), st))
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx, c), x.grad) < 1e-2
    assert rel_l2(dgamma, wgt.grad) < 5e-3 and rel_l2(dbeta, bias.grad) < 5e-3
    if rb:
        assert rel_l2(nchw(dres, c), r.grad) < 6e-3


def test_maxpool_backward_ew_and_channel_sum(cuda):
    g = torch.Generator(device="cuda").manual_seed(11)
    n, c, h, w = 2, 64, 18, 22
    x = bf(F.relu(torch.randn((n, c, h, w), device="cuda", generator=g))).requires_grad_(True)     # many exact ties at 0
    y = F.max_pool2d(x, 3, 2, 1)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dy)
    xs, dys = slab_from(x.detach()), slab_from(dy)
    dxs = E.Slab(n, h, w, c, "cuda")
    st = N.stream_ptr()
    N.check(N.lib().snb_maxpool3x3s2_backward(N.c_vp(xs.t.data_ptr()), n, h, w, c, c, N.c_vp(dys.t.data_ptr()), c,
                                              N.c_vp(dxs.t.data_ptr()), c, st))
    assert rel_l2(nchw(dxs, c), x.grad) < 4e-3                     # sums of up to 4 bf16 gradients, rounded once
    # elementwise helpers
    a, b = slab_from(dy), slab_from(bf(torch.randn(y.shape, device="cuda", generator=g)))
    o = E.Slab(n, y.shape[2], y.shape[3], c, "cuda")
    px = n * y.shape[2] * y.shape[3]
    N.check(N.lib().snb_ew_nhwc(N.c_vp(a.t.data_ptr()), c, N.c_vp(b.t.data_ptr()), c, N.c_vp(o.t.data_ptr()), c, px, c, 0, 0.0, st))
    assert torch.equal(o.t, (a.t.float() + b.t.float()).to(torch.bfloat16))
    N.check(N.lib().snb_ew_nhwc(N.c_vp(a.t.data_ptr()), c, N.c_vp(b.t.data_ptr()), c, N.c_vp(o.t.data_ptr()), c, px, c, 1, 0.01, st))
    assert torch.equal(o.t, torch.where(b.t.float() > 0, a.t.float(), a.t.float() * 0.01).to(torch.bfloat16))
    s = torch.empty(c, device="cuda")
    work = torch.zeros(3 * c + 2, dtype=torch.float64, device="cuda")
    N.check(N.lib().snb_channel_sum_nhwc(N.c_vp(a.t.data_ptr()), px, c, c, N.ptr(s), N.ptr(work), st))
    assert (s - a.t.float().sum(dim=(0, 1, 2))).abs().max().item() < 1e-3
