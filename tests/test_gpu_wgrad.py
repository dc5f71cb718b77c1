"""-m gpu: tcgen05 weight gradients (csrc/conv_wgrad.cu) through the C ABI against torch autograd on bf16-rounded
operands, in the packed layout of the forward weights, for every layer geometry of LinkNet34 (lib/models/linknet.py)."""
import pytest
import torch
import torch.nn.functional as F

import snb_b200  # noqa: F401
from snb_b200 import _native as N
from snb_b200 import engine as E

pytestmark = pytest.mark.gpu
F64 = torch.float64


def bf(t):
    return t.to(torch.bfloat16).float()


def slab_from(x_nchw, extra=0):
    n, c, h, w = x_nchw.shape
    s = E.Slab(n, h, w, c + extra, "cuda")
    s.t.copy_(torch.randn(s.t.shape, device="cuda").to(torch.bfloat16))      # neighbours in the slab hold garbage
    s.t[..., :c].copy_(x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16))
    return s


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def run_wgrad(kind, xs, cin, dys, cout, shape, valid=0):
    dw = torch.zeros(shape, dtype=torch.float32, device="cuda")
    op = E.WgradOp(kind, xs.view(0, cin), dys.view(0, cout), dw, valid=valid)
    op(N.stream_ptr())
    torch.cuda.synchronize()
    return dw, op


# (cin, cout, h, w, n): resnet34 stride-1 blocks at their LinkNet34 sizes (batch 8 x 256^2 -> 64^2 ... 8^2), ragged sizes
@pytest.mark.parametrize("cin,cout,h,w,n", [(64, 64, 64, 64, 2), (128, 128, 32, 32, 3), (256, 256, 16, 16, 2),
                                              (512, 512, 8, 8, 8), (64, 64, 7, 9, 1), (96, 160, 20, 12, 2)])
def test_conv3x3_weight_gradient(cuda, cin, cout, h, w, n):
    g = torch.Generator(device="cuda").manual_seed(cin + cout + h)
    x = bf(torch.randn((n, cin, h, w), device="cuda", generator=g))
    wt = torch.zeros((cout, cin, 3, 3), device="cuda", requires_grad=True)
    y = F.conv2d(x, wt, None, padding=1)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dy)
    xs, dys = slab_from(x, extra=8), slab_from(dy, extra=16)
    dw, op = run_wgrad(N.CONV_3X3, xs, cin, dys, cout, (9, cout, cin))
    want = E.pack_conv3x3(wt.grad, F64)
    assert rel_l2(dw, want) < 1e-4, rel_l2(dw, want)
    assert op.flops == 2.0 * n * h * w * cin * cout * 9
    # accumulation: a second launch doubles the result; padded gradient buffers keep their padding untouched
    dwp = torch.full((9, cout + 8, cin + 32), 5.0, dtype=torch.float32, device="cuda")
    op2 = E.WgradOp(N.CONV_3X3, xs.view(0, cin), dys.view(0, cout), dwp)
    op2(N.stream_ptr())
    op2(N.stream_ptr())
    torch.cuda.synchronize()
    assert rel_l2(dwp[:, :cout, :cin] - 5.0, 2 * want) < 1e-4
    assert torch.all(dwp[:, cout:] == 5.0) and torch.all(dwp[:, :, cin:] == 5.0)


@pytest.mark.parametrize("cin,cout,h,w", [(512, 128, 8, 8), (128, 256, 16, 16), (32, 64, 64, 64), (160, 64, 32, 48)])
def test_conv1x1_weight_gradient(cuda, cin, cout, h, w):
    """decoder conv1 / conv3 and the stem as a conv1x1 over its im2col rows (K = 160, not a multiple of 64)"""
    g = torch.Generator(device="cuda").manual_seed(cin * 3 + cout)
    n = 2
    x = bf(torch.randn((n, cin, h, w), device="cuda", generator=g))
    wt = torch.zeros((cout, cin, 1, 1), device="cuda", requires_grad=True)
    y = F.conv2d(x, wt)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dy)
    dw, _ = run_wgrad(N.CONV_1X1, slab_from(x), cin, slab_from(dy, extra=8), cout, (1, cout, cin))
    assert rel_l2(dw, E.pack_conv1x1(wt.grad, F64)) < 1e-4


@pytest.mark.parametrize("cin,cout,h,w", [(64, 128, 32, 32), (256, 512, 16, 16), (64, 128, 12, 20)])
def test_stride2_conv3x3_weight_gradient_via_space_to_depth(cuda, cin, cout, h, w):
    """resnet34 down-sampling conv (k3 s2 p1) = 4-tap conv over the space-to-depth copy: the packed gradient holds the
    9 real (tap, parity) blocks of pack_conv3x3_s2; the 1x1 s2 shortcut reads the first channel quarter of the same copy"""
    g = torch.Generator(device="cuda").manual_seed(cin + cout + w)
    n = 2
    x = bf(torch.randn((n, cin, h, w), device="cuda", generator=g))
    wt = torch.zeros((cout, cin, 3, 3), device="cuda", requires_grad=True)
    wd = torch.zeros((cout, cin, 1, 1), device="cuda", requires_grad=True)
    y = F.conv2d(x, wt, None, stride=2, padding=1)
    yd = F.conv2d(x, wd, None, stride=2)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    dyd = bf(torch.randn(yd.shape, device="cuda", generator=g))
    (y * dy).sum().backward()
    (yd * dyd).sum().backward()
    xs = slab_from(x)
    x4 = E.Slab(n, h // 2, w // 2, 4 * cin, "cuda")
    N.check(N.lib().snb_space_to_depth2(N.c_vp(xs.t.data_ptr()), n, h, w, cin, cin, N.c_vp(x4.t.data_ptr()), 4 * cin, N.stream_ptr()))
    dw, _ = run_wgrad(N.CONV_2X2, x4, 4 * cin, slab_from(dy), cout, (4, cout, 4 * cin), valid=1)
    want = E.pack_conv3x3_s2(wt.grad, F64)
    mask = E.pack_conv3x3_s2(torch.ones_like(wt), F64) != 0          # the 7 unused (tap, parity) blocks hold other products
    assert rel_l2(dw * mask, want) < 1e-4
    dwd, _ = run_wgrad(N.CONV_1X1, x4, cin, slab_from(dyd), cout, (1, cout, cin))
    assert rel_l2(dwd, E.pack_conv1x1(wd.grad, F64)) < 1e-4


@pytest.mark.parametrize("c,h,w", [(128, 8, 8), (64, 16, 16), (32, 32, 24)])
def test_conv_transpose4x4_weight_gradient(cuda, c, h, w):
    """decoder deconv2: ConvTranspose2d(c, c, 4, 2, 1): 4 phases x 4 taps, dY phases are stride-2 views"""
    g = torch.Generator(device="cuda").manual_seed(c + h)
    n = 2
    x = bf(torch.randn((n, c, h, w), device="cuda", generator=g))
    wt = torch.zeros((c, c, 4, 4), device="cuda", requires_grad=True)
    y = F.conv_transpose2d(x, wt, None, stride=2, padding=1)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dy)
    dw, _ = run_wgrad(N.CONVT_4X4_S2, slab_from(x), c, slab_from(dy), c, (16, c, c))
    assert rel_l2(dw, E.pack_convT4x4(wt.grad, F64)) < 1e-4


def test_linknet_head_weight_gradients(cuda):
    """finaldeconv1 (ConvTranspose k3 s2 uncropped, 64 -> 32), finalconv2 (valid conv3x3 32 -> 32), finalconv3 (conv k2 p1,
    32 -> 1): lib/models/linknet.py:58-62"""
    g = torch.Generator(device="cuda").manual_seed(9)
    n, h, w = 2, 24, 40
    x = bf(torch.randn((n, 64, h, w), device="cuda", generator=g))
    w1 = torch.zeros((64, 32, 3, 3), device="cuda", requires_grad=True)
    y1 = F.conv_transpose2d(x, w1, None, stride=2)
    dy1 = bf(torch.randn(y1.shape, device="cuda", generator=g))
    y1.backward(dy1)
    dw, _ = run_wgrad(N.CONVT_3X3_S2_FULL, slab_from(x), 64, slab_from(dy1), 32, (16, 32, 64))
    want = E.pack_convT3x3(w1.grad, 64, 32, F64)
    mask = E.pack_convT3x3(torch.ones_like(w1), 64, 32, F64) != 0
    assert rel_l2(dw * mask, want) < 1e-4
    x2 = bf(torch.randn((n, 32, 2 * h + 1, 2 * w + 1), device="cuda", generator=g))
    w2 = torch.zeros((32, 32, 3, 3), device="cuda", requires_grad=True)
    y2 = F.conv2d(x2, w2)
    dy2 = bf(torch.randn(y2.shape, device="cuda", generator=g))
    y2.backward(dy2)
    dw2, _ = run_wgrad(N.CONV_3X3, slab_from(x2), 32, slab_from(dy2), 32, (9, 32, 32), valid=1)
    assert rel_l2(dw2, E.pack_conv3x3(w2.grad, F64)) < 1e-4
    x3 = bf(torch.randn((n, 32, 2 * h - 1, 2 * w - 1), device="cuda", generator=g))
    w3 = torch.zeros((1, 32, 2, 2), device="cuda", requires_grad=True)
    y3 = F.conv2d(x3, w3, None, padding=1)
    dy3 = bf(torch.randn(y3.shape, device="cuda", generator=g))
    y3.backward(dy3)
    dw3, _ = run_wgrad(N.CONV_2X2, slab_from(x3), 32, slab_from(dy3, extra=7), 1, (4, 32, 32))
    assert rel_l2(dw3[:, :1], E.pack_conv2x2(w3.grad, 32, 1, F64)) < 1e-4
    assert torch.count_nonzero(dw3[:, 1:]) == 0


# ------------------------------------------------------------------------------------------------ input gradients
def nchw(slab, c):
    return slab.t[..., :c].float().permute(0, 3, 1, 2).contiguous()


def zero_bias(c):
    return torch.zeros(c, device="cuda")


def test_stride2_block_input_gradient(cuda):
    """resnet34 down-sampling block input: d x = dgrad(conv3x3 s2 p1) + dgrad(conv1x1 s2) = SNB_CONV_2X2_ADJ over the output
    gradient into space-to-depth channels, the shortcut's conv1x1 added in place on the first channel quarter (residual
    epilogue), then snb_depth_to_space2 (accumulating onto a gradient already in the slab)."""
    g = torch.Generator(device="cuda").manual_seed(21)
    n, cin, cout, h, w = 2, 64, 128, 24, 40
    x = bf(torch.randn((n, cin, h, w), device="cuda", generator=g)).requires_grad_(True)
    wt = bf(torch.randn((cout, cin, 3, 3), device="cuda", generator=g) * 0.05)
    wd = bf(torch.randn((cout, cin, 1, 1), device="cuda", generator=g) * 0.1)
    y, yd = F.conv2d(x, wt, None, stride=2, padding=1), F.conv2d(x, wd, None, stride=2)
    dy, dyd = bf(torch.randn(y.shape, device="cuda", generator=g)), bf(torch.randn(yd.shape, device="cuda", generator=g))
    prior = bf(torch.randn(x.shape, device="cuda", generator=g))
    (y * dy).sum().backward()
    (yd * dyd).sum().backward()
    dys, dyds = slab_from(dy), slab_from(dyd)
    dx4 = E.Slab(n, h // 2, w // 2, 4 * cin, "cuda")
    st = N.stream_ptr()
    E.ConvOp(N.CONV_2X2_ADJ, dys.view(0, cout), dx4.view(), E.pack_conv3x3_s2_dgrad(wt), zero_bias(4 * cin), relu=False)(st)
    E.ConvOp(N.CONV_1X1, dyds.view(0, cout), dx4.view(0, cin), E.pack_conv_dgrad(wd), zero_bias(cin), relu=False,
             residual=dx4.view(0, cin))(st)
    dx = slab_from(prior)
    N.check(N.lib().snb_depth_to_space2(N.c_vp(dx4.t.data_ptr()), n, h, w, cin, 4 * cin, N.c_vp(dx.t.data_ptr()), cin, 1, st))
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx, cin), x.grad + prior) < 1e-2


@pytest.mark.parametrize("c,h,w", [(128, 8, 8), (32, 16, 24)])
def test_conv_transpose4x4_input_gradient(cuda, c, h, w):
    """decoder deconv2 (k4 s2 p1): conv3x3 over the space-to-depth copy of the output gradient"""
    g = torch.Generator(device="cuda").manual_seed(c)
    n = 2
    x = bf(torch.randn((n, c, h, w), device="cuda", generator=g)).requires_grad_(True)
    wt = bf(torch.randn((c, c, 4, 4), device="cuda", generator=g) * 0.05)
    y = F.conv_transpose2d(x, wt, None, stride=2, padding=1)
    dy = bf(torch.randn(y.shape, device="cuda", generator=g))
    y.backward(dy)
    dys = slab_from(dy)
    d4 = E.Slab(n, h, w, 4 * c, "cuda")
    st = N.stream_ptr()
    N.check(N.lib().snb_space_to_depth2(N.c_vp(dys.t.data_ptr()), n, 2 * h, 2 * w, c, c, N.c_vp(d4.t.data_ptr()), 4 * c, st))
    dx = E.Slab(n, h, w, c, "cuda")
    E.ConvOp(N.CONV_3X3, d4.view(), dx.view(), E.pack_convT4x4_dgrad(wt), zero_bias(c), relu=False)(st)
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx, c), x.grad) < 1e-2


def test_linknet_head_input_gradients(cuda):
    """finalconv3 (conv k2 p1, 32 -> 1), finalconv2 (valid conv3x3) and finaldeconv1 (ConvTranspose k3 s2, uncropped):
    SNB_CONV_2X2_ADJ (valid), conv3x3 in "full" mode, SNB_CONV_2X2_ADJ over the blocked gradient of odd size."""
    g = torch.Generator(device="cuda").manual_seed(33)
    n, h, w = 2, 24, 40
    st = N.stream_ptr()
    # conv k2 p1: input (2h-1), output 2h
    x3 = bf(torch.randn((n, 32, 2 * h - 1, 2 * w - 1), device="cuda", generator=g)).requires_grad_(True)
    w3 = bf(torch.randn((1, 32, 2, 2), device="cuda", generator=g) * 0.2)
    y3 = F.conv2d(x3, w3, None, padding=1)
    dy3 = bf(torch.randn(y3.shape, device="cuda", generator=g))
    y3.backward(dy3)
    dl = E.Slab(n, 2 * h, 2 * w, 32, "cuda")
    dl.t.zero_()
    dl.t[..., 0].copy_(dy3[:, 0].to(torch.bfloat16))
    dx3 = E.Slab(n, 2 * h - 1, 2 * w - 1, 32, "cuda")
    E.ConvOp(N.CONV_2X2_ADJ, dl.view(), dx3.view(), E.pack_conv2x2_dgrad(w3, 32, 32), zero_bias(32), relu=False, valid=1)(st)
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx3, 32), x3.grad) < 1e-2
    # valid conv3x3: input (2h+1), output (2h-1)
    x2 = bf(torch.randn((n, 32, 2 * h + 1, 2 * w + 1), device="cuda", generator=g)).requires_grad_(True)
    w2 = bf(torch.randn((32, 32, 3, 3), device="cuda", generator=g) * 0.1)
    y2 = F.conv2d(x2, w2)
    dy2 = bf(torch.randn(y2.shape, device="cuda", generator=g))
    y2.backward(dy2)
    dx2 = E.Slab(n, 2 * h + 1, 2 * w + 1, 32, "cuda")
    E.ConvOp(N.CONV_3X3, slab_from(dy2).view(0, 32), dx2.view(), E.pack_conv_dgrad(w2), zero_bias(32), relu=False, valid=2)(st)
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx2, 32), x2.grad) < 1e-2
    # ConvTranspose k3 s2 uncropped: input h, output 2h+1
    x1 = bf(torch.randn((n, 64, h, w), device="cuda", generator=g)).requires_grad_(True)
    w1 = bf(torch.randn((64, 32, 3, 3), device="cuda", generator=g) * 0.1)
    y1 = F.conv_transpose2d(x1, w1, None, stride=2)
    dy1 = bf(torch.randn(y1.shape, device="cuda", generator=g))
    y1.backward(dy1)
    d4 = E.Slab(n, h + 1, w + 1, 128, "cuda")
    dy1s = slab_from(dy1)
    N.check(N.lib().snb_space_to_depth2(N.c_vp(dy1s.t.data_ptr()), n, 2 * h + 1, 2 * w + 1, 32, 32, N.c_vp(d4.t.data_ptr()), 128, st))
    dx1 = E.Slab(n, h, w, 64, "cuda")
    E.ConvOp(N.CONV_2X2_ADJ, d4.view(), dx1.view(), E.pack_convT3x3_full_dgrad(w1), zero_bias(64), relu=False, valid=1)(st)
    torch.cuda.synchronize()
    assert rel_l2(nchw(dx1, 64), x1.grad) < 1e-2


def test_gather_segments_and_dropout_scale(cuda):
    """snb_gather_segments packs / unpacks through index maps; snb_scale_nc_nhwc applies a per-(image, channel) scale."""
    import ctypes

    g = torch.Generator(device="cuda").manual_seed(3)
    srcs = [torch.randn(n, device="cuda", generator=g) for n in (5000, 37, 2048)]
    idxs = [torch.randint(-1, s.numel(), (m,), device="cuda", generator=g, dtype=torch.int32) for s, m in zip(srcs, (3000, 1025, 7))]
    for bf16 in (False, True):
        dsts = [torch.full((i.numel(),), 9.0, device="cuda", dtype=torch.bfloat16 if bf16 else torch.float32) for i in idxs]
        table, first = [], 0
        for s, i, d in zip(srcs, idxs, dsts):
            table.append((s.data_ptr(), d.data_ptr(), i.data_ptr(), i.numel(), first))
            first += (i.numel() + 1023) // 1024
        segs = torch.tensor(table, dtype=torch.int64, device="cuda")
        N.check(N.lib().snb_gather_segments(N.ptr(segs), len(table), first, 1 if bf16 else 0, N.stream_ptr()))
        torch.cuda.synchronize()
        for s, i, d in zip(srcs, idxs, dsts):
            want = torch.where(i >= 0, s[i.clamp(min=0).long()], torch.zeros((), device="cuda"))
            assert torch.equal(d, want.to(d.dtype))
    n, h, w, c = 3, 9, 11, 64
    x = E.Slab(n, h, w, c, "cuda")
    x.t.copy_(torch.randn(x.t.shape, device="cuda", generator=g).to(torch.bfloat16))
    sc = (torch.rand((n, c), device="cuda", generator=g) > 0.5).float() * 2.0
    o = E.Slab(n, h, w, c, "cuda")
    N.check(N.lib().snb_scale_nc_nhwc(N.c_vp(x.t.data_ptr()), n, h * w, c, c, N.ptr(sc), N.c_vp(o.t.data_ptr()), c, N.stream_ptr()))
    assert torch.equal(o.t, (x.t.float() * sc.view(n, 1, 1, c)).to(torch.bfloat16))
