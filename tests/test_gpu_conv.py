"""-m gpu: the tcgen05 implicit-GEMM convolution through the C ABI against torch fp32 on bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

import snb_b200  # noqa: F401
from snb_b200 import _native as N
from snb_b200 import engine as E

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 2, 3, 4], ids=["tap", "halo", "halo+bres", "default", "pairs"], autouse=True)
def conv_mode(request, monkeypatch):
    """Every test runs against all main-loop variants, the default (3: fused ConvTranspose phases, wave-aware N tile)
    included (SNB_CONV_MODE is read at snb_conv_create)."""
    monkeypatch.setenv("SNB_CONV_MODE", str(request.param))
    return request.param


def bf(t):
    return t.to(torch.bfloat16).float()


def rand_slab(n, h, w, c, gen):
    s = E.Slab(n, h, w, c, "cuda")
    s.t.copy_(torch.randn((n, h, w, c), device="cuda", generator=gen).to(torch.bfloat16))
    return s


def nchw(view):
    return view.torch().float().permute(0, 3, 1, 2).contiguous()


def check(got, want, tol=2e-2):
    err = (got - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err <= tol * scale, "max abs err %g vs scale %g" % (err, scale)


@pytest.mark.parametrize("n,h,w,cin,cout,relu", [
    (1, 8, 16, 64, 64, True),        # exactly one M tile, one K chunk per tap
    (2, 32, 32, 64, 128, True),
    (1, 16, 16, 128, 256, False),
    (1, 24, 40, 256, 512, True),     # two N tiles, ragged spatial tiles (40 = 2.5 x 16)
    (3, 16, 32, 96, 32, True),       # BK = 32 path (SW64), BN = 32 store path
    (1, 64, 64, 192, 128, True),
    (1, 7, 7, 64, 64, True),         # smaller than one tile: TMA OOB fill + clipped store
    (13, 32, 32, 512, 512, True),    # bench.py's deep layers: 208 tiles of N=256 on 148 SMs -> the default mode narrows N to 128
    (13, 16, 16, 512, 512, True),
])
def test_conv3x3(cuda, n, h, w, cin, cout, relu):
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout)
    src = rand_slab(n, h, w, cin, g)
    dst = E.Slab(n, h, w, cout, "cuda")
    dst.t.fill_(float("nan"))
    wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    op = E.ConvOp(N.CONV_3X3, src.view(), dst.view(), E.pack_conv3x3(wt), bias, relu=relu)
    op(N.stream_ptr())
    torch.cuda.synchronize()
    want = F.conv2d(nchw(src.view()), bf(wt), bias, padding=1)
    if relu:
        want = F.relu(want)
    check(nchw(dst.view()), want)
    assert op.flops == 2.0 * n * h * w * cin * cout * 9


def test_conv_into_concat_slab_and_from_slab_slice(cuda):
    """Producer writes at a channel offset of a wider slab; consumer reads a channel range of a slab."""
    g = torch.Generator(device="cuda").manual_seed(3)
    n, h, w = 2, 16, 32
    src = rand_slab(n, h, w, 192, g)
    dst = E.Slab(n, h, w, 96, "cuda")
    dst.t.zero_()
    wt = torch.randn((64, 128, 3, 3), device="cuda", generator=g) * 0.03
    bias = torch.randn(64, device="cuda", generator=g)
    op = E.ConvOp(N.CONV_3X3, src.view(64, 128), dst.view(32, 64), E.pack_conv3x3(wt), bias)
    op(N.stream_ptr())
    torch.cuda.synchronize()
    want = F.relu(F.conv2d(nchw(src.view(64, 128)), bf(wt), bias, padding=1))
    check(nchw(dst.view(32, 64)), want)
    assert torch.count_nonzero(dst.t[..., :32]) == 0                      # neighbours in the slab untouched


def test_first_layer_as_patch_gemm(cuda):
    g = torch.Generator(device="cuda").manual_seed(4)
    n, h, w = 2, 32, 48
    x = torch.randn((n, 3, h, w), device="cuda", generator=g)
    rows = E.Slab(n, h, w, 32, "cuda")
    N.check(N.lib().snb_nchw_f32_to_patch32(N.ptr(x), n, 3, h, w, N.c_vp(rows.t.data_ptr()), 0, N.stream_ptr()))
    wt = torch.randn((64, 3, 3, 3), device="cuda", generator=g) * 0.2
    bias = torch.randn(64, device="cuda", generator=g)
    dst = E.Slab(n, h, w, 64, "cuda")
    op = E.ConvOp(N.CONV_1X1, rows.view(), dst.view(), E.pack_first_conv3x3(wt), bias)
    op(N.stream_ptr())
    torch.cuda.synchronize()
    check(nchw(dst.view()), F.relu(F.conv2d(bf(x), bf(wt), bias, padding=1)))


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 8, 16, 64, 64), (2, 16, 16, 512, 256), (1, 32, 32, 128, 32),
                                            (1, 12, 20, 256, 64)])
def test_conv_transpose_4x4_s2(cuda, n, h, w, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    src = rand_slab(n, h, w, cin, g)
    dst = E.Slab(n, 2 * h, 2 * w, cout + 32, "cuda")                      # written at channel offset 0 of a wider slab
    dst.t.zero_()
    wt = torch.randn((cin, cout, 4, 4), device="cuda", generator=g) * (2.0 / (4 * cin)) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    op = E.ConvOp(N.CONVT_4X4_S2, src.view(), dst.view(0, cout), E.pack_convT4x4(wt), bias)
    op(N.stream_ptr())
    torch.cuda.synchronize()
    want = F.relu(F.conv_transpose2d(nchw(src.view()), bf(wt), bias, stride=2, padding=1))
    check(nchw(dst.view(0, cout)), want)
    assert torch.count_nonzero(dst.t[..., cout:]) == 0
    assert op.flops == 2.0 * n * h * w * cin * cout * 16


@pytest.mark.parametrize("sigmoid", [False, True])
def test_fused_head(cuda, sigmoid):
    g = torch.Generator(device="cuda").manual_seed(9)
    n, h, w = 2, 24, 32
    src = rand_slab(n, h, w, 96, g)
    wt = torch.randn((32, 96, 3, 3), device="cuda", generator=g) * 0.05
    bias = torch.randn(32, device="cuda", generator=g) * 0.1
    hw = torch.randn(32, device="cuda", generator=g) * 0.3
    out = torch.full((n, h, w), float("nan"), device="cuda")
    op = E.ConvOp(N.CONV_3X3, src.view(), None, E.pack_conv3x3(wt), bias, head=(hw, 0.125, sigmoid, out))
    op(N.stream_ptr())
    torch.cuda.synchronize()
    feat = F.relu(F.conv2d(nchw(src.view()), bf(wt), bias, padding=1))
    want = (feat * hw.view(1, 32, 1, 1)).sum(1) + 0.125
    if sigmoid:
        want = torch.sigmoid(want)
    assert (out - want).abs().max().item() < 2e-3


def test_maxpool_and_exit_layout(cuda):
    g = torch.Generator(device="cuda").manual_seed(5)
    src = rand_slab(2, 16, 24, 96, g)
    dst = E.Slab(2, 8, 12, 64, "cuda")
    E.PoolOp(src.view(32, 64), dst.view())(N.stream_ptr())
    want = F.max_pool2d(nchw(src.view(32, 64)), 2, 2)
    assert torch.equal(nchw(dst.view()), want)
    out = torch.empty((2, 64, 8, 12), device="cuda")
    N.check(N.lib().snb_nhwc_bf16_to_nchw_f32(N.c_vp(dst.t.data_ptr()), 2, 8, 12, 64, 64, N.ptr(out), N.stream_ptr()))
    assert torch.equal(out, want)


def test_bad_descriptors_raise(cuda):
    s = E.Slab(1, 8, 16, 48, "cuda")
    d = E.Slab(1, 8, 16, 64, "cuda")
    w = torch.zeros((9, 64, 48), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError):
        E.ConvOp(N.CONV_3X3, s.view(), d.view(), w, torch.zeros(64, device="cuda"))   # cin not a multiple of 32


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 64, 64), (1, 16, 24, 128, 256), (1, 48, 40, 64, 128),
                                            (1, 32, 16, 256, 512)])
def test_fused_maxpool_epilogue(cuda, conv_mode, n, h, w, cin, cout):
    """The stage-final conv writes conv output and its 2x2 max-pool from the same epilogue (lib/models/unet16.py:64)."""
    if conv_mode == 0:
        pytest.skip("tap mode keeps the separate pooling kernel")
    g = torch.Generator(device="cuda").manual_seed(cin + 3 * cout)
    src = rand_slab(n, h, w, cin, g)
    dst = E.Slab(n, h, w, cout + 32, "cuda")
    pooled = E.Slab(n, h // 2, w // 2, cout, "cuda")
    dst.t.zero_()
    pooled.t.fill_(float("nan"))
    wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    op = E.ConvOp(N.CONV_3X3, src.view(), dst.view(32, cout), E.pack_conv3x3(wt), bias, pool_dst=pooled.view())
    op(N.stream_ptr())
    torch.cuda.synchronize()
    want = F.relu(F.conv2d(nchw(src.view()), bf(wt), bias, padding=1))
    check(nchw(dst.view(32, cout)), want)
    # the pooled tensor is exactly the max-pool of what the kernel itself stored
    assert torch.equal(nchw(pooled.view()), F.max_pool2d(nchw(dst.view(32, cout)), 2, 2))


# ------------------------------------------------------------------------------------------------ TF32 ("fp32 mode")
def rand_slab_f32(n, h, w, c, gen):
    s = E.Slab(n, h, w, c, "cuda", torch.float32)
    s.t.copy_(E.round_tf32(torch.randn((n, h, w, c), device="cuda", generator=gen)))
    return s


def nchw32(view):
    return view.torch().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("kind,n,h,w,cin,cout", [
    ("c3", 1, 16, 16, 32, 32), ("c3", 2, 32, 24, 64, 128), ("c3", 1, 16, 32, 48, 64), ("c3", 1, 24, 40, 128, 512),
    ("ct", 1, 8, 16, 64, 64), ("ct", 2, 16, 16, 128, 32), ("ct", 1, 12, 20, 256, 256)])
def test_tf32_convolutions(cuda, conv_mode, kind, n, h, w, cin, cout):
    """fp32 storage + kind::tf32: operands are TF32-representable, so the only difference to torch fp32 is the
    accumulation order (well inside the 1e-4 budget of the fp32 mode)."""
    if conv_mode == 4:
        pytest.skip("CTA pairs are built for bf16 only")
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    src = rand_slab_f32(n, h, w, cin, g)
    bias = torch.randn(cout, device="cuda", generator=g)
    if kind == "c3":
        dst = E.Slab(n, h, w, cout + 16, "cuda", torch.float32)
        dst.t.zero_()
        pooled = E.Slab(n, h // 2, w // 2, cout, "cuda", torch.float32)
        wt = E.round_tf32(torch.randn((cout, cin, 3, 3), device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5)
        fuse = conv_mode != 0
        op = E.ConvOp(N.CONV_3X3, src.view(), dst.view(16, cout), E.pack_conv3x3(wt, torch.float32), bias,
                      pool_dst=pooled.view() if fuse else None)
        op(N.stream_ptr())
        want = F.relu(F.conv2d(nchw32(src.view()), wt, bias, padding=1))
        got = nchw32(dst.view(16, cout))
        if fuse:
            assert torch.equal(nchw32(pooled.view()), F.max_pool2d(got, 2, 2))
        assert torch.count_nonzero(dst.t[..., :16]) == 0
    else:
        dst = E.Slab(n, 2 * h, 2 * w, cout, "cuda", torch.float32)
        wt = E.round_tf32(torch.randn((cin, cout, 4, 4), device="cuda", generator=g) * (2.0 / (4 * cin)) ** 0.5)
        op = E.ConvOp(N.CONVT_4X4_S2, src.view(), dst.view(), E.pack_convT4x4(wt, torch.float32), bias)
        op(N.stream_ptr())
        want = F.relu(F.conv_transpose2d(nchw32(src.view()), wt, bias, stride=2, padding=1))
        got = nchw32(dst.view())
    torch.cuda.synchronize()
    # stored outputs are rounded to TF32 (2^-11 relative); the accumulation itself agrees to ~1e-6
    assert (got - E.round_tf32(want)).abs().max().item() <= 1.5e-3 * want.abs().max().item()
    assert (got - want).abs().max().item() <= 6e-4 * want.abs().max().item()


def test_tf32_first_layer_and_head(cuda, conv_mode):
    if conv_mode == 4:
        pytest.skip("CTA pairs are built for bf16 only")
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(21)
    n, h, w = 2, 32, 48
    x = torch.randn((n, 3, h, w), device="cuda", generator=g)
    rows = E.Slab(n, h, w, 32, "cuda", torch.float32)
    N.check(N.lib().snb_nchw_f32_to_patch32(N.ptr(x), n, 3, h, w, N.c_vp(rows.t.data_ptr()), 1, N.stream_ptr()))
    wt = E.round_tf32(torch.randn((64, 3, 3, 3), device="cuda", generator=g) * 0.2)
    bias = torch.randn(64, device="cuda", generator=g)
    dst = E.Slab(n, h, w, 64, "cuda", torch.float32)
    E.ConvOp(N.CONV_1X1, rows.view(), dst.view(), E.pack_first_conv3x3(wt, torch.float32), bias)(N.stream_ptr())
    want = F.relu(F.conv2d(E.round_tf32(x), wt, bias, padding=1))
    assert (nchw32(dst.view()) - want).abs().max().item() <= 6e-4 * want.abs().max().item()
    # fused 1x1 head on a 96 -> 32 conv
    src = rand_slab_f32(n, h, w, 96, g)
    w2 = E.round_tf32(torch.randn((32, 96, 3, 3), device="cuda", generator=g) * 0.05)
    b2 = torch.randn(32, device="cuda", generator=g) * 0.1
    hw = torch.randn(32, device="cuda", generator=g) * 0.3
    out = torch.empty((n, h, w), device="cuda")
    E.ConvOp(N.CONV_3X3, src.view(), None, E.pack_conv3x3(w2, torch.float32), b2, head=(hw, 0.125, True, out))(N.stream_ptr())
    feat = F.relu(F.conv2d(nchw32(src.view()), w2, b2, padding=1))
    want = torch.sigmoid((feat * hw.view(1, 32, 1, 1)).sum(1) + 0.125)
    assert (out - want).abs().max().item() < 2e-6


@pytest.mark.parametrize("n,h,w,cin", [(1, 16, 16, 64), (2, 24, 40, 96), (1, 7, 7, 160), (3, 32, 16, 448)])
def test_fused_preactivation_prologue(cuda, conv_mode, n, h, w, cin):
    """conv3x3(relu(x * scale + shift)) with the pre-activation applied to the operand tiles inside the kernel
    (FCDenseNet DenseLayer, lib/models/tiramisu.py:9-19); zero padding applies after the activation."""
    if conv_mode == 0:
        pytest.skip("tap mode has no prologue warps")
    g = torch.Generator(device="cuda").manual_seed(cin)
    src = rand_slab(n, h, w, cin + 32, g)                       # wider slab: the conv reads the first cin channels
    dst = E.Slab(n, h, w, 32, "cuda")
    wt = torch.randn((32, cin, 3, 3), device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(32, device="cuda", generator=g)
    scale = torch.rand(cin, device="cuda", generator=g) + 0.5
    shift = torch.randn(cin, device="cuda", generator=g) * 0.3 + 0.2      # positive shifts make relu(shift) != 0 at the border
    op = E.ConvOp(N.CONV_3X3, src.view(0, cin), dst.view(), E.pack_conv3x3(wt), bias, relu=False, pre=(scale, shift))
    op(N.stream_ptr())
    torch.cuda.synchronize()
    x = nchw(src.view(0, cin))
    z = bf(F.relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)))
    want = F.conv2d(z, bf(wt), bias, padding=1)
    check(nchw(dst.view()), want)


# ------------------------------------------------------------------------- LinkNet34 layer shapes (lib/models/linknet.py)
def _need_halo(conv_mode):
    if conv_mode == 0:
        pytest.skip("valid / widened-grid / residual convolutions exist in halo mode only")


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 16, 16, 64, 64), (2, 24, 40, 64, 128), (1, 8, 8, 256, 512), (1, 32, 32, 128, 256)])
def test_stride2_conv3x3_via_space_to_depth(cuda, conv_mode, n, h, w, cin, cout):
    """resnet34 down-sampling block: conv3x3 stride 2 padding 1 == 4-tap conv over the space-to-depth tensor; the
    stride-2 conv1x1 of the shortcut == conv1x1 over its first channel quarter."""
    _need_halo(conv_mode)
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    src = rand_slab(n, h, w, cin, g)
    x4 = E.Slab(n, h // 2, w // 2, 4 * cin, "cuda")
    N.check(N.lib().snb_space_to_depth2(N.c_vp(src.view().ptr), n, h, w, cin, cin, N.c_vp(x4.view().ptr), 4 * cin, N.stream_ptr()))
    want4 = F.pixel_unshuffle(nchw(src.view()), 2)                        # channel = c*4 + py*2 + px
    want4 = want4.view(n, cin, 4, h // 2, w // 2).transpose(1, 2).reshape(n, 4 * cin, h // 2, w // 2)
    assert torch.equal(nchw(x4.view()), want4)
    wt = torch.randn((cout, cin, 3, 3), device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    dst = E.Slab(n, h // 2, w // 2, cout, "cuda")
    dst.t.fill_(float("nan"))
    op = E.ConvOp(N.CONV_2X2, x4.view(), dst.view(), E.pack_conv3x3_s2(wt), bias, valid=True)
    op(N.stream_ptr())
    check(nchw(dst.view()), F.relu(F.conv2d(nchw(src.view()), bf(wt), bias, stride=2, padding=1)))
    w1 = torch.randn((cout, cin, 1, 1), device="cuda", generator=g) * (1.0 / cin) ** 0.5
    d1 = E.Slab(n, h // 2, w // 2, cout, "cuda")
    E.ConvOp(N.CONV_1X1, x4.view(0, cin), d1.view(), E.pack_conv1x1(w1), bias, relu=False)(N.stream_ptr())
    check(nchw(d1.view()), F.conv2d(nchw(src.view()), bf(w1), bias, stride=2))


@pytest.mark.parametrize("n,h,w,c,after", [(1, 16, 16, 64, False), (2, 24, 40, 128, False), (1, 8, 16, 512, False),
                                           (2, 16, 24, 256, True)])
def test_residual_and_leaky_epilogue(cuda, conv_mode, n, h, w, c, after):
    """BasicBlock tail relu(conv + identity) and LinkNet skip leaky(conv) + skip, the residual read in the epilogue."""
    _need_halo(conv_mode)
    g = torch.Generator(device="cuda").manual_seed(c + after)
    src, res = rand_slab(n, h, w, c, g), rand_slab(n, h, w, c + 32, g)
    wt = torch.randn((c, c, 3, 3), device="cuda", generator=g) * (2.0 / (9 * c)) ** 0.5
    bias = torch.randn(c, device="cuda", generator=g)
    dst = E.Slab(n, h, w, c, "cuda")
    slope = 0.01 if after else 0.0
    E.ConvOp(N.CONV_3X3, src.view(), dst.view(), E.pack_conv3x3(wt), bias, act_slope=slope, residual=res.view(32, c),
             res_after_act=after)(N.stream_ptr())
    y, r = F.conv2d(nchw(src.view()), bf(wt), bias, padding=1), nchw(res.view(32, c))
    want = F.leaky_relu(y, slope) + r if after else F.relu(y + r)
    check(nchw(dst.view()), want)
    # conv1x1 with the skip added after the leaky-ReLU (decoder conv3 + encoder feature)
    w1 = torch.randn((c, c, 1, 1), device="cuda", generator=g) * (1.0 / c) ** 0.5
    d1 = E.Slab(n, h, w, c, "cuda")
    E.ConvOp(N.CONV_1X1, src.view(), d1.view(), E.pack_conv1x1(w1), bias, act_slope=0.01, residual=res.view(0, c),
             res_after_act=True)(N.stream_ptr())
    check(nchw(d1.view()), F.leaky_relu(F.conv2d(nchw(src.view()), bf(w1), bias), 0.01) + nchw(res.view(0, c)))


@pytest.mark.parametrize("n,h,w", [(1, 16, 16), (2, 24, 40), (1, 33, 17)])
def test_linknet_head_layers(cuda, conv_mode, n, h, w):
    """finaldeconv1 (ConvTranspose k3 s2, uncropped 2h+1), finalconv2 (valid conv3x3), finalconv3 (conv k2 p1, h+1)."""
    _need_halo(conv_mode)
    g = torch.Generator(device="cuda").manual_seed(h * w)
    src = rand_slab(n, h, w, 64, g)
    wt = torch.randn((64, 32, 3, 3), device="cuda", generator=g) * 0.1
    bias = torch.randn(32, device="cuda", generator=g)
    f1 = E.Slab(n, 2 * h + 1, 2 * w + 1, 32, "cuda")
    f1.t.fill_(float("nan"))
    op = E.ConvOp(N.CONVT_3X3_S2_FULL, src.view(), f1.view(), E.pack_convT3x3(wt, 64, 32), bias, act_slope=0.01)
    op(N.stream_ptr())
    check(nchw(f1.view()), F.leaky_relu(F.conv_transpose2d(nchw(src.view()), bf(wt), bias, stride=2), 0.01))
    assert op.flops == 2.0 * n * h * w * 64 * 32 * 9
    w2 = torch.randn((32, 32, 3, 3), device="cuda", generator=g) * 0.08
    f2 = E.Slab(n, 2 * h - 1, 2 * w - 1, 32, "cuda")
    f2.t.fill_(float("nan"))
    E.ConvOp(N.CONV_3X3, f1.view(), f2.view(), E.pack_conv3x3(w2), bias, act_slope=0.01, valid=True)(N.stream_ptr())
    check(nchw(f2.view()), F.leaky_relu(F.conv2d(nchw(f1.view()), bf(w2)[:, :, :, :], bias), 0.01))
    w3 = torch.randn((32, 32, 2, 2), device="cuda", generator=g) * 0.1
    f3 = E.Slab(n, 2 * h, 2 * w, 32, "cuda")
    f3.t.fill_(float("nan"))
    E.ConvOp(N.CONV_2X2, f2.view(), f3.view(), E.pack_conv2x2(w3), bias, relu=False)(N.stream_ptr())
    check(nchw(f3.view()), F.conv2d(nchw(f2.view()), bf(w3), bias, padding=1))
    # the same layer through the fused head (channel 0 picked), as LinkNet34Plan runs finalconv3
    out = torch.full((n, 2 * h, 2 * w), float("nan"), device="cuda")
    pick = torch.zeros(32, device="cuda")
    pick[0] = 1.0
    E.ConvOp(N.CONV_2X2, f2.view(), None, E.pack_conv2x2(w3), bias, relu=False, head=(pick, 0.0, False, out))(N.stream_ptr())
    want = F.conv2d(nchw(f2.view()), bf(w3), bias, padding=1)[:, 0]
    assert (out - want).abs().max().item() < 2e-3 * max(1.0, want.abs().max().item())


def test_resnet_stem_helpers(cuda):
    """7x7/s2 stem rows (im2col for a K=160 GEMM) and MaxPool2d(3, 2, 1) against torch."""
    g = torch.Generator(device="cuda").manual_seed(77)
    n, h, w = 2, 32, 48
    x = torch.randn((n, 3, h, w), device="cuda", generator=g)
    rows = E.Slab(n, h // 2, w // 2, 160, "cuda")
    N.check(N.lib().snb_stem7x7_rows(N.ptr(x), n, 3, h, w, N.c_vp(rows.t.data_ptr()), 160, N.stream_ptr()))
    wt = torch.randn((64, 3, 7, 7), device="cuda", generator=g) * 0.1
    bias = torch.randn(64, device="cuda", generator=g)
    stem = E.Slab(n, h // 2, w // 2, 64, "cuda")
    E.ConvOp(N.CONV_1X1, rows.view(), stem.view(), E.pack_stem7x7(wt, 160), bias)(N.stream_ptr())
    check(nchw(stem.view()), F.relu(F.conv2d(bf(x), bf(wt), bias, stride=2, padding=3)))
    pooled = E.Slab(n, h // 4, w // 4, 64, "cuda")
    N.check(N.lib().snb_maxpool3x3s2(N.c_vp(stem.view().ptr), n, h // 2, w // 2, 64, 64, N.c_vp(pooled.view().ptr), 64,
                                     N.stream_ptr()))
    assert torch.equal(nchw(pooled.view()), F.max_pool2d(nchw(stem.view()), 3, 2, 1))


@pytest.mark.parametrize("n,h,w,cin,pre", [
    (1, 14, 6, 32, False),           # exactly one tile, BK = 32
    (2, 28, 24, 64, True),           # 2 x 4 tiles, BK = 64
    (1, 224, 224, 96, True),         # FCDenseNet67 first dense block shapes, BK = 32
    (3, 20, 13, 288, True),          # ragged tiles, resident weights (4.5 chunks -> BK = 32, 9 chunks)
    (1, 7, 7, 448, True),            # smaller than a tile, streamed weights (do not fit next to the pipeline)
    (2, 56, 56, 640, False),         # streamed weights, many K chunks
])
def test_scatter_conv3x3_cout16(cuda, conv_mode, n, h, w, cin, pre):
    """FCDenseNet growth-rate layer relu(bn(x)) -> conv3x3(cin -> 16) as one N = 144 GEMM per tile
    (csrc/conv_scatter.cu) against torch; written into 16 channels in the middle of a wider slab."""
    if conv_mode != 1:
        pytest.skip("independent of SNB_CONV_MODE: run once")
    g = torch.Generator(device="cuda").manual_seed(cin + h)
    src = rand_slab(n, h, w, cin + 32, g)
    dst = E.Slab(n, h, w, 64, "cuda")
    dst.t.fill_(7.0)
    wt = torch.randn((16, cin, 3, 3), device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    bias = torch.randn(16, device="cuda", generator=g)
    scale = torch.rand(cin, device="cuda", generator=g) + 0.5
    shift = torch.randn(cin, device="cuda", generator=g) * 0.3 + 0.2
    op = E.ScatterConvOp(src.view(0, cin), dst.view(32, 16), E.pack_conv3x3_scatter(wt), bias,
                         pre=(scale, shift) if pre else None)
    op(N.stream_ptr())
    op(N.stream_ptr())                                                  # relaunch: barriers / TMEM are per launch
    torch.cuda.synchronize()
    x = nchw(src.view(0, cin))
    if pre:
        x = bf(F.relu(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)))
    want = F.conv2d(x, bf(wt), bias, padding=1)
    check(nchw(dst.view(32, 16)), want)
    assert torch.all(dst.t[..., :32] == 7.0) and torch.all(dst.t[..., 48:] == 7.0)      # neighbours untouched
    assert op.flops == 2.0 * n * h * w * cin * 16 * 9


@pytest.mark.parametrize("n,h,w,cout,relu", [(2, 64, 96, 64, True), (1, 13, 24, 64, False), (3, 40, 8, 32, True), (1, 512, 512, 64, True)])
def test_first_layer_from_packed_3_channel_tile(cuda, n, h, w, cout, relu):
    """SNB_CONV_FIRST_3X3: conv3x3 (Cin = 3, padding 1) whose K = 32 operand rows are built in shared memory from the packed
    NHWC bf16 tile, against torch on the same bf16-rounded input and weights; ragged heights, one- and many-tile widths."""
    g = torch.Generator(device="cuda").manual_seed(h + w)
    x = torch.randn((n, 3, h, w), device="cuda", generator=g)
    src = E.Slab(n, h, w, 3, "cuda")
    N.check(N.lib().snb_nchw_f32_to_nhwc3(N.ptr(x.contiguous()), n, h, w, N.c_vp(src.t.data_ptr()), N.stream_ptr()))
    assert torch.equal(src.t, x.permute(0, 2, 3, 1).to(torch.bfloat16))
    dst = E.Slab(n, h, w, cout + 32, "cuda")
    dst.t.fill_(7.0)
    wt = torch.randn((cout, 3, 3, 3), device="cuda", generator=g) * 0.3
    bias = torch.randn(cout, device="cuda", generator=g)
    op = E.ConvOp(N.CONV_FIRST_3X3, src.view(), dst.view(32, cout), E.pack_first_conv3x3(wt), bias, relu=relu)
    op(N.stream_ptr())
    torch.cuda.synchronize()
    want = F.conv2d(bf(x), bf(wt), bias, padding=1)
    if relu:
        want = F.relu(want)
    check(nchw(dst.view(32, cout)), want)
    assert torch.all(dst.t[..., :32] == 7.0)
    assert op.flops == 2.0 * n * h * w * 27 * cout
