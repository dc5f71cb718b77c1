"""Static execution plans for the encoder-decoder bodies (TernausNet family) on top of the C ABI.

A plan is built once per (weights, batch, height, width): NHWC bf16 channel slabs are allocated, every producer
is pointed at its channel offset inside the consumer's concat slab (so `torch.cat` of
lib/models/unet16.py:122-127 never runs), weights are packed K-major bf16 per tap, and each layer becomes one
`snb_conv` handle (TMA descriptors baked in).  `run()` only enqueues kernels on the current stream: no
allocation, no synchronisation, CUDA-graph capturable.
"""
import ctypes
import os

import torch

from . import _native as N


def head_op(plan):
    """The op of an inference plan that writes plan.out (the conv with the fused 1x1 head), or None."""
    for op in reversed(getattr(plan, "ops", [])):
        if getattr(op, "has_head", False):
            return op
    return None


# ------------------------------------------------------------------------------------------- precision
PRECISIONS = {"bf16": torch.bfloat16, "tf32": torch.float32}


def round_tf32(t):
    """fp32 -> nearest TF32 value (10-bit mantissa, ties away from zero like cvt.rna.tf32.f32), still stored as fp32."""
    bits = t.detach().float().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


def _f(t):
    """detached float32 copy of a parameter, or the float64 tensor itself (index maps are pushed through the pack functions
    as float64 so that every element keeps its identity)."""
    t = t.detach()
    return t if t.dtype == torch.float64 else t.float()


def to_storage(t, dtype):
    """Weights in their device storage type: bf16, or fp32 rounded to TF32 for the TF32 ("fp32 mode") path; float64 keeps the
    values exactly (index maps and expected gradients in the packed layout)."""
    if dtype == torch.float64:
        return t.detach().double()
    return round_tf32(t) if dtype == torch.float32 else t.detach().to(torch.bfloat16)


# ------------------------------------------------------------------------------------------- weight packing
def pack_conv3x3(weight, dtype=torch.bfloat16):
    """nn.Conv2d weight [Cout, Cin, 3, 3] -> [9][Cout][Cin] in `dtype`, tap = ky*3 + kx."""
    cout, cin = weight.shape[:2]
    return to_storage(weight.detach().permute(2, 3, 0, 1).reshape(9, cout, cin), dtype).contiguous()


def pack_conv1x1(weight, dtype=torch.bfloat16):
    cout, cin = weight.shape[:2]
    return to_storage(weight.detach().reshape(1, cout, cin), dtype).contiguous()


def pack_first_conv3x3(weight, dtype=torch.bfloat16):
    """First layer [Cout, C<=3, 3, 3] -> [1][Cout][32] matching the PATCH32 rows: k = (ky*3+kx)*C + c."""
    cout, cin = weight.shape[:2]
    w = weight.detach().permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    out = torch.zeros((1, cout, 32), dtype=dtype, device=weight.device)
    out[0, :, :9 * cin] = to_storage(w, dtype)
    return out.contiguous()


# out[2y+py] = sum in[y+dy] * W[ky]:  py=0 -> (dy=0, ky=1), (dy=-1, ky=3);  py=1 -> (dy=+1, ky=0), (dy=0, ky=2)
_CONVT_K = ((1, 3), (0, 2))


def pack_convT4x4(weight, dtype=torch.bfloat16):
    """nn.ConvTranspose2d(k=4, s=2, p=1) weight [Cin, Cout, 4, 4] -> [4 phases * 4 taps][Cout][Cin] in `dtype`."""
    cin, cout = weight.shape[:2]
    w = weight.detach()
    taps = []
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    taps.append(w[:, :, _CONVT_K[py][ty], _CONVT_K[px][tx]].t())
    return to_storage(torch.stack(taps), dtype).contiguous()


# --------------------------------------------------------------------------------------------------- slabs
class Slab:
    """NHWC buffer (bf16, or fp32 in TF32 mode); `view(c0, c)` names a channel range (a concat slot)."""

    def __init__(self, n, h, w, c, device, dtype=torch.bfloat16):
        self.n, self.h, self.w, self.c = n, h, w, c
        self.t = torch.empty((n, h, w, c), dtype=dtype, device=device)

    def view(self, c0=0, c=None):
        return SlabView(self, c0, self.c - c0 if c is None else c)


class SlabView:
    def __init__(self, slab, c0, c):
        assert 0 <= c0 and c0 + c <= slab.c
        self.slab, self.c0, self.c = slab, c0, c

    @property
    def ptr(self):
        return self.slab.t.data_ptr() + self.slab.t.element_size() * self.c0

    @property
    def cstride(self):
        return self.slab.c

    def torch(self):
        return self.slab.t[..., self.c0:self.c0 + self.c]


class ConvOp:
    """One snb_conv handle; keeps every tensor it points at alive."""

    def __init__(self, kind, src, dst, weight, bias, relu=True, head=None, pool_dst=None, upsample2x=False, pre=None,
                 act_slope=0.0, residual=None, res_after_act=False, valid=False, real=None):
        self.keep = (src, dst, weight, bias, head, pool_dst, pre, residual)
        self.has_head = head is not None
        d = N.ConvDesc()
        d.kind = kind
        d.dtype = N.CONV_TF32 if src.slab.t.dtype == torch.float32 else N.CONV_BF16
        if weight.dtype != src.slab.t.dtype:
            raise ValueError("weights (%s) and activations (%s) must share the storage type" % (weight.dtype, src.slab.t.dtype))
        d.relu = 1 if relu else 0
        d.n, d.h, d.w = src.slab.n, src.slab.h, src.slab.w
        d.cin, d.in_cstride = src.c, src.cstride
        d.d_in = src.ptr
        d.d_weight = weight.data_ptr()
        d.d_bias = bias.data_ptr()
        if head is None:
            d.cout, d.out_cstride = dst.c, dst.cstride
            d.d_out = dst.ptr
        else:
            head_w, head_b, head_sigmoid, head_out = head
            d.cout, d.out_cstride = 32, 32
            d.d_out = None
            d.d_head_w = head_w.data_ptr()
            d.head_b = float(head_b)
            d.head_sigmoid = 1 if head_sigmoid else 0
            d.d_head_out = head_out.data_ptr()
        d.out_upsample2x = 1 if upsample2x else 0
        d.act_slope = float(act_slope)
        d.valid = int(valid)          # 0 / False = padded, 1 / True = valid, 2 = full (conv3x3)
        if residual is not None:
            d.d_residual, d.res_cstride = residual.ptr, residual.cstride
            d.res_after_act = 1 if res_after_act else 0
        if pre is not None:
            d.d_pre_scale, d.d_pre_shift = pre[0].data_ptr(), pre[1].data_ptr()
        if pool_dst is not None:
            d.d_pool_out = pool_dst.ptr
            d.pool_cstride = pool_dst.cstride
        if kind == N.CONV_FIRST_3X3:
            if weight.shape[1] != d.cout or weight.shape[2] != 32 or src.c != 3 or src.cstride != 3:
                raise ValueError("the first-layer kernel takes a packed 3-channel tile and a [1][cout][32] weight")
        elif weight.shape[1] != d.cout or weight.shape[2] != d.cin:
            raise ValueError("packed weight %s does not match cout=%d cin=%d" % (tuple(weight.shape), d.cout, d.cin))
        self.desc = (int(d.kind), int(d.h), int(d.w), int(d.cin), int(d.cout))
        self._h = ctypes.c_void_p()
        N.check(N.lib().snb_conv_create(ctypes.byref(d), ctypes.byref(self._h)))
        self.flops = N.lib().snb_conv_flops(self._h)
        if real is not None:
            # algorithmic FLOPs: zero-padded channels do not count; real = (K actually contracted, real Cout)
            self.flops *= (real[0] / float(d.cin)) * (real[1] / float(d.cout))
        self.launches = 1

    def __call__(self, stream):
        N.check(N.lib().snb_conv_launch(self._h, stream))

    def set_head_out(self, ptr):
        """Fused-head convs only: the float [n][h][w] output goes to `ptr` (a device address) from the next launch on."""
        N.check(N.lib().snb_conv_set_head_out(self._h, N.c_vp(ptr)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().snb_conv_destroy(h)
            except Exception:
                pass


class WgradOp:
    """One snb_wgrad handle: dw[phase * taps + tap][co][ci] += sum_pixels dout[pixel][co] * src[pixel + tap][ci] on the
    tensor cores (csrc/conv_wgrad.cu), in the packed layout of the forward weights.  `src` / `dout` are the forward
    conv's input view and the gradient of its output; `dw` a float tensor [T][>= cout][>= cin] the caller zeroes."""

    def __init__(self, kind, src, dout, dw, valid=0, cin=None, cout=None):
        self.keep = (src, dout, dw)
        d = N.WgradDesc()
        d.kind, d.valid = kind, int(valid)
        d.n, d.h, d.w = src.slab.n, src.slab.h, src.slab.w
        d.cin, d.in_cstride = (src.c if cin is None else cin), src.cstride
        d.cout, d.dout_cstride = (dout.c if cout is None else cout), dout.cstride
        d.d_in, d.d_dout = src.ptr, dout.ptr
        if dw.dtype != torch.float32 or dw.dim() != 3 or not dw.is_contiguous():
            raise ValueError("dw must be a contiguous float tensor [taps][cout][cin]")
        d.d_dweight = dw.data_ptr()
        d.dw_cout, d.dw_cin = dw.shape[1], dw.shape[2]
        self._h = ctypes.c_void_p()
        N.check(N.lib().snb_wgrad_create(ctypes.byref(d), ctypes.byref(self._h)))
        self.flops = N.lib().snb_wgrad_flops(self._h)
        self.launches = 1

    def __call__(self, stream):
        N.check(N.lib().snb_wgrad_launch(self._h, stream))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().snb_wgrad_destroy(h)
            except Exception:
                pass


def pack_conv3x3_scatter(weight, dtype=torch.bfloat16):
    """[16, Cin, 3, 3] -> [144][Cin] for snb_conv_scatter: row tap * 16 + co = W[co][:, ky, kx], tap = ky * 3 + kx."""
    cout, cin = weight.shape[:2]
    assert cout == 16
    return to_storage(weight.detach().float().permute(2, 3, 0, 1).reshape(9 * cout, cin), dtype).contiguous()


class ScatterConvOp:
    """conv3x3 with 16 output channels as one N = 144 GEMM per tile (csrc/conv_scatter.cu); keeps its tensors alive."""

    def __init__(self, src, dst, weight, bias, pre=None):
        if dst.c != 16 or weight.shape != (144, src.c):
            raise ValueError("scatter conv needs 16 output channels and a [144][cin] weight")
        self.keep = (src, dst, weight, bias, pre)
        self.desc = (0, src.slab.h, src.slab.w, src.c, 16)
        self._h = ctypes.c_void_p()
        N.check(N.lib().snb_conv_scatter_create(
            N.c_vp(src.ptr), src.slab.n, src.slab.h, src.slab.w, src.c, src.cstride, N.ptr(weight), N.ptr(bias),
            N.ptr(pre[0]) if pre is not None else N.c_vp(0), N.ptr(pre[1]) if pre is not None else N.c_vp(0),
            N.c_vp(dst.ptr), dst.cstride, ctypes.byref(self._h)))
        self.flops = N.lib().snb_conv_scatter_flops(self._h)
        self.launches = 1

    def __call__(self, stream):
        N.check(N.lib().snb_conv_scatter_launch(self._h, stream))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                N.lib().snb_conv_scatter_destroy(h)
            except Exception:
                pass


def first_layer(plan, S, dtype, n, h, w, dst, weight, bias, relu=True, cout_real=None):
    """Input buffer + first conv3x3 (Cin = 3) of a plan.  bf16: the packed 3-channel tile (6 bytes per pixel) and the kernel
    that builds its operand rows in shared memory (SNB_CONV_FIRST_3X3); TF32 mode (and widths that are not multiples of 8):
    the PATCH32 rows in HBM and a K = 32 conv1x1 over them."""
    cout = weight.shape[0]
    real = (9 * weight.shape[1], cout if cout_real is None else cout_real)
    if dtype == torch.bfloat16 and w % 8 == 0 and weight.shape[1] == 3 and cout in (32, 64) and os.environ.get("SNB_FIRST_PATCH32", "0") != "1":
        plan.x_in3 = S(h, w, 3)
        plan.x_patch = None
        op = ConvOp(N.CONV_FIRST_3X3, plan.x_in3.view(), dst, pack_first_conv3x3(weight, dtype), bias, relu=relu)
        op.flops *= real[1] / float(cout)
        return op
    plan.x_in3 = None
    plan.x_patch = S(h, w, 32)
    return ConvOp(N.CONV_1X1, plan.x_patch.view(), dst, pack_first_conv3x3(weight, dtype), bias, relu=relu, real=real)


class PoolOp:
    def __init__(self, src, dst):
        self.keep = (src, dst)
        s = src.slab
        self.args = (N.c_vp(src.ptr), s.n, s.h, s.w, src.c, src.cstride, N.c_vp(dst.ptr), dst.cstride, s.t.element_size())
        self.flops = 0.0
        self.launches = 1

    def __call__(self, stream):
        N.check(N.lib().snb_maxpool2x2(*self.args, stream))


class VGGUNetPlan:
    """UNet11 / UNet16 forward (lib/models/unet11.py:106-122, unet16.py:113-131) as a list of kernel launches.

    enc: list of stages, each a list of (weight, bias) nn.Conv2d parameters (3x3, ReLU after each).
    decs: [center, dec5, dec4, dec3, dec2] as (conv_w, conv_b, convT_w, convT_b); dec1 = (w, b); final = (w, b).
    """

    def __init__(self, enc, decs, dec1, final, n, h, w, device, sigmoid, dtype=torch.bfloat16):
        if h % 32 or w % 32:
            raise ValueError("height and width must be multiples of 32 (five 2x2 poolings)")
        if final[0].shape[0] != 1 or final[0].shape[1] != 32 or dec1[0].shape[0] != 32:
            raise NotImplementedError("fused head expects num_classes == 1 and num_filters == 32")
        self.n, self.h, self.w = n, h, w
        self.device = device
        self.dtype = dtype
        self.ops = []
        # SNB_CONV_MODE=0 (tap-mode A/B runs) has no fused pooling; SNB_FUSE_POOL=0 keeps the separate kernel
        fuse_pool = os.environ.get("SNB_CONV_MODE", "3") != "0" and os.environ.get("SNB_FUSE_POOL", "1") != "0"
        f32 = lambda b: b.detach().float().contiguous()
        S = lambda hh, ww, c: Slab(n, hh, ww, c, device, dtype)

        # concat slabs: [decoder output | encoder skip], torch.cat([dec, conv], 1) order
        skip_c = [st[-1][0].shape[0] for st in enc]                # 64, 128, 256, 512, 512
        up_c = [d[2].shape[1] for d in decs]                       # center..dec2 ConvT output channels
        # decs[i] output feeds the slab of encoder stage 4-i (center -> conv5 slab, dec2 -> conv1 slab)
        slabs = []
        for s in range(5):
            hh, ww = h >> s, w >> s
            slabs.append(S(hh, ww, up_c[4 - s] + skip_c[s]))
        self.slabs = slabs

        cur = None
        for s, stage in enumerate(enc):
            hh, ww = h >> s, w >> s
            for li, (wt, bs) in enumerate(stage):
                cout = wt.shape[0]
                last = li == len(stage) - 1
                dst = slabs[s].view(up_c[4 - s], cout) if last else S(hh, ww, cout).view()
                # the stage's last conv also writes the 2x2 max-pooled tensor from its epilogue
                pooled = S(hh // 2, ww // 2, cout).view() if last else None
                fuse = pooled is not None and fuse_pool and not (s == 0 and li == 0)
                if s == 0 and li == 0:
                    self.ops.append(first_layer(self, S, dtype, n, h, w, dst, wt, f32(bs)))
                else:
                    self.ops.append(ConvOp(N.CONV_3X3, cur, dst, pack_conv3x3(wt, dtype), f32(bs),
                                           pool_dst=pooled if fuse else None))
                cur = dst
            if not fuse:
                self.ops.append(PoolOp(cur, pooled))
            cur = pooled

        # center consumes the pooled conv5; dec5..dec2 consume the concat slabs
        for i, (cw, cb, tw, tb) in enumerate(decs):
            s = 5 - i  # resolution level of this block's input
            hh, ww = h >> s, w >> s
            src = cur if i == 0 else slabs[s].view()
            mid = S(hh, ww, cw.shape[0]).view()
            self.ops.append(ConvOp(N.CONV_3X3, src, mid, pack_conv3x3(cw, dtype), f32(cb)))
            dst = slabs[s - 1].view(0, tw.shape[1])
            self.ops.append(ConvOp(N.CONVT_4X4_S2, mid, dst, pack_convT4x4(tw, dtype), f32(tb)))

        self.out = torch.empty((n, h, w), dtype=torch.float32, device=device)
        head = (final[0].detach().reshape(32).float().contiguous(), float(final[1].detach().reshape(-1)[0]), sigmoid,
                self.out)
        self.ops.append(ConvOp(N.CONV_3X3, slabs[0].view(), None, pack_conv3x3(dec1[0], dtype), f32(dec1[1]), head=head))
        self.ops[-1].flops += 2.0 * n * h * w * 32          # the fused 1x1 head (32 -> 1)
        self.flops = sum(op.flops for op in self.ops)
        self.launches = sum(op.launches for op in self.ops)

    def load_nchw(self, x):
        """float [n,3,h,w] CUDA tensor -> the first layer's input (packed 3-channel bf16 tile, or PATCH32 operand rows)."""
        x = x.contiguous()
        if self.x_in3 is not None:
            N.check(N.lib().snb_nchw_f32_to_nhwc3(N.ptr(x), x.shape[0], x.shape[2], x.shape[3],
                                                  N.c_vp(self.x_in3.t.data_ptr()), N.stream_ptr()))
            return
        N.check(N.lib().snb_nchw_f32_to_patch32(N.ptr(x), x.shape[0], x.shape[1], x.shape[2], x.shape[3],
                                                N.c_vp(self.x_patch.t.data_ptr()),
                                                1 if self.x_patch.t.dtype == torch.float32 else 0, N.stream_ptr()))

    def input_layout(self):
        """(snb_split_norm_u8 layout, destination pointer) of the tiled predictor's fused split"""
        if self.x_in3 is not None:
            return N.LAYOUT_NHWC3_BF16, self.x_in3.t.data_ptr()
        f32 = self.x_patch.t.dtype == torch.float32
        return (N.LAYOUT_PATCH32_F32 if f32 else N.LAYOUT_PATCH32), self.x_patch.t.data_ptr()

    def run(self):
        """Enqueue the whole forward on the current stream; result lands in self.out [n,h,w] float32."""
        st = N.stream_ptr()
        for op in self.ops:
            op(st)
        return self.out


def fold_bn(conv_w, conv_b, bn):
    """conv -> BatchNorm2d(eval) as one conv: w' = w * g / sqrt(var + eps), b' = (b - mean) * g / sqrt(var + eps) + beta."""
    if bn is None:
        return conv_w.detach().float(), conv_b.detach().float()
    gamma, beta, mean, var, eps = bn
    scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
    return conv_w.detach().float() * scale.view(-1, 1, 1, 1), (conv_b.detach().float() - mean.detach().float()) * scale + beta.detach().float()


class ZFUNetPlan:
    """ZF_UNET forward (lib/models/zf_unet.py:60-95) in eval mode: BatchNorm folded into the conv weights, Dropout2d
    off, MaxPool2d fused into the producing conv's epilogue, nn.Upsample(2) fused into the producing conv's store
    (the tile is TMA-stored to the 2x2 replicated positions of the consumer's concat slab), final 1x1 fused as head.

    blocks: 11 double-conv modules in forward order (conv_224, conv_112, conv_56, conv_28, conv_14, conv_7, up_conv_14,
    up_conv_28, up_conv_56, up_conv_112, up_conv_224), each ((w1, b1, bn1), (w2, b2, bn2)); final = (w, b).
    """

    def __init__(self, blocks, final, n, h, w, device, sigmoid, dtype=torch.bfloat16):
        self.dtype = dtype
        if h % 32 or w % 32:
            raise ValueError("height and width must be multiples of 32 (five 2x2 poolings)")
        filters = blocks[0][0][0].shape[0]
        if final[0].shape[0] != 1 or filters != 32:
            raise NotImplementedError("fused head expects num_classes == 1 and filters == 32")
        self.n, self.h, self.w, self.device = n, h, w, device
        self.ops = []
        S = lambda hh, ww, c: Slab(n, hh, ww, c, device, dtype)
        enc_c = [blocks[i][1][0].shape[0] for i in range(5)]          # 32, 64, 128, 256, 512
        up_c = [blocks[5][1][0].shape[0]] + [blocks[6 + i][1][0].shape[0] for i in range(4)]  # channels unpooled into level 4..0
        # concat slabs per level: [unpool(deeper) | encoder skip]  (torch.cat([self.unpool(..), conv_X], dim=1))
        slabs = [S(h >> l, w >> l, up_c[4 - l] + enc_c[l]) for l in range(5)]
        self.slabs = slabs

        def conv(src, dst, layer, first=False, **kw):
            wt, bs = fold_bn(*layer)
            if first:
                self.ops.append(first_layer(self, S, dtype, n, h, w, dst, wt, bs.contiguous()))
            else:
                self.ops.append(ConvOp(N.CONV_3X3, src, dst, pack_conv3x3(wt, dtype), bs.contiguous(), **kw))

        cur = None
        for l in range(5):                                             # encoder: double conv, skip into the slab, pool
            hh, ww = h >> l, w >> l
            l1, l2 = blocks[l]
            mid = S(hh, ww, enc_c[l]).view()
            conv(cur, mid, l1, first=(l == 0))
            pooled = S(hh // 2, ww // 2, enc_c[l]).view()
            conv(mid, slabs[l].view(up_c[4 - l], enc_c[l]), l2, pool_dst=pooled)
            cur = pooled
        for i in range(5):                                             # conv_7, up_conv_14 .. up_conv_112: store upsampled
            lvl = 5 - i                                                # resolution level of this block
            hh, ww = h >> lvl, w >> lvl
            l1, l2 = blocks[5 + i]
            src = cur if i == 0 else slabs[lvl].view()
            c_out = l1[0].shape[0]
            mid = S(hh, ww, c_out).view()
            conv(src, mid, l1)
            conv(mid, slabs[lvl - 1].view(0, c_out), l2, upsample2x=True)
        l1, l2 = blocks[10]                                            # up_conv_224 + conv_final
        mid = S(h, w, 32).view()
        conv(slabs[0].view(), mid, l1)
        self.out = torch.empty((n, h, w), dtype=torch.float32, device=device)
        wt, bs = fold_bn(*l2)
        head = (final[0].detach().reshape(32).float().contiguous(), float(final[1].detach().reshape(-1)[0]), sigmoid,
                self.out)
        self.ops.append(ConvOp(N.CONV_3X3, mid, None, pack_conv3x3(wt, dtype), bs.contiguous(), head=head))
        self.ops[-1].flops += 2.0 * n * h * w * 32          # the fused 1x1 head (32 -> 1)
        self.flops = sum(op.flops for op in self.ops)
        self.launches = sum(op.launches for op in self.ops)

    load_nchw = VGGUNetPlan.load_nchw
    input_layout = VGGUNetPlan.input_layout
    run = VGGUNetPlan.run


def fold_norm(conv_w, conv_b, norm):
    """conv -> BatchNorm2d(eval) or InPlaceABN(eval) as one conv.  norm = (gamma, beta, mean, var, eps, abn); the ABN scale
    is |gamma| + eps (mapillary's kernel)."""
    gamma, beta, mean, var, eps, abn = norm
    g = (gamma.detach().float().abs() + eps) if abn else gamma.detach().float()
    scale = g / torch.sqrt(var.detach().float() + eps)
    b0 = conv_b.detach().float() if conv_b is not None else torch.zeros_like(scale)
    return conv_w.detach().float() * scale.view(-1, 1, 1, 1), (b0 - mean.detach().float()) * scale + beta.detach().float()


class UNetPlan:
    """UNet / UNetABN forward (lib/models/unet.py:79-107, unet_abn.py:80-107) in eval mode: BatchNorm / InPlaceABN folded
    into the convolutions, MaxPool2d fused into the producing conv's epilogue, nn.Upsample(2, nearest) fused into the
    producing conv's store (the tile is TMA-stored to its 2x2 replicated positions inside the consumer's concat slab,
    torch.cat([skip, upsampled]) order), Dropout2d inactive, the 1x1 `outc` fused as head.

    blocks: the 9 double_conv modules in forward order (inc, down1-4, up1-4), each ((w1, b1, norm1), (w2, b2, norm2));
    final = (w, b) of outc; act_slope = 0 (ReLU) or 0.01 (InPlaceABN's leaky-ReLU)."""

    def __init__(self, blocks, final, n, h, w, device, sigmoid, dtype=torch.bfloat16, act_slope=0.0):
        self.dtype = dtype
        if h % 16 or w % 16:
            raise ValueError("height and width must be multiples of 16 (four 2x2 poolings)")
        nf = blocks[0][0][0].shape[0]
        if final[0].shape[0] != 1 or nf != 32:
            raise NotImplementedError("fused head expects n_classes == 1 and n_filters == 32")
        self.n, self.h, self.w, self.device = n, h, w, device
        self.ops = []
        S = lambda hh, ww, c: Slab(n, hh, ww, c, device, dtype)
        ec = [blocks[i][1][0].shape[0] for i in range(5)]                    # nf, 2nf, 4nf, 8nf, 8nf
        upc = [blocks[7][1][0].shape[0], blocks[6][1][0].shape[0], blocks[5][1][0].shape[0], ec[4]]   # into level 0..3
        slabs = [S(h >> l, w >> l, ec[l] + upc[l]) for l in range(4)]       # [skip | upsampled deeper features]
        self.slabs = slabs

        def conv(src, dst, layer, first=False, **kw):
            wt, bs = fold_norm(*layer)
            if first:
                self.ops.append(first_layer(self, S, dtype, n, h, w, dst, wt, bs.contiguous())
                                if act_slope == 0.0 else self._first_leaky(S, dtype, n, h, w, dst, wt, bs.contiguous(), act_slope))
            else:
                self.ops.append(ConvOp(N.CONV_3X3, src, dst, pack_conv3x3(wt, dtype), bs.contiguous(), act_slope=act_slope, **kw))

        cur = None
        for l in range(5):                                                   # inc, down1..down4
            hh, ww = h >> l, w >> l
            l1, l2 = blocks[l]
            mid = S(hh, ww, ec[l]).view()
            conv(cur, mid, l1, first=(l == 0))
            if l < 4:
                pooled = S(hh // 2, ww // 2, ec[l]).view()
                conv(mid, slabs[l].view(0, ec[l]), l2, pool_dst=pooled)
                cur = pooled
            else:                                                            # x5: upsampled straight into level 3's slab
                conv(mid, slabs[3].view(ec[3], ec[4]), l2, upsample2x=True)
        for i in range(3):                                                   # up1..up3 at levels 3, 2, 1
            lvl = 3 - i
            hh, ww = h >> lvl, w >> lvl
            l1, l2 = blocks[5 + i]
            c_out = l1[0].shape[0]
            mid = S(hh, ww, c_out).view()
            conv(slabs[lvl].view(), mid, l1)
            conv(mid, slabs[lvl - 1].view(ec[lvl - 1], c_out), l2, upsample2x=True)
        l1, l2 = blocks[8]                                                   # up4 + outc
        mid = S(h, w, nf).view()
        conv(slabs[0].view(), mid, l1)
        self.out = torch.empty((n, h, w), dtype=torch.float32, device=device)
        wt, bs = fold_norm(*l2)
        head = (final[0].detach().reshape(32).float().contiguous(), float(final[1].detach().reshape(-1)[0]), sigmoid, self.out)
        self.ops.append(ConvOp(N.CONV_3X3, mid, None, pack_conv3x3(wt, dtype), bs.contiguous(), head=head, act_slope=act_slope))
        self.ops[-1].flops += 2.0 * n * h * w * 32
        self.flops = sum(op.flops for op in self.ops)
        self.launches = sum(op.launches for op in self.ops)

    def _first_leaky(self, S, dtype, n, h, w, dst, wt, bias, slope):
        """first conv with a leaky activation (UNetABN): the 3-channel first-layer kernel only knows ReLU, so the PATCH32
        operand rows and the K = 32 conv1x1 (common epilogue with the activation slope) are used"""
        self.x_in3 = None
        self.x_patch = S(h, w, 32)
        return ConvOp(N.CONV_1X1, self.x_patch.view(), dst, pack_first_conv3x3(wt, dtype), bias, act_slope=slope,
                      real=(9 * wt.shape[1], wt.shape[0]))

    load_nchw = VGGUNetPlan.load_nchw
    input_layout = VGGUNetPlan.input_layout
    run = VGGUNetPlan.run


# ---------------------------------------------------------------------------------------------- FCDenseNet
# out[2y+py] = sum in[y+dy] * W[ky] for ConvTranspose2d(k=3, s=2, p=0) cropped to [0, 2h): py=0 -> (dy=0, ky=0),
# (dy=-1, ky=2); py=1 -> (dy=0, ky=1) and one unused slot (zero weights)
_CONVT3_K = ((0, 2), (1, None))


def pack_convT3x3(weight, cin_pad, cout_pad, dtype=torch.bfloat16):
    """nn.ConvTranspose2d(k=3, s=2, p=0) weight [Cin, Cout, 3, 3] -> bf16 [16 tap slots][cout_pad][cin_pad]."""
    cin, cout = weight.shape[:2]
    w = _f(weight)
    out = torch.zeros((16, cout_pad, cin_pad), dtype=w.dtype, device=weight.device)
    i = 0
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    ky, kx = _CONVT3_K[py][ty], _CONVT3_K[px][tx]
                    if ky is not None and kx is not None:
                        out[i, :cout, :cin] = w[:, :, ky, kx].t()
                    i += 1
    return to_storage(out, dtype).contiguous()


def pad_conv_weight(weight, cin_map, cin_pad, cout_pad):
    """[Cout, Cin, kh, kw] -> zero-padded [cout_pad, cin_pad, kh, kw] with input channel j moved to slab position
    cin_map[j] (the slab keeps concat members in a different order than torch.cat)."""
    cout, cin = weight.shape[:2]
    out = torch.zeros((cout_pad, cin_pad) + tuple(weight.shape[2:]), dtype=torch.float32, device=weight.device)
    out[:cout, cin_map] = weight.detach().float()
    return out


class BnReluOp:
    """relu(batch_norm_eval(x)) of a slab range into a scratch slab (snb_bn_relu_nhwc)."""

    def __init__(self, src, dst, scale, shift):
        self.keep = (src, dst, scale, shift)
        s = src.slab
        self.args = (N.c_vp(src.ptr), s.n, s.h, s.w, src.c, src.cstride, N.c_vp(scale.data_ptr()),
                     N.c_vp(shift.data_ptr()), N.c_vp(dst.ptr), dst.c, dst.cstride)
        self.flops = 0.0
        self.launches = 1

    def __call__(self, stream):
        N.check(N.lib().snb_bn_relu_nhwc(*self.args, stream))


def _pad32(c):
    return (c + 31) // 32 * 32


class FCDenseNetPlan:
    """FCDenseNet forward (lib/models/tiramisu.py:168-184) in eval mode.

    One slab per resolution level holds [dense-block input | down features | TransitionUp output | up features | 16
    zero channels]; every producer writes at its channel offset (no torch.cat), narrow outputs (growth rate 16, first
    conv 48, ConvT 80) are stored as 32-wide tiles whose zero tail lands on the slot of the layer that runs next.
    The per-consumer pre-activation BatchNorm+ReLU runs as an elementwise kernel into a scratch slab which the
    convolution then reads (fusing it into the operand path of the conv kernel is the next step).
    `spec` comes from snb_b200.lib.models.tiramisu.FCDenseNet._spec().
    """

    def __init__(self, spec, n, h, w, device, sigmoid):
        L = len(spec['down'])
        if h % (1 << L) or w % (1 << L):
            raise ValueError("height and width must be multiples of %d" % (1 << L))
        if spec['final'][0].shape[0] != 1:
            raise NotImplementedError("fused head expects n_classes == 1")
        self.n, self.h, self.w, self.device = n, h, w, device
        self.dtype = torch.bfloat16     # the pre-activation BN+ReLU path is bf16 only
        # SNB_FUSE_PRE=0 keeps the separate BN+ReLU pass (A/B runs); tap mode has no prologue warps
        fuse_pre = os.environ.get("SNB_FUSE_PRE", "1") != "0" and os.environ.get("SNB_CONV_MODE", "3") != "0"
        # SNB_SCATTER=0 keeps the tap-list kernel for the growth-rate layers (A/B runs)
        scatter = os.environ.get("SNB_SCATTER", "1") != "0"
        # measured on B200 (profiles/): both formulations are bound by the shared-memory operand bandwidth of the
        # tensor core (128 B/clk: the tap-list kernel re-reads the activation rows once per tap, the scatter kernel
        # pays for its fp32 partial planes once per tile), so the scatter kernel wins from ~160 input channels on
        scatter_min_cin = int(os.environ.get("SNB_SCATTER_MIN_CIN", "160"))
        self.ops = []
        g = spec['growth']
        dev = device
        S = lambda hh, ww, c: Slab(n, hh, ww, c, dev)
        f32 = lambda t: t.detach().float().contiguous()

        def bn_params(bn, cmap, cpad):
            gamma, beta, mean, var, eps = bn
            scale = gamma.detach().float() / torch.sqrt(var.detach().float() + eps)
            shift = beta.detach().float() - mean.detach().float() * scale
            sc = torch.zeros(cpad, dtype=torch.float32, device=dev)
            sh = torch.zeros(cpad, dtype=torch.float32, device=dev)
            sc[cmap], sh[cmap] = scale, shift
            return sc, sh

        def padded_bias(b, cpad):
            out = torch.zeros(cpad, dtype=torch.float32, device=dev)
            out[:b.numel()] = b.detach().float()
            return out

        # ---- slabs: level l holds c_in[l] + down + (up part on the way back)
        c_first = spec['first'][0].shape[0]
        c_in, c = [], c_first
        for layers in spec['down']:
            c_in.append(c)
            c += g * len(layers)
        c_bott_in = c
        n_bott = len(spec['bottleneck'])
        up_prev = [g * n_bott] + [g * len(b) for b in spec['up'][:-1]]      # channels arriving through each TransitionUp
        slabs = []
        for l in range(L):
            skip = c_in[l] + g * len(spec['down'][l])
            i_up = L - 1 - l                                                 # up block that runs at level l
            slabs.append(S(h >> l, w >> l, _pad32(skip + up_prev[i_up] + g * len(spec['up'][i_up])) + 32))
        bott = S(h >> L, w >> L, _pad32(c_bott_in + g * n_bott) + 32)
        for sl in slabs + [bott]:
            sl.t.zero_()                                                     # zero tails are read as K padding
        self.slabs, self.bott = slabs, bott
        max_c = max(sl.c for sl in slabs + [bott])
        scratch = [S(h >> l, w >> l, max_c) for l in range(L + 1)]          # bn-relu output per level
        for sc in scratch:
            sc.t.zero_()

        def dense_layer(slab, lvl, cin, cmap, layer, out_off):
            """BN -> ReLU -> conv3x3(cin -> g) reading slab[0:cin], writing slab[out_off : out_off + 32]."""
            bn, wt, bs = layer
            cpad = _pad32(cin)
            sc, sh = bn_params(bn, cmap, cpad)
            wp = pad_conv_weight(wt, cmap, cpad, 32)
            if scatter and g == 16 and fuse_pre and wt.shape[0] == 16 and out_off % 16 == 0 and cpad >= scatter_min_cin:
                # growth-rate conv as ONE N = 144 GEMM per tile (taps in the N dimension), BN+ReLU in the operand path
                w16 = pad_conv_weight(wt, cmap, cpad, 16)
                self.ops.append(ScatterConvOp(slab.view(0, cpad), slab.view(out_off, 16), pack_conv3x3_scatter(w16),
                                              padded_bias(bs, 16), pre=(sc, sh)))
                self.ops[-1].flops *= cin / float(cpad)
                return
            if fuse_pre:
                # BatchNorm+ReLU applied to the operand tiles inside the conv kernel: the slab is read once, in place
                self.ops.append(ConvOp(N.CONV_3X3, slab.view(0, cpad), slab.view(out_off, 32), pack_conv3x3(wp),
                                       padded_bias(bs, 32), relu=False, pre=(sc, sh), real=(cin, wt.shape[0])))
                return
            z = scratch[lvl].view(0, cpad)
            self.ops.append(BnReluOp(slab.view(0, cpad), z, sc, sh))
            self.ops.append(ConvOp(N.CONV_3X3, z, slab.view(out_off, 32), pack_conv3x3(wp), padded_bias(bs, 32), relu=False,
                                   real=(cin, wt.shape[0])))

        # ---- first conv: 3 -> 48 (no activation), stored 64 wide
        wf, bf = spec['first']
        cf_pad = _pad32(c_first) if c_first % 64 == 0 else (c_first + 63) // 64 * 64
        wfp = torch.zeros((cf_pad, wf.shape[1], 3, 3), dtype=torch.float32, device=dev)
        wfp[:c_first] = wf.detach().float()
        self.ops.append(first_layer(self, S, torch.bfloat16, n, h, w, slabs[0].view(0, cf_pad), wfp, padded_bias(bf, cf_pad),
                                    relu=False, cout_real=c_first))

        # ---- down path
        ident = lambda k: torch.arange(k, device=dev)
        for l in range(L):
            cur = c_in[l]
            for layer in spec['down'][l]:
                dense_layer(slabs[l], l, cur, ident(cur), layer, cur)
                cur += g
            # TransitionDown: BN -> ReLU -> conv1x1(cur -> cur) -> maxpool, into the next level's input channels
            bn, wt, bs = spec['trans_down'][l]
            cpad = _pad32(cur)
            sc, sh = bn_params(bn, ident(cur), cpad)
            z = scratch[l].view(0, cpad)
            self.ops.append(BnReluOp(slabs[l].view(0, cpad), z, sc, sh))
            tmp = S(h >> l, w >> l, cpad)
            wp = pad_conv_weight(wt, ident(cur), cpad, cpad)
            self.ops.append(ConvOp(N.CONV_1X1, z, tmp.view(), pack_conv1x1(wp), padded_bias(bs, cpad), relu=False, real=(cur, cur)))
            dst = (slabs[l + 1] if l + 1 < L else bott).view(0, cur)
            self.ops.append(PoolOp(tmp.view(0, cur), dst))

        # ---- bottleneck: dense layers on the growing slab; its output is only the new features
        cur = c_bott_in
        for layer in spec['bottleneck']:
            dense_layer(bott, L, cur, ident(cur), layer, cur)
            cur += g
        new_src = bott.view(c_bott_in, _pad32(g * n_bott))                   # 80 new channels + zero tail

        # ---- up path
        for i in range(L):
            l = L - 1 - i
            skip = c_in[l] + g * len(spec['down'][l])
            cu = up_prev[i]
            wt, bs = spec['trans_up'][i]
            cu_pad = _pad32(cu)
            # TransitionUp: ConvTranspose2d(k3, s2) cropped to the skip size, written right after the skip channels
            self.ops.append(ConvOp(N.CONVT_3X3_S2, new_src, slabs[l].view(skip, cu_pad),
                                   pack_convT3x3(wt, new_src.c, cu_pad), padded_bias(bs, cu_pad), relu=False, real=(cu, cu)))
            # reference channel order of the block input is [up | skip | new...]; the slab holds [skip | up | new...]
            cur = skip + cu
            base_map = torch.cat([torch.arange(skip, skip + cu, device=dev), torch.arange(0, skip, device=dev)])
            for k, layer in enumerate(spec['up'][i]):
                cmap = torch.cat([base_map, torch.arange(skip + cu, cur, device=dev)])
                dense_layer(slabs[l], l, cur, cmap, layer, cur)
                cur += g
            new_src = slabs[l].view(skip + cu, _pad32(g * len(spec['up'][i])))
            last_cur, last_map = cur, torch.cat([base_map, torch.arange(skip + cu, cur, device=dev)])

        # ---- final 1x1 conv (C -> 1) through the fused head: a 32-row conv whose row 0 is the real filter
        wt, bs = spec['final']
        cpad = _pad32(last_cur)
        wp = pad_conv_weight(wt, last_map, cpad, 32)
        self.out = torch.empty((n, h, w), dtype=torch.float32, device=dev)
        pick = torch.zeros(32, dtype=torch.float32, device=dev)
        pick[0] = 1.0
        self.ops.append(ConvOp(N.CONV_1X1, slabs[0].view(0, cpad), None, pack_conv1x1(wp), padded_bias(bs, 32), relu=False,
                               head=(pick, 0.0, sigmoid, self.out), real=(last_cur, 1)))
        self.flops = sum(op.flops for op in self.ops)
        self.launches = sum(op.launches for op in self.ops)

    load_nchw = VGGUNetPlan.load_nchw
    input_layout = VGGUNetPlan.input_layout
    run = VGGUNetPlan.run


# ------------------------------------------------------------------------------------------------ LinkNet34
def pack_conv3x3_s2(weight, dtype=torch.bfloat16):
    """Stride-2 conv3x3 (padding 1) as a 4-tap conv over the space-to-depth tensor: [Cout, Cin, 3, 3] ->
    [4 taps][Cout][4 * Cin]; tap (ty, tx) reads the 2x2 block at offset (ty - 1, tx - 1), block channel
    (py*2+px)*Cin + ci holds input pixel (2y'+py, 2x'+px).  Row 2y+ky-1 is (block y-1, py=1) for ky=0 and
    (block y, py=ky-1) for ky=1,2; unused (tap, parity) pairs stay zero."""
    cout, cin = weight.shape[:2]
    w = _f(weight)
    out = torch.zeros((4, cout, 4 * cin), dtype=w.dtype, device=weight.device)
    kmap = {(0, 1): 0, (1, 0): 1, (1, 1): 2}          # (tap index t, parity p) -> kernel index
    for (ty, py), ky in kmap.items():
        for (tx, px), kx in kmap.items():
            q = py * 2 + px
            out[ty * 2 + tx, :, q * cin:(q + 1) * cin] = w[:, :, ky, kx]
    return to_storage(out, dtype).contiguous()


def pack_conv2x2(weight, cin_pad=None, cout_pad=None, dtype=torch.bfloat16):
    """nn.Conv2d(k=2, padding=1) weight [Cout, Cin, 2, 2] -> [4][cout_pad][cin_pad], tap = ky*2 + kx."""
    cout, cin = weight.shape[:2]
    w = _f(weight)
    out = torch.zeros((4, cout_pad or cout, cin_pad or cin), dtype=w.dtype, device=weight.device)
    out[:, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(4, cout, cin)
    return to_storage(out, dtype).contiguous()


def pack_stem7x7(weight, k_pad, dtype=torch.bfloat16):
    """Stride-2 7x7 stem [Cout, C, 7, 7] -> [1][Cout][k_pad] matching snb_stem7x7_rows: k = (ky*7+kx)*C + c."""
    cout, cin = weight.shape[:2]
    w = _f(weight)
    out = torch.zeros((1, cout, k_pad), dtype=w.dtype, device=weight.device)
    out[0, :, :49 * cin] = w.permute(0, 2, 3, 1).reshape(cout, 49 * cin)
    return to_storage(out, dtype).contiguous()


# ---- input-gradient operands: every dgrad of LinkNet34 is one of the FORWARD kernels on transformed weights --------------
def pack_conv_dgrad(weight, cin_pad=None, cout_pad=None, dtype=torch.bfloat16):
    """Stride-1 conv3x3 (padding 1, or none: run the adjoint with valid=2) / conv1x1 [Cout, Cin, k, k]: the adjoint is the
    same convolution with the taps flipped and Cin / Cout swapped -> [k*k][cin_pad][cout_pad]."""
    cout, cin, k = weight.shape[0], weight.shape[1], weight.shape[2]
    wt = _pad_mat(weight.detach().double(), cout_pad or cout, cin_pad or cin).flip(2, 3).transpose(0, 1).contiguous()
    return pack_conv3x3(wt, dtype) if k == 3 else pack_conv1x1(wt, dtype)


def pack_conv3x3_s2_dgrad(weight, dtype=torch.bfloat16):
    """Stride-2 conv3x3 (padding 1) [Cout, Cin, 3, 3]: its input gradient in space-to-depth form is a 4-tap convolution with
    taps {0,+1}^2 (SNB_CONV_2X2_ADJ) from Cout to 4 * Cin channels: [4][4 * Cin][Cout] = pack_conv3x3_s2 with the taps
    reversed and transposed; snb_depth_to_space2 then restores the pixel grid."""
    return to_storage(pack_conv3x3_s2(weight, torch.float64).flip(0).transpose(1, 2), dtype).contiguous()


_CONVT_D4 = ((0, -1), (1, 0))      # ConvTranspose k4 s2 p1: input row offset per (phase parity, tap), csrc tap tables
_CONVT_D3 = ((0, -1), (0, 0))      # ConvTranspose k3 s2 p0


def pack_convT4x4_dgrad(weight, cin_pad=None, cout_pad=None, dtype=torch.bfloat16):
    """nn.ConvTranspose2d(k=4, s=2, p=1) [Cin, Cout, 4, 4]: input gradient = conv3x3 (padding 1) over the space-to-depth
    copy of the output gradient (4 * cout_pad channels, block = sub-pixel phase) -> [9][cin_pad][4 * cout_pad]; 16 of the
    36 (tap, phase) blocks are non-zero."""
    cin, cout = weight.shape[:2]
    cin_pad, cout_pad = cin_pad or cin, cout_pad or cout
    w = weight.detach().double()
    out = torch.zeros((9, cin_pad, 4 * cout_pad), dtype=torch.float64, device=weight.device)
    for py in range(2):
        for px in range(2):
            ph = py * 2 + px
            for ty in range(2):
                for tx in range(2):
                    dy, dx = -_CONVT_D4[py][ty], -_CONVT_D4[px][tx]          # the gradient reads dOut4[r - d]
                    tap = (dy + 1) * 3 + (dx + 1)
                    out[tap, :cin, ph * cout_pad:ph * cout_pad + cout] += w[:, :, _CONVT_K[py][ty], _CONVT_K[px][tx]]
    return to_storage(out, dtype).contiguous()


def pack_convT3x3_full_dgrad(weight, cin_pad=None, cout_pad=None, dtype=torch.bfloat16):
    """nn.ConvTranspose2d(k=3, s=2, p=0) uncropped [Cin, Cout, 3, 3] (output 2h+1): input gradient = 4-tap convolution with
    taps {0,+1}^2 (SNB_CONV_2X2_ADJ, valid) over the space-to-depth copy of the output gradient (h+1 blocks of
    4 * cout_pad channels) -> [4][cin_pad][4 * cout_pad]."""
    cin, cout = weight.shape[:2]
    cin_pad, cout_pad = cin_pad or cin, cout_pad or cout
    w = weight.detach().double()
    out = torch.zeros((4, cin_pad, 4 * cout_pad), dtype=torch.float64, device=weight.device)
    for py in range(2):
        for px in range(2):
            ph = py * 2 + px
            for ty in range(2):
                for tx in range(2):
                    ky, kx = _CONVT3_K[py][ty], _CONVT3_K[px][tx]
                    if ky is None or kx is None:
                        continue
                    sy, sx = -_CONVT_D3[py][ty], -_CONVT_D3[px][tx]
                    out[sy * 2 + sx, :cin, ph * cout_pad:ph * cout_pad + cout] += w[:, :, ky, kx]
    return to_storage(out, dtype).contiguous()


def pack_conv2x2_dgrad(weight, cin_pad=None, cout_pad=None, dtype=torch.bfloat16):
    """nn.Conv2d(k=2, padding=1) [Cout, Cin, 2, 2]: input gradient = 4-tap convolution with taps {0,+1}^2 (valid) and the
    kernel reversed -> [4][cin_pad][cout_pad]."""
    cout, cin = weight.shape[:2]
    w = _pad_mat(weight.detach().double(), cout_pad or cout, cin_pad or cin)
    return to_storage(w.flip(2, 3).permute(2, 3, 1, 0).reshape(4, w.shape[1], w.shape[0]), dtype).contiguous()


def _pad_mat(w, cout_pad, cin_pad):
    out = torch.zeros((cout_pad, cin_pad) + tuple(w.shape[2:]), dtype=w.dtype if w.dtype == torch.float64 else torch.float32,
                      device=w.device)
    out[:w.shape[0], :w.shape[1]] = w
    return out


def _pad_vec(b, n):
    out = torch.zeros(n, dtype=torch.float32, device=b.device)
    out[:b.numel()] = b
    return out


class SimpleOp:
    """A helper-kernel launch with fixed arguments (space-to-depth, 3x3/s2 max-pool)."""

    def __init__(self, fn_name, args, keep):
        self.fn_name, self.args, self.keep = fn_name, args, keep
        self.flops, self.launches = 0.0, 1

    def __call__(self, stream):
        N.check(getattr(N.lib(), self.fn_name)(*self.args, stream))


class LinkNet34Plan:
    """LinkNet34 forward (lib/models/linknet.py:65-90) in eval mode on the native kernels.

    ResNet-34 encoder: BatchNorm folded into the (bias-free) convolutions; the 7x7/s2 stem is a GEMM over im2col rows;
    stride-2 blocks run on a space-to-depth copy (conv3x3/s2 -> 4-tap conv, 1x1/s2 -> conv1x1 on its first quarter);
    the block's identity / down-sample branch enters the second conv's epilogue as a residual operand before the ReLU.
    Decoders: conv1x1 -> ConvTranspose k4s2 -> conv1x1, each with InPlaceABN folded (scale |w|+eps as in mapillary's
    kernel, leaky-ReLU 0.01 in the epilogue); the additive skips `decoderK(x) + eK` are residual operands applied
    after the activation.  Head: ConvTranspose k3s2 (uncropped, 2h+1) -> valid conv3x3 -> conv k2 p1, Dropout2d off.
    `spec` comes from snb_b200.lib.models.linknet.LinkNet34._spec().
    """

    STEM_K = 160

    def __init__(self, spec, n, h, w, device, sigmoid):
        if h % 32 or w % 32:
            raise ValueError("height and width must be multiples of 32")
        if spec['final3'][0].shape[0] != 1:
            raise NotImplementedError("fused head expects num_classes == 1")
        self.n, self.h, self.w, self.device = n, h, w, device
        self.dtype = torch.bfloat16
        self.ops = []
        S = lambda hh, ww, c: Slab(n, hh, ww, c, device)
        p32 = lambda c: (c + 31) // 32 * 32

        def fold(wt, bias, bn, transposed=False):
            """conv (+ optional bias) followed by BatchNorm(eval) / InPlaceABN(eval) -> (weight, bias) in fp32."""
            gamma, beta, mean, var, eps, abn = bn
            g = (gamma.detach().float().abs() + eps) if abn else gamma.detach().float()
            scale = g / torch.sqrt(var.detach().float() + eps)
            b0 = bias.detach().float() if bias is not None else torch.zeros_like(scale)
            shape = (1, -1, 1, 1) if transposed else (-1, 1, 1, 1)
            return wt.detach().float() * scale.view(shape), (b0 - mean.detach().float()) * scale + beta.detach().float()

        # ---- stem: 7x7/s2 conv + BN + ReLU as a GEMM over im2col rows, then MaxPool 3x3/s2
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        # network input: normalised float NCHW (what the tiled predictor's split kernel writes with LAYOUT_NCHW_F32)
        self.x_nchw = torch.empty((n, 3, h, w), dtype=torch.float32, device=device)
        self.x_rows = S(h2, w2, self.STEM_K)
        self.ops.append(SimpleOp("snb_stem7x7_rows", (N.c_vp(self.x_nchw.data_ptr()), n, 3, h, w,
                                                      N.c_vp(self.x_rows.t.data_ptr()), self.STEM_K), (self.x_nchw, self.x_rows)))
        wt, bs = fold(spec['stem'][0], None, spec['stem'][1])
        stem = S(h2, w2, 64)
        self.ops.append(ConvOp(N.CONV_1X1, self.x_rows.view(), stem.view(), pack_stem7x7(wt, self.STEM_K), bs.contiguous(),
                               real=(147, 64)))
        cur = S(h4, w4, 64)
        self.ops.append(SimpleOp("snb_maxpool3x3s2", (N.c_vp(stem.view().ptr), n, h2, w2, 64, 64, N.c_vp(cur.view().ptr), 64),
                                 (stem, cur)))
        cur, ch, hh, ww = cur.view(), 64, h4, w4

        # ---- encoders (BasicBlocks)
        skips = []
        for blocks in spec['encoders']:
            for blk in blocks:
                cout = blk['conv1'].shape[0]
                if blk['down'] is not None:                      # stride-2 block
                    hh, ww = hh // 2, ww // 2
                    x4 = S(hh, ww, 4 * ch)
                    self.ops.append(SimpleOp("snb_space_to_depth2", (N.c_vp(cur.ptr), n, 2 * hh, 2 * ww, ch, cur.cstride,
                                                                     N.c_vp(x4.view().ptr), 4 * ch), (cur, x4)))
                    w1, b1 = fold(blk['conv1'], None, blk['bn1'])
                    t = S(hh, ww, cout).view()
                    self.ops.append(ConvOp(N.CONV_2X2, x4.view(), t, pack_conv3x3_s2(w1), b1.contiguous(), valid=True,
                                           real=(4 * ch * 9 / 16.0, cout)))   # 9 of the 16 (tap, parity) blocks are real
                    wd, bd = fold(blk['down'][0], None, blk['down'][1])
                    ident = S(hh, ww, cout).view()
                    self.ops.append(ConvOp(N.CONV_1X1, x4.view(0, ch), ident, pack_conv1x1(wd), bd.contiguous(), relu=False))
                else:
                    w1, b1 = fold(blk['conv1'], None, blk['bn1'])
                    t = S(hh, ww, cout).view()
                    self.ops.append(ConvOp(N.CONV_3X3, cur, t, pack_conv3x3(w1), b1.contiguous()))
                    ident = cur
                w2_, b2 = fold(blk['conv2'], None, blk['bn2'])
                y = S(hh, ww, cout).view()
                self.ops.append(ConvOp(N.CONV_3X3, t, y, pack_conv3x3(w2_), b2.contiguous(), residual=ident))
                cur, ch = y, cout
            skips.append(cur)

        # ---- decoders with additive skips
        def decoder(x, cin, spec_d, n_out, hh, ww, skip):
            mid = cin // 4
            mp = p32(mid)
            w1, b1 = fold(spec_d['conv1'][0], spec_d['conv1'][1], spec_d['abn1'])
            a = S(hh, ww, mp).view()
            self.ops.append(ConvOp(N.CONV_1X1, x, a, pack_conv1x1(_pad_mat(w1, mp, cin)), _pad_vec(b1, mp), act_slope=0.01,
                                   real=(cin, mid)))
            w2_, b2 = fold(spec_d['deconv2'][0], spec_d['deconv2'][1], spec_d['abn2'], transposed=True)
            b_ = S(2 * hh, 2 * ww, mp).view()
            self.ops.append(ConvOp(N.CONVT_4X4_S2, a, b_, pack_convT4x4(_pad_mat(w2_, mp, mp)), _pad_vec(b2, mp), act_slope=0.01,
                                   real=(mid, mid)))
            w3, b3 = fold(spec_d['conv3'][0], spec_d['conv3'][1], spec_d['abn3'])
            y = S(2 * hh, 2 * ww, n_out).view()
            self.ops.append(ConvOp(N.CONV_1X1, b_, y, pack_conv1x1(_pad_mat(w3, n_out, mp)), b3.contiguous(), act_slope=0.01,
                                   residual=skip, res_after_act=True, real=(mid, n_out)))
            return y

        e1, e2, e3, e4 = skips
        h32, w32 = h // 32, w // 32
        d4 = decoder(e4, 512, spec['decoders'][3], 256, h32, w32, e3)
        d3 = decoder(d4, 256, spec['decoders'][2], 128, 2 * h32, 2 * w32, e2)
        d2 = decoder(d3, 128, spec['decoders'][1], 64, 4 * h32, 4 * w32, e1)
        d1 = decoder(d2, 64, spec['decoders'][0], 64, 8 * h32, 8 * w32, None)

        # ---- final classifier: ConvT k3 s2 (2h+1) -> LeakyReLU -> conv3x3 valid (2h-1) -> LeakyReLU -> conv k2 p1 (2h)
        wt, bs = spec['final1']
        f1 = S(h + 1, w + 1, 32)
        self.ops.append(ConvOp(N.CONVT_3X3_S2_FULL, d1, f1.view(), pack_convT3x3(wt, 64, 32), bs.detach().float().contiguous(),
                               act_slope=0.01))
        wt, bs = spec['final2']
        f3 = S(h - 1, w - 1, 32)
        self.ops.append(ConvOp(N.CONV_3X3, f1.view(), f3.view(), pack_conv3x3(wt), bs.detach().float().contiguous(),
                               act_slope=0.01, valid=True))
        wt, bs = spec['final3']
        self.out = torch.empty((n, h, w), dtype=torch.float32, device=device)
        pick = torch.zeros(32, dtype=torch.float32, device=device)
        pick[0] = 1.0
        self.ops.append(ConvOp(N.CONV_2X2, f3.view(), None, pack_conv2x2(wt, 32, 32), _pad_vec(bs.detach().float(), 32),
                               relu=False, head=(pick, 0.0, sigmoid, self.out), real=(32, 1)))
        self.flops = sum(op.flops for op in self.ops)
        self.launches = sum(op.launches for op in self.ops)

    def load_nchw(self, x):
        self.x_nchw.copy_(x)

    run = VGGUNetPlan.run


class BnTrainOp:
    """Training-mode BatchNorm2d / InPlaceABN on an NHWC bf16 slab (snb_bn_train_nhwc): batch statistics, in-place update
    of the module's running statistics, normalise + activation (+ residual before or after it).  `bn` = (weight, bias,
    running_mean, running_var, eps, momentum); parameters are read through their own pointers (no copies) unless the
    slab is wider than the layer (zero-padded channels), in which case padded temporaries are used and the running
    statistics are copied back."""

    def __init__(self, src, dst, bn, abn, slope, residual=None, res_after_act=False):
        weight, bias, rmean, rvar, eps, momentum = bn
        c, cpad, dev = weight.numel(), src.c, src.slab.t.device
        if dst.c != cpad or (residual is not None and residual.c != cpad):
            raise ValueError("source, destination and residual must have the same channel count")
        self.padded = c != cpad
        f = lambda t: t.detach()
        if self.padded:
            pad = lambda t, fill: torch.cat([f(t).float(), torch.full((cpad - c,), fill, dtype=torch.float32, device=dev)])
            self.gamma, self.beta = pad(weight, 0.0), pad(bias, 0.0)
            self.rmean, self.rvar = pad(rmean, 0.0), pad(rvar, 1.0)
            self.module_stats = (rmean, rvar, c)
        else:
            self.gamma, self.beta, self.rmean, self.rvar = f(weight), f(bias), rmean, rvar
        for t in (self.gamma, self.beta, self.rmean, self.rvar):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("BatchNorm parameters and buffers must be contiguous float32")
        self.scale = torch.empty(cpad, dtype=torch.float32, device=dev)
        self.shift = torch.empty(cpad, dtype=torch.float32, device=dev)
        self.mean = torch.empty(cpad, dtype=torch.float32, device=dev)     # kept for the backward pass
        self.var = torch.empty(cpad, dtype=torch.float32, device=dev)
        self.work = torch.zeros(3 * cpad + 2, dtype=torch.float64, device=dev)      # zero at rest (include/snb_b200.h)
        self.keep = (src, dst, residual, weight, bias, rmean, rvar)
        self.pixels = src.slab.n * src.slab.h * src.slab.w
        self._cfg = (cpad, abn, eps, momentum, slope, res_after_act)
        self.rebind()
        self.flops, self.launches = 0.0, 2

    def rebind(self):
        """(re)build the argument list after self.gamma / self.beta were pointed at other tensors"""
        src, dst, residual = self.keep[:3]
        cpad, abn, eps, momentum, slope, res_after_act = self._cfg
        self.args = (N.c_vp(src.ptr), self.pixels, cpad, src.cstride, N.ptr(self.gamma), N.ptr(self.beta), 1 if abn else 0,
                     float(eps), float(momentum), N.ptr(self.rmean), N.ptr(self.rvar), float(slope),
                     N.c_vp(residual.ptr if residual is not None else 0), residual.cstride if residual is not None else 0,
                     1 if res_after_act else 0, N.c_vp(dst.ptr), dst.cstride, N.ptr(self.scale), N.ptr(self.shift),
                     N.ptr(self.mean), N.ptr(self.var), N.ptr(self.work))

    def refresh_stats(self):
        """padded layers: re-read the module's running statistics (they are copied back after each run)"""
        rmean, rvar = self.keep[5:7]
        c = rmean.numel()
        self.rmean[:c].copy_(rmean)
        self.rvar[:c].copy_(rvar)

    def refresh(self):
        """padded layers keep copies of the parameters: re-read them (running statistics are copied back after each run)"""
        weight, bias, rmean, rvar = self.keep[3:7]
        c = weight.numel()
        self.gamma[:c].copy_(weight.detach())
        self.beta[:c].copy_(bias.detach())
        self.rmean[:c].copy_(rmean)
        self.rvar[:c].copy_(rvar)

    def __call__(self, stream):
        N.check(N.lib().snb_bn_train_nhwc(*self.args, stream))
        if self.padded:
            rmean, rvar, c = self.module_stats
            rmean.data.copy_(self.rmean[:c])       # through .data: the plan's own update must not look like an external
            rvar.data.copy_(self.rvar[:c])         # modification of the buffers (their version counters stay put)
