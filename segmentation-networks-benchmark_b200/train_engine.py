"""Training-step plan of LinkNet34 (BASELINE configs[1]; reference lib/models/linknet.py:65-90 under torch_train.py:183-190):
train-mode forward (batch statistics), fused loss, and the whole backward pass as STATIC lists of kernel launches on
preallocated buffers, so that forward and backward each replay as one CUDA graph.

Every contraction runs on the tcgen05 kernels:
  forward         conv_tcgen05.cu (the inference kernels, nothing folded), BatchNorm / InPlaceABN on batch statistics
  input gradients the same forward kernels on transformed weights (engine.pack_*_dgrad): stride-1 conv3x3 / conv1x1 = the
                  adjoint convolution; stride-2 conv3x3 = 4-tap convolution into space-to-depth channels + depth-to-space;
                  ConvTranspose k4 s2 = conv3x3 over the space-to-depth gradient; the head through "full" / adjoint taps
  weight gradients conv_wgrad.cu (K = pixels, MN-major operands), in the packed layout of the forward weights
Parameters are packed into the bf16 operands, and packed fp32 weight gradients are scattered back into the parameters'
own layouts, by ONE multi-segment gather launch each (index maps obtained by pushing an index tensor through the same
pack functions).  Gradients of all parameters live in one flat fp32 arena.
"""
import ctypes
import os

import torch

from . import _native as N
from . import engine as E
from .engine import BnTrainOp, ConvOp, SimpleOp, Slab, WgradOp


def _index_map(fn, shape, device):
    """Push 1-based element indices of a parameter through a pack function: -> (int32 [packed numel] source index or -1,
    packed shape)."""
    n = 1
    for d in shape:
        n *= int(d)
    src = (torch.arange(n, dtype=torch.float64, device=device) + 1).reshape(shape)
    packed = fn(src)
    if packed.dtype != torch.float64:
        raise TypeError("pack function does not preserve float64 (index maps need exact values)")
    idx = (packed.reshape(-1).round().long() - 1).to(torch.int32)
    return idx.contiguous(), tuple(packed.shape)


def _inverse_map(idx, n):
    """For every source element its position in the packed tensor (each element must appear exactly once)."""
    pos = torch.nonzero(idx >= 0).reshape(-1)
    inv = torch.full((n,), -1, dtype=torch.int32, device=idx.device)
    inv[idx[pos].long()] = pos.to(torch.int32)
    if pos.numel() != n or bool((inv < 0).any()):
        raise ValueError("pack function is not a bijection onto its non-zero entries")
    return inv.contiguous()


class GatherTable:
    """Device table of snb_gather_seg rows; one launch moves every segment."""

    def __init__(self, device, dst_bf16):
        self.device, self.dst_bf16 = device, 1 if dst_bf16 else 0
        self.rows, self.keep, self.blocks, self.table = [], [], 0, None
        self.launches, self.flops = 1, 0.0

    def add(self, src, dst, idx):
        if src.dtype != torch.float32 or not src.is_contiguous() or not dst.is_contiguous() or idx.dtype != torch.int32:
            raise ValueError("gather segments need contiguous float sources, contiguous destinations and int32 maps")
        if dst.numel() != idx.numel():
            raise ValueError("destination and index map differ in size")
        self.rows.append((src.data_ptr(), dst.data_ptr(), idx.data_ptr(), idx.numel(), self.blocks))
        self.blocks += (idx.numel() + 1023) // 1024
        self.keep += [src, dst, idx]
        self.table = None

    def __call__(self, stream):
        if not self.rows:
            return
        if self.table is None:
            self.table = torch.tensor(self.rows, dtype=torch.int64, device=self.device)
        N.check(N.lib().snb_gather_segments(N.ptr(self.table), len(self.rows), self.blocks, self.dst_bf16, stream))


class BnBwdOp:
    """snb_bn_backward_nhwc on preallocated buffers (BatchNorm / InPlaceABN + activation + residual backward)."""

    def __init__(self, raw, g_out, bnop, abn, eps, slope, res_before, draw, dres, dgamma, dbeta):
        cpad = raw.c
        self.keep = (raw, g_out, bnop, res_before, draw, dres, dgamma, dbeta)
        self.work = torch.zeros(3 * cpad + 2, dtype=torch.float64, device=raw.slab.t.device)
        pixels = raw.slab.n * raw.slab.h * raw.slab.w
        self.args = (N.c_vp(raw.ptr), raw.cstride, N.c_vp(g_out.ptr), g_out.cstride, pixels, cpad, N.ptr(bnop.scale),
                     N.ptr(bnop.shift), N.ptr(bnop.mean), N.ptr(bnop.var), N.ptr(bnop.gamma), 1 if abn else 0, float(eps),
                     float(slope), N.c_vp(res_before.ptr if res_before is not None else 0),
                     res_before.cstride if res_before is not None else 0, N.c_vp(draw.ptr), draw.cstride,
                     N.c_vp(dres.ptr if dres is not None else 0), dres.cstride if dres is not None else 0, N.ptr(dgamma),
                     N.ptr(dbeta), N.ptr(self.work))
        self.flops, self.launches = 0.0, 2

    def __call__(self, stream):
        N.check(N.lib().snb_bn_backward_nhwc(*self.args, stream))


class TorchOp:
    """A few ATen calls on fixed tensors (copies / fills that are not worth a kernel of their own); graph-capturable."""

    def __init__(self, fn, launches=1):
        self.fn, self.launches, self.flops = fn, launches, 0.0

    def __call__(self, stream):
        self.fn()


def _p32(c):
    return (c + 31) // 32 * 32


class LinkNet34TrainPlan:
    """LinkNet34 in train() mode for one (batch, height, width): `run()` = forward (logits in self.out), `backward(dlogits)`
    = gradients of every parameter.  `model` is the snb_b200.lib.models.LinkNet34 module; its parameters and BatchNorm
    buffers are read and updated through their own storages (the plan is rebuilt if they are replaced)."""

    STEM_K = 160

    def __init__(self, model, n, h, w, device, linear=False):
        # linear=True (tests only): every ReLU / leaky-ReLU becomes the identity, so gradients can be compared tightly
        if h % 32 or w % 32:
            raise ValueError("height and width must be multiples of 32")
        if model.finalconv3.weight.shape[0] != 1:
            raise NotImplementedError("fused head expects num_classes == 1")
        for p in model.parameters():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("parameters must be contiguous float32 tensors")
        self.n, self.h, self.w, self.device = n, h, w, device
        self.model, self.linear = model, linear
        self.ops, self.bwd_ops, self.bn_modules = [], [], []
        self.repack = GatherTable(device, True)       # parameters -> bf16 operands of the forward and dgrad convolutions
        self.repack32 = GatherTable(device, False)    # biases / norm weights -> zero-padded float vectors
        self.unpack = GatherTable(device, False)      # packed fp32 weight gradients -> parameter layouts
        self._pending_shortcut = None
        self._drop_is_ones = True
        self._nbt = None
        self.generation = 0                           # bumped by every forward: a stale backward must not run
        self.use_graph = os.environ.get("SNB_TRAIN_GRAPH", "1") != "0"
        self._graphs = {}
        dev = device
        S = lambda hh, ww, c: Slab(n, hh, ww, c, dev)
        zeros = lambda c: torch.zeros(c, dtype=torch.float32, device=dev)

        # ---- flat gradient arena: one slot per parameter (1-D slots padded to the 32-channel slab width)
        offs, total = {}, 0
        for p in model.parameters():
            size = _p32(p.numel()) if p.dim() == 1 else p.numel()
            offs[p] = total
            total += (size + 31) // 32 * 32
        self.grad_arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad_slot = {p: self.grad_arena[o:o + (_p32(p.numel()) if p.dim() == 1 else p.numel())] for p, o in offs.items()}
        self.grads = {p: self.grad_slot[p][:p.numel()].view(p.shape) for p in offs}
        self._wgrad_bufs = []

        def packed(param, fn):
            """bf16 operand of `param` under the pack function `fn`, kept current by the repack table."""
            idx, shape = _index_map(fn, tuple(param.shape), dev)
            t = torch.empty(shape, dtype=torch.bfloat16, device=dev)
            self.repack.add(param.detach(), t.view(-1), idx)
            return t, idx, shape

        def wgrad_buffer(param, idx, shape):
            """packed fp32 gradient of `param` + its scatter back into the parameter's own layout"""
            buf = torch.zeros(shape, dtype=torch.float32, device=dev)
            self._wgrad_bufs.append(buf)
            self.unpack.add(buf.view(-1), self.grads[param].view(-1), _inverse_map(idx, param.numel()))
            return buf

        def padded_vec(param, c):
            """float vector of `param` zero-padded to c entries (the parameter itself when it already has c): bias / norm
            weight operands of layers whose slab is wider than the layer"""
            if param.numel() == c:
                return param.detach()
            t = torch.zeros(c, dtype=torch.float32, device=dev)
            idx = torch.full((c,), -1, dtype=torch.int32, device=dev)
            idx[:param.numel()] = torch.arange(param.numel(), dtype=torch.int32, device=dev)
            self.repack32.add(param.detach(), t, idx)
            return t

        tape = []

        def conv_bn(kind, src, conv, fn, cout, hh, ww, m, abn, slope, bias=None, residual=None, res_after_act=False,
                    valid=0, dsrc=None, **extra):
            """conv -> BatchNorm / ABN (batch statistics) -> activation (+ residual).  `src` is what the forward kernel reads
            (possibly a space-to-depth copy or im2col rows); `dsrc` the tensor whose gradient the backward produces."""
            wp, idx, shape = packed(conv.weight, fn)
            raw = S(hh, ww, cout).view()
            self.ops.append(ConvOp(kind, src, raw, wp, zeros(cout) if bias is None else bias, relu=False, valid=valid))
            out = S(hh, ww, cout).view()
            sl = -1.0 if linear else slope
            bnop = BnTrainOp(raw, out, (m.weight, m.bias, m.running_mean, m.running_var, m.eps, m.momentum), abn, sl, residual,
                             res_after_act)
            if bnop.padded:          # keep the padded copies of the norm weight / bias current through the gather table
                bnop.gamma, bnop.beta = padded_vec(m.weight, cout), padded_vec(m.bias, cout)
                bnop.rebind()
            self.ops.append(bnop)
            if isinstance(m, torch.nn.BatchNorm2d):
                self.bn_modules.append(m)
            node = dict(kind='conv_bn', fwd_kind=kind, valid=valid, conv=conv, src=src, dsrc=dsrc, raw=raw, out=out, bnop=bnop,
                        bn=m, abn=abn, slope=sl, res=residual, res_after=res_after_act, widx=idx, wshape=shape)
            node.update(extra)
            tape.append(node)
            return out

        # ---- stem: 7x7/s2 conv as a GEMM over im2col rows, BatchNorm + ReLU, MaxPool 3x3/s2
        h2, w2, h4, w4 = h // 2, w // 2, h // 4, w // 4
        self.x_nchw = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
        self.x_rows = S(h2, w2, self.STEM_K)
        self.ops.append(SimpleOp("snb_stem7x7_rows", (N.c_vp(self.x_nchw.data_ptr()), n, 3, h, w,
                                                      N.c_vp(self.x_rows.t.data_ptr()), self.STEM_K), (self.x_nchw, self.x_rows)))
        stem = conv_bn(N.CONV_1X1, self.x_rows.view(), model.firstconv, lambda t: E.pack_stem7x7(t, self.STEM_K, torch.float64),
                       64, h2, w2, model.firstbn, False, 0.0, dsrc=None)
        cur = S(h4, w4, 64)
        self.ops.append(SimpleOp("snb_maxpool3x3s2", (N.c_vp(stem.ptr), n, h2, w2, 64, 64, N.c_vp(cur.view().ptr), 64),
                                 (stem, cur)))
        tape.append(dict(kind='maxpool', src=stem, out=cur.view()))
        cur, ch, hh, ww = cur.view(), 64, h4, w4

        F64 = torch.float64
        skips = []
        for li in range(1, 5):
            for blk in getattr(model, 'encoder%d' % li):
                cout = blk.conv1.weight.shape[0]
                if blk.downsample is not None:
                    hh, ww = hh // 2, ww // 2
                    x4 = S(hh, ww, 4 * ch)
                    self.ops.append(SimpleOp("snb_space_to_depth2", (N.c_vp(cur.ptr), n, 2 * hh, 2 * ww, ch, cur.cstride,
                                                                     N.c_vp(x4.view().ptr), 4 * ch), (cur, x4)))
                    dx4 = S(hh, ww, 4 * ch)          # gradient of the space-to-depth copy (both consumers write into it)
                    t = conv_bn(N.CONV_2X2, x4.view(), blk.conv1, lambda q: E.pack_conv3x3_s2(q, F64), cout, hh, ww, blk.bn1,
                                False, 0.0, valid=1, dsrc=cur, dx4=dx4, role='s2_main')
                    ident = conv_bn(N.CONV_1X1, x4.view(0, ch), blk.downsample[0], lambda q: E.pack_conv1x1(q, F64), cout, hh, ww,
                                    blk.downsample[1], False, -1.0, dsrc=cur, dx4=dx4, role='s2_shortcut')
                else:
                    t = conv_bn(N.CONV_3X3, cur, blk.conv1, lambda q: E.pack_conv3x3(q, F64), cout, hh, ww, blk.bn1, False, 0.0,
                                dsrc=cur)
                    ident = cur
                cur = conv_bn(N.CONV_3X3, t, blk.conv2, lambda q: E.pack_conv3x3(q, F64), cout, hh, ww, blk.bn2, False, 0.0,
                              residual=ident, dsrc=t)
                ch = cout
            skips.append(cur)

        def decoder(x, cin, d, n_out, hh, ww, skip):
            mid = cin // 4
            mp = _p32(mid)
            bias = lambda conv, c: padded_vec(conv.bias, c)
            # (conv biases in front of a batch norm only move the running mean: the mean subtraction removes them from the
            # output, so their gradient is exactly zero and their slots of the gradient arena stay zero)
            a = conv_bn(N.CONV_1X1, x, d.conv1, lambda q: E.pack_conv1x1(E._pad_mat(q, mp, cin), F64), mp, hh, ww, d.abn1, True,
                        d.abn1.slope, bias=bias(d.conv1, mp), dsrc=x, real_cin=cin, real_cout=mid)
            b_ = conv_bn(N.CONVT_4X4_S2, a, d.deconv2, lambda q: E.pack_convT4x4(E._pad_mat(q, mp, mp), F64), mp, 2 * hh, 2 * ww,
                         d.abn2, True, d.abn2.slope, bias=bias(d.deconv2, mp), dsrc=a, real_cin=mid, real_cout=mid, mp=mp)
            return conv_bn(N.CONV_1X1, b_, d.conv3, lambda q: E.pack_conv1x1(E._pad_mat(q, n_out, mp), F64), n_out, 2 * hh,
                           2 * ww, d.abn3, True, d.abn3.slope, bias=bias(d.conv3, n_out), residual=skip, res_after_act=True,
                           dsrc=b_, real_cin=mid, real_cout=n_out)

        e1, e2, e3, e4 = skips
        h32, w32 = h // 32, w // 32
        d4 = decoder(e4, 512, model.decoder4, 256, h32, w32, e3)
        d3 = decoder(d4, 256, model.decoder3, 128, 2 * h32, 2 * w32, e2)
        d2 = decoder(d3, 128, model.decoder2, 64, 4 * h32, 4 * w32, e1)
        d1 = decoder(d2, 64, model.decoder1, 64, 8 * h32, 8 * w32, None)

        # ---- head: Dropout2d -> ConvT k3 s2 (2h+1) -> LeakyReLU -> conv3x3 valid (2h-1) -> LeakyReLU -> conv k2 p1 (2h)
        self.drop_p = float(model.finaldrop1.p)
        self.drop_scale = torch.ones((n, 64), dtype=torch.float32, device=dev)     # keep mask / (1 - p), drawn per step
        self._drop_injected = False
        d1d = S(d1.slab.h, d1.slab.w, 64).view()
        self.ops.append(SimpleOp("snb_scale_nc_nhwc", (N.c_vp(d1.ptr), n, d1.slab.h * d1.slab.w, 64, d1.cstride,
                                                       N.ptr(self.drop_scale), N.c_vp(d1d.ptr), d1d.cstride), (d1, d1d)))
        tape.append(dict(kind='dropout', src=d1, out=d1d))
        slope1, slope2 = model.finalrelu1.negative_slope, model.finalrelu2.negative_slope
        if linear:
            slope1 = slope2 = 1.0        # leaky-ReLU with slope 1 is the identity
        f32 = lambda t: t.detach().float().contiguous()
        f1 = S(h + 1, w + 1, 32)
        wp, idx, shape = packed(model.finaldeconv1.weight, lambda q: E.pack_convT3x3(q, 64, 32, F64))
        self.ops.append(ConvOp(N.CONVT_3X3_S2_FULL, d1d, f1.view(), wp, model.finaldeconv1.bias.detach(), act_slope=slope1))
        tape.append(dict(kind='conv_act', fwd_kind=N.CONVT_3X3_S2_FULL, valid=0, conv=model.finaldeconv1, src=d1d, out=f1.view(),
                         slope=slope1, widx=idx, wshape=shape))
        f3 = S(h - 1, w - 1, 32)
        wp, idx, shape = packed(model.finalconv2.weight, lambda q: E.pack_conv3x3(q, F64))
        self.ops.append(ConvOp(N.CONV_3X3, f1.view(), f3.view(), wp, model.finalconv2.bias.detach(), act_slope=slope2, valid=1))
        tape.append(dict(kind='conv_act', fwd_kind=N.CONV_3X3, valid=1, conv=model.finalconv2, src=f1.view(), out=f3.view(),
                         slope=slope2, widx=idx, wshape=shape))
        self.out = torch.empty((n, h, w), dtype=torch.float32, device=dev)
        pick = torch.zeros(32, dtype=torch.float32, device=dev)
        pick[0] = 1.0
        wp, idx, shape = packed(model.finalconv3.weight, lambda q: E.pack_conv2x2(q, 32, 32, F64))
        self.head_bias = padded_vec(model.finalconv3.bias, 32)                     # finalconv3.bias in a 32-wide vector
        self.ops.append(ConvOp(N.CONV_2X2, f3.view(), None, wp, self.head_bias, relu=False, head=(pick, 0.0, False, self.out)))
        tape.append(dict(kind='conv_head', fwd_kind=N.CONV_2X2, valid=0, conv=model.finalconv3, src=f3.view(), widx=idx, wshape=shape))
        self.flops = sum(op.flops for op in self.ops)

        self._build_backward(tape)
        self.bwd_flops = sum(op.flops for op in self.bwd_ops)
        self.launches = sum(op.launches for op in self.ops)
        self.bwd_launches = sum(op.launches for op in self.bwd_ops)
        self.refresh()

    # ------------------------------------------------------------------------------------------------ backward list
    def _build_backward(self, tape):
        n, dev, model = self.n, self.device, self.model
        S = lambda hh, ww, c: Slab(n, hh, ww, c, dev)
        zeros = lambda c: torch.zeros(c, dtype=torch.float32, device=dev)
        ops = self.bwd_ops
        grads = {}                  # id(activation slab) -> gradient slab (written by the first producer, then accumulated)
        F64 = torch.float64

        def grad_of(view):
            return grads.get(id(view.slab))

        def new_grad(view):
            g = S(view.slab.h, view.slab.w, view.slab.c)
            grads[id(view.slab)] = g
            return g

        def packed(param, fn):
            idx, shape = _index_map(fn, tuple(param.shape), dev)
            t = torch.empty(shape, dtype=torch.bfloat16, device=dev)
            self.repack.add(param.detach(), t.view(-1), idx)
            return t

        def wgrad(node, src, dout, cin=None, cout=None):
            conv = node['conv']
            buf = torch.zeros(node['wshape'], dtype=torch.float32, device=dev)
            self._wgrad_bufs.append(buf)
            self.unpack.add(buf.view(-1), self.grads[conv.weight].view(-1), _inverse_map(node['widx'], conv.weight.numel()))
            ops.append(WgradOp(node['fwd_kind'], src, dout, buf, valid=node['valid'], cin=cin, cout=cout))

        def conv_dgrad_into(kind, gin, target_view, weight, valid=0):
            """input gradient through a forward kernel: writes the target's gradient slab, or adds to it in place (residual
            epilogue) when another consumer has already written it"""
            g = grad_of(target_view)
            first = g is None
            if first:
                g = new_grad(target_view)
            dst = g.view(0, weight.shape[1])
            ops.append(ConvOp(kind, gin, dst, weight, zeros(weight.shape[1]), relu=False, valid=valid,
                              residual=None if first else dst))

        def bias_grad(conv, gslab, channels):
            """channel sums of the output gradient -> the bias slot of the gradient arena"""
            cw = (channels + 7) // 8 * 8
            work = torch.zeros(3 * cw + 2, dtype=torch.float64, device=dev)
            pixels = n * gslab.h * gslab.w
            ops.append(SimpleOp("snb_channel_sum_nhwc", (N.c_vp(gslab.t.data_ptr()), pixels, cw, gslab.c,
                                                         N.ptr(self.grad_slot[conv.bias]), N.ptr(work)), (gslab, work)))

        # d loss / d logits arrives as float [n, h, w]; the head's output gradient slab keeps it in channel 0 of 32
        self.dlogits = torch.zeros((n, self.h, self.w), dtype=torch.float32, device=dev)
        dl = S(self.h, self.w, 32)
        dl.t.zero_()
        ops.append(TorchOp(lambda: dl.t[..., 0].copy_(self.dlogits)))
        ops.append(TorchOp(lambda: torch._foreach_zero_(self._wgrad_bufs), launches=2))   # the wgrad kernels accumulate

        for node in reversed(tape):
            kind = node['kind']
            if kind == 'conv_head':
                conv, src = node['conv'], node['src']
                wgrad(node, src, dl.view(0, 8), cout=1)
                bias_grad(conv, dl, 1)
                conv_dgrad_into(N.CONV_2X2_ADJ, dl.view(), src, packed(conv.weight, lambda q: E.pack_conv2x2_dgrad(q, 32, 32, F64)),
                                valid=1)
            elif kind == 'conv_act':
                conv, src, out = node['conv'], node['src'], node['out']
                g = grad_of(out)
                # through the leaky-ReLU, in place: dz = g * (out > 0 ? 1 : slope)
                ops.append(SimpleOp("snb_ew_nhwc", (N.c_vp(g.t.data_ptr()), g.c, N.c_vp(out.ptr), out.cstride,
                                                    N.c_vp(g.t.data_ptr()), g.c, n * g.h * g.w, g.c, 1, float(node['slope'])),
                                    (g, out)))
                wgrad(node, src, g.view())
                bias_grad(conv, g, conv.bias.numel())
                if node['fwd_kind'] == N.CONV_3X3:       # valid conv3x3: "full" convolution with the flipped kernel
                    conv_dgrad_into(N.CONV_3X3, g.view(), src, packed(conv.weight, lambda q: E.pack_conv_dgrad(q, dtype=F64)),
                                    valid=2)
                else:                                     # ConvTranspose k3 s2 uncropped: 4 taps over the blocked gradient
                    g4 = S((g.h + 1) // 2, (g.w + 1) // 2, 4 * g.c)
                    ops.append(SimpleOp("snb_space_to_depth2", (N.c_vp(g.t.data_ptr()), n, g.h, g.w, g.c, g.c,
                                                                N.c_vp(g4.t.data_ptr()), g4.c), (g, g4)))
                    conv_dgrad_into(N.CONV_2X2_ADJ, g4.view(), src,
                                    packed(conv.weight, lambda q: E.pack_convT3x3_full_dgrad(q, 64, 32, F64)), valid=1)
            elif kind == 'dropout':
                src, out = node['src'], node['out']
                g = grad_of(out)
                gs = new_grad(src)
                ops.append(SimpleOp("snb_scale_nc_nhwc", (N.c_vp(g.t.data_ptr()), n, g.h * g.w, g.c, g.c, N.ptr(self.drop_scale),
                                                          N.c_vp(gs.t.data_ptr()), gs.c), (g, gs)))
            elif kind == 'maxpool':
                src, out = node['src'], node['out']
                g = grad_of(out)
                gs = new_grad(src)
                ops.append(SimpleOp("snb_maxpool3x3s2_backward", (N.c_vp(src.ptr), n, src.slab.h, src.slab.w, src.c, src.cstride,
                                                                  N.c_vp(g.t.data_ptr()), g.c, N.c_vp(gs.t.data_ptr()), gs.c),
                                    (src, g, gs)))
            else:   # conv_bn
                out, raw, bnop, m, conv = node['out'], node['raw'], node['bnop'], node['bn'], node['conv']
                g_out = grad_of(out)
                res, after = node['res'], node['res_after']
                if res is not None and after:
                    # out = act(bn) + res: the skip sees the same gradient.  The decoder output's gradient slab BECOMES the
                    # skip's gradient slab (its batch-norm backward below has consumed it before anyone adds to it)
                    assert grad_of(res) is None
                    grads[id(res.slab)] = g_out
                cpad = raw.c
                draw = S(raw.slab.h, raw.slab.w, cpad)
                rb = res is not None and not after
                dres = None
                if rb:          # act(bn + res): the identity branch receives dz; first writer of that gradient slab
                    assert grad_of(res) is None
                    dres = new_grad(res)
                ops.append(BnBwdOp(raw, g_out.view(), bnop, node['abn'], m.eps, node['slope'], res if rb else None, draw.view(),
                                   dres.view() if rb else None, self.grad_slot[m.weight], self.grad_slot[m.bias]))
                fk, src, dsrc = node['fwd_kind'], node['src'], node['dsrc']
                cin, cout = node.get('real_cin'), node.get('real_cout')
                wgrad(node, src, draw.view(), cin=cin, cout=cout)
                if dsrc is None:
                    continue                                  # the stem: no gradient with respect to the image
                role = node.get('role')
                if role == 's2_shortcut':
                    # conv1x1 / s2 on the first channel quarter of the space-to-depth copy: added in place AFTER the main
                    # branch has written all of dx4 (the main branch comes later in this reverse walk)
                    wd = packed(conv.weight, lambda q: E.pack_conv_dgrad(q, dtype=F64))
                    dx4 = node['dx4']
                    dst = dx4.view(0, wd.shape[1])
                    node['dx4_pending'] = ConvOp(N.CONV_1X1, draw.view(), dst, wd, zeros(wd.shape[1]), relu=False, residual=dst)
                    self._pending_shortcut = node['dx4_pending']
                elif role == 's2_main':
                    dx4 = node['dx4']
                    wd = packed(conv.weight, lambda q: E.pack_conv3x3_s2_dgrad(q, F64))
                    ops.append(ConvOp(N.CONV_2X2_ADJ, draw.view(), dx4.view(), wd, zeros(dx4.c), relu=False))
                    ops.append(self._pending_shortcut)
                    self._pending_shortcut = None
                    g = grad_of(dsrc)
                    acc = 1 if g is not None else 0
                    if g is None:
                        g = new_grad(dsrc)
                    ch = dsrc.c
                    ops.append(SimpleOp("snb_depth_to_space2", (N.c_vp(dx4.t.data_ptr()), n, dsrc.slab.h, dsrc.slab.w, ch, dx4.c,
                                                                N.c_vp(g.t.data_ptr()), g.c, acc), (dx4, g)))
                elif fk == N.CONVT_4X4_S2:
                    mp = node['mp']
                    d4 = S(draw.h // 2, draw.w // 2, 4 * cpad)
                    ops.append(SimpleOp("snb_space_to_depth2", (N.c_vp(draw.t.data_ptr()), n, draw.h, draw.w, cpad, cpad,
                                                                N.c_vp(d4.t.data_ptr()), d4.c), (draw, d4)))
                    conv_dgrad_into(N.CONV_3X3, d4.view(), dsrc, packed(conv.weight, lambda q: E.pack_convT4x4_dgrad(q, mp, mp, F64)))
                else:       # stride-1 conv3x3 (padding 1) / conv1x1: the adjoint convolution
                    k = conv.kernel_size[0]
                    cin_pad = dsrc.slab.c
                    wd = packed(conv.weight, lambda q: E.pack_conv_dgrad(q, cin_pad, cpad, F64))
                    conv_dgrad_into(N.CONV_3X3 if k == 3 else N.CONV_1X1, draw.view(), dsrc, wd)
        ops.append(self.unpack)                      # packed weight gradients -> parameter layouts (one launch)
        self._dl = dl
        self._grad_slabs = grads

    # ------------------------------------------------------------------------------------------------ running
    def refresh(self):
        """Re-pack every bf16 operand (forward and input-gradient weights) and the padded float vectors from the module's
        current parameters: two gather launches, no handle, slab or tensor map changes.  Biases, BatchNorm / ABN weights and
        running statistics that are not padded are read through the parameters' own pointers and need nothing."""
        with torch.no_grad():
            st = N.stream_ptr()
            self.repack(st)
            self.repack32(st)
            for op in self.ops:
                if isinstance(op, BnTrainOp) and op.padded:
                    op.refresh_stats()

    def set_dropout_mask(self, keep):
        """Inject the Dropout2d keep mask (bool / 0-1 tensor [n, 64]) for the next forward (parity tests); otherwise a fresh
        Bernoulli(1 - p) mask is drawn per forward while p > 0."""
        if self.drop_p >= 1.0:
            raise ValueError("Dropout2d with p = 1 drops everything")
        self.drop_scale.copy_(keep.to(self.drop_scale.dtype) / (1.0 - self.drop_p))
        self._drop_injected, self._drop_is_ones = True, False

    def load_nchw(self, x):
        self.x_nchw.copy_(x)

    def _replay(self, name, ops):
        """Run a static op list on the current stream: eagerly the first time (which also serves as the warm-up), captured
        into a CUDA graph right after, replayed from then on."""
        st = N.stream_ptr()
        g = self._graphs.get(name) if self.use_graph else None
        if g is not None:
            g.replay()
            return
        for op in ops:
            op(st)
        if self.use_graph:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cst = N.stream_ptr()
                for op in ops:
                    op(cst)
            self._graphs[name] = g

    def run(self):
        """Train-mode forward on the current stream; logits land in self.out [n, h, w] float32."""
        self.generation += 1
        if self._drop_injected:
            self._drop_injected = False
        elif self.drop_p > 0:
            with torch.no_grad():
                self.drop_scale.bernoulli_(1.0 - self.drop_p).div_(1.0 - self.drop_p)
            self._drop_is_ones = False
        elif not self._drop_is_ones:
            self.drop_scale.fill_(1.0)
            self._drop_is_ones = True
        self._replay('fwd', self.ops)
        if self._nbt is None:                # nn.BatchNorm2d bookkeeping (momentum is fixed, the counter only counts)
            self._nbt = [m.num_batches_tracked for m in self.bn_modules]
        torch._foreach_add_(self._nbt, 1)    # one fused launch instead of one per module
        return self.out

    def backward(self, dlogits=None):
        """Gradients of every parameter given d loss / d logits (float [N, 1, H, W] or [N, H, W]; None = already written
        into self.dlogits) -> {parameter: gradient} (views into the plan's flat gradient arena, valid until the next
        backward)."""
        if dlogits is not None:
            self.dlogits.copy_(dlogits.detach().reshape(self.n, self.h, self.w))
        self._replay('bwd', self.bwd_ops)
        return self.grads

    def grads_in(self, flat):
        """The per-parameter views of a flat tensor laid out like the gradient arena (e.g. a clone of it)."""
        base = self.grad_arena.data_ptr()
        out = {}
        for p, g in self.grads.items():
            off = (g.data_ptr() - base) // 4
            out[p] = flat[off:off + p.numel()].view(p.shape)
        return out
