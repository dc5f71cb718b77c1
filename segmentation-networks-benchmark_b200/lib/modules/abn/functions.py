"""`inplace_abn` autograd function (reference lib/modules/abn/functions.py:62-122) on the native CUDA kernels.

Same call signature, same in-place semantics (x is overwritten with the activated output, the running statistics are
updated in place in training mode) and the same saved tensors (z, var, weight, bias).  The arithmetic the reference
delegates to the external `inplace_abn` extension runs in csrc/abn.cu through snb_abn_forward / snb_abn_backward.
"""
import torch
import torch.autograd as autograd
from torch.autograd.function import once_differentiable

from .... import _native as N

ACT_LEAKY_RELU = "leaky_relu"
ACT_ELU = "elu"
ACT_NONE = "none"
_ACT_CODE = {ACT_NONE: 0, ACT_LEAKY_RELU: 1, ACT_ELU: 2}


def _dims(x):
    if x.dim() < 2:
        raise ValueError("expected an input with a channel dimension, got shape %s" % (tuple(x.shape),))
    n, c = x.shape[0], x.shape[1]
    hw = 1
    for s in x.shape[2:]:
        hw *= s
    return n, c, hw


def _check_input(x):
    N.require_cuda()
    if not x.is_cuda:
        raise RuntimeError("inplace_abn runs on CUDA tensors only (no CPU fallback)")
    if x.dtype != torch.float32:
        raise ValueError("inplace_abn expects float32, got %s" % x.dtype)


class InPlaceABN(autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training=True, momentum=0.1, eps=1e-05,
                activation=ACT_LEAKY_RELU, slope=0.01):
        _check_input(x)
        if activation not in _ACT_CODE:
            raise ValueError("unknown activation %r" % (activation,))
        ctx.training, ctx.momentum, ctx.eps, ctx.activation, ctx.slope = training, momentum, eps, activation, slope
        ctx.affine = weight is not None and bias is not None
        # the reference calls x.contiguous() (functions.py:72): a non-contiguous input is copied, and the copy is what gets
        # normalised and returned
        in_place = x.is_contiguous()
        if not in_place:
            x = x.contiguous()
        n, c, hw = _dims(x)
        weight = weight.contiguous() if ctx.affine else None
        bias = bias.contiguous() if ctx.affine else None
        work = torch.empty(2 * c, dtype=torch.float64, device=x.device)
        if training:
            mean = torch.empty(c, dtype=torch.float32, device=x.device)
            var = torch.empty(c, dtype=torch.float32, device=x.device)
        else:
            mean, var = None, running_var.contiguous()
        with torch.cuda.device(x.device):
            N.check(N.lib().snb_abn_forward(N.ptr(x), n, c, hw, N.ptr(weight), N.ptr(bias), N.ptr(running_mean),
                                            N.ptr(running_var), 1 if training else 0, float(momentum), float(eps),
                                            _ACT_CODE[activation], float(slope), N.ptr(mean), N.ptr(var) if training else N.c_vp(0),
                                            N.ptr(work), N.stream_ptr()))
        # the reference also marks running_mean / running_var dirty in training mode (functions.py:87); torch >= 2 rejects
        # dirty tensors that are not outputs, and buffers need no autograd bookkeeping, so only x is marked
        if in_place:
            ctx.mark_dirty(x)
        ctx.var = var
        ctx.save_for_backward(x, var, weight if ctx.affine else x.new_empty(0), bias if ctx.affine else x.new_empty(0))
        return x

    @staticmethod
    @once_differentiable
    def backward(ctx, dz):
        z, var, weight, bias = ctx.saved_tensors
        dz = dz.contiguous()
        n, c, hw = _dims(z)
        dx = torch.empty_like(dz)
        work = torch.empty(2 * c, dtype=torch.float64, device=z.device)
        dweight = torch.empty(c, dtype=torch.float32, device=z.device) if ctx.affine else None
        dbias = torch.empty(c, dtype=torch.float32, device=z.device) if ctx.affine else None
        with torch.cuda.device(z.device):
            N.check(N.lib().snb_abn_backward(N.ptr(z), N.ptr(dz), n, c, hw, N.ptr(var), N.ptr(weight if ctx.affine else None),
                                             N.ptr(bias if ctx.affine else None), 1 if ctx.training else 0, float(ctx.eps),
                                             _ACT_CODE[ctx.activation], float(ctx.slope), N.ptr(dx), N.ptr(dweight),
                                             N.ptr(dbias), N.ptr(work), N.stream_ptr()))
        return dx, dweight, dbias, None, None, None, None, None, None, None


inplace_abn = InPlaceABN.apply

__all__ = ["inplace_abn"]
