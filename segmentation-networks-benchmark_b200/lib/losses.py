"""Losses of the hot path (reference lib/losses.py:31-75) on top of ONE fused device reduction.

`snb_loss_iou_reduce` returns {sum bce, sum p*t, sum p, sum t} (+ integer tp/fp/fn/tn) in a single pass over
logits and targets; the classes below turn those partial sums into the reference's scalars with the same
formulas, including its quirk of feeding logsigmoid(x) into BCE-with-logits (lib/losses.py:51-53).
When the logits require grad the scalar carries autograd history: the backward is one elementwise kernel
(`snb_loss_grad`) that rebuilds d loss / d logits from the same four sums, so `loss.backward()` works as in
torch_train.py:186-189 without any host synchronisation.
"""
import torch
from torch.nn.modules.loss import _Loss

from .. import _native as N

_TARGET_DT = {torch.int64: N.DT_I64, torch.uint8: N.DT_U8, torch.float32: N.DT_F32, torch.bool: N.DT_U8}


def fused_sums(outputs, targets):
    """-> (sums float64[4] = [sum bce, sum p*t, sum p, sum t], counts int64[4] = [tp, fp, fn, tn]) on the device."""
    N.require_cuda()
    if not outputs.is_cuda or not targets.is_cuda:
        raise RuntimeError("loss/metric reductions run on CUDA tensors only (no CPU fallback)")
    if outputs.numel() != targets.numel():
        raise ValueError("outputs and targets must have the same number of elements")
    x = outputs.detach()
    if x.dtype != torch.float32:
        x = x.float()
    x = x.contiguous()
    t = targets.detach()
    if t.dtype not in _TARGET_DT:
        t = t.float()
    t = t.contiguous()
    sums = torch.empty(4, dtype=torch.float64, device=x.device)
    counts = torch.empty(4, dtype=torch.int64, device=x.device)
    N.check(N.lib().snb_loss_iou_reduce(N.ptr(x), N.ptr(t), _TARGET_DT[t.dtype], x.numel(), N.ptr(sums),
                                        N.ptr(counts), N.stream_ptr()))
    return sums, counts


class _FusedLoss(torch.autograd.Function):
    """c_bce * sum_i BCE_i + c_jac * SmoothJaccard(smooth) as one reduction (forward) and one elementwise kernel
    (backward); covers SmoothJaccardLoss, BCEWithSigmoidLoss and their weighted combination."""

    @staticmethod
    def forward(ctx, outputs, targets, c_bce, c_jac, smooth):
        sums, _ = fused_sums(outputs, targets)
        ctx.save_for_backward(outputs, targets, sums)
        ctx.coef = (float(c_bce), float(c_jac), float(smooth))
        jac = 1 - (sums[1] + smooth) / (sums[2] + sums[3] - sums[1] + smooth)
        return (sums[0] * c_bce).float() + jac.float() * c_jac if c_jac else (sums[0] * c_bce).float()

    @staticmethod
    def backward(ctx, grad_out):
        outputs, targets, sums = ctx.saved_tensors
        c_bce, c_jac, smooth = ctx.coef
        x = outputs.detach().float().contiguous()
        t = targets.detach()
        if t.dtype not in _TARGET_DT:
            t = t.float()
        t = t.contiguous()
        grad = torch.empty_like(x)
        g = grad_out.detach().float().contiguous()
        N.check(N.lib().snb_loss_grad(N.ptr(x), N.ptr(t), _TARGET_DT[t.dtype], x.numel(), N.ptr(sums), N.ptr(g), c_bce,
                                      c_jac, smooth, N.ptr(grad), N.stream_ptr()))
        return grad.view_as(outputs).to(outputs.dtype), None, None, None, None


def _needs_grad(outputs):
    return torch.is_grad_enabled() and outputs.requires_grad


class SmoothJaccardLoss(_Loss):
    """1 - (I + smooth) / (U - I + smooth), I = sum p*t, U = sum p + sum t (lib/losses.py:31-43)."""

    def __init__(self, smooth=100):
        super(SmoothJaccardLoss, self).__init__()
        self.smooth = smooth

    def forward(self, output, target):
        if _needs_grad(output):
            return _FusedLoss.apply(output, target, 0.0, 1.0, self.smooth)
        s, _ = fused_sums(output, target)
        intersection, union = s[1], s[2] + s[3]
        jac = (intersection + self.smooth) / (union - intersection + self.smooth)
        return (1 - jac).float()


class BCEWithSigmoidLoss(_Loss):
    """mean BCE-with-logits of logsigmoid(outputs) (lib/losses.py:46-53, the reference's double squash)."""

    def __init__(self, size_average=True, reduce=True):
        super().__init__()
        self.size_average = size_average
        self.reduce = reduce

    def forward(self, outputs, targets):
        if not self.reduce:
            raise NotImplementedError("reduce=False (per-element loss) is not on the fused path")
        scale = 1.0 / outputs.numel() if self.size_average else 1.0
        if _needs_grad(outputs):
            return _FusedLoss.apply(outputs, targets, scale, 0.0, 0.0)
        s, _ = fused_sums(outputs, targets)
        return (s[0] * scale).float()


class BCEWithLogitsLossAndSmoothJaccard(_Loss):
    """(bce_weight * BCE + jaccard_weight * SmoothJaccard) / (bce_weight + jaccard_weight), lib/losses.py:56-75."""

    def __init__(self, bce_weight=1, jaccard_weight=0.5):
        super(BCEWithLogitsLossAndSmoothJaccard, self).__init__()
        self.bce_loss = BCEWithSigmoidLoss()
        self.jac_loss = SmoothJaccardLoss()
        self.bce_weight = bce_weight
        self.jaccard_weight = jaccard_weight

    def forward(self, outputs, targets):
        if _needs_grad(outputs):
            tot = float(self.bce_weight + self.jaccard_weight)
            return _FusedLoss.apply(outputs, targets, self.bce_weight / (tot * outputs.numel()), self.jaccard_weight / tot,
                                    self.jac_loss.smooth)
        s, _ = fused_sums(outputs, targets)  # one pass feeds both terms
        bce = s[0] / outputs.numel()
        smooth = self.jac_loss.smooth
        jac = 1 - (s[1] + smooth) / (s[2] + s[3] - s[1] + smooth)
        loss1 = bce.float() * self.bce_weight
        loss2 = jac.float() * self.jaccard_weight
        return (loss1 + loss2) / (self.bce_weight + self.jaccard_weight)
