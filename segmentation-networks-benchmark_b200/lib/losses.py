"""Losses of the hot path (reference lib/losses.py:18-101, binary classes) on top of ONE fused device reduction.

`snb_loss_iou_reduce` returns {sum bce, sum p*t, sum p, sum t, sum focal} (+ integer tp/fp/fn/tn) from ONE kernel launch
over logits and targets; the classes below (JaccardLoss, SmoothJaccardLoss, BCEWithSigmoidLoss,
BCEWithLogitsLossAndSmoothJaccard, FocalLossBinary: the binary losses of lib/losses.py:18-101) turn those partial sums
into the reference's scalars with the same formulas, including its quirk of feeding logsigmoid(x) into BCE-with-logits
(lib/losses.py:51-53).
When the logits require grad the scalar carries autograd history: the backward is one elementwise kernel
(`snb_loss_grad`) that rebuilds d loss / d logits from the same four sums, so `loss.backward()` works as in
torch_train.py:186-189 without any host synchronisation.
"""
import torch
from torch.nn.modules.loss import _Loss

from .. import _native as N

_TARGET_DT = {torch.int64: N.DT_I64, torch.uint8: N.DT_U8, torch.float32: N.DT_F32, torch.bool: N.DT_U8}


def _prep(outputs, targets):
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
    return x, t.contiguous()


def fused_sums(outputs, targets, focal_gamma=None, per_element=False):
    """-> (sums float64[5] = [sum bce, sum p*t, sum p, sum t, sum focal], counts int64[4] = [tp, fp, fn, tn]) on the
    device, from ONE kernel launch; with per_element=True also the float tensor of per-element BCE values."""
    x, t = _prep(outputs, targets)
    sums = torch.empty(5, dtype=torch.float64, device=x.device)
    counts = torch.empty(4, dtype=torch.int64, device=x.device)
    elem = torch.empty_like(x) if per_element else None
    with torch.cuda.device(x.device):
        N.check(N.lib().snb_loss_iou_reduce(N.ptr(x), N.ptr(t), _TARGET_DT[t.dtype], x.numel(),
                                            -1.0 if focal_gamma is None else float(focal_gamma), N.ptr(elem), N.ptr(sums),
                                            N.ptr(counts), N.ptr(N.reduce_workspace()), N.stream_ptr()))
    return (sums, counts, elem) if per_element else (sums, counts)


def _loss_grad(outputs, targets, sums, grad_out, per_element, c_bce, c_focal, gamma, c_jac, smooth_num, smooth_den):
    x, t = _prep(outputs, targets)
    grad = torch.empty_like(x)
    g = grad_out.detach().float().contiguous()
    with torch.cuda.device(x.device):
        N.check(N.lib().snb_loss_grad(N.ptr(x), N.ptr(t), _TARGET_DT[t.dtype], x.numel(), N.ptr(sums), N.ptr(g),
                                      1 if per_element else 0, c_bce, c_focal, gamma, c_jac, smooth_num, smooth_den,
                                      N.ptr(grad), N.stream_ptr()))
    return grad.view_as(outputs).to(outputs.dtype)


class _FusedLoss(torch.autograd.Function):
    """c_bce * sum_i BCE_i + c_focal * sum_i focal_i + c_jac * (1 - (I + s_num) / (U - I + s_den)) as one reduction
    (forward) and one elementwise kernel (backward); covers JaccardLoss, SmoothJaccardLoss, BCEWithSigmoidLoss,
    FocalLossBinary and the weighted BCE + Jaccard combination."""

    @staticmethod
    def forward(ctx, outputs, targets, c_bce, c_focal, gamma, c_jac, smooth_num, smooth_den):
        sums, _ = fused_sums(outputs, targets, focal_gamma=gamma if c_focal else None)
        ctx.save_for_backward(outputs, targets, sums)
        ctx.coef = (float(c_bce), float(c_focal), float(gamma), float(c_jac), float(smooth_num), float(smooth_den))
        return _combine(sums, *ctx.coef)

    @staticmethod
    def backward(ctx, grad_out):
        outputs, targets, sums = ctx.saved_tensors
        return (_loss_grad(outputs, targets, sums, grad_out, False, *ctx.coef),) + (None,) * 7


def _combine(sums, c_bce, c_focal, gamma, c_jac, smooth_num, smooth_den):
    out = None
    if c_bce:
        out = (sums[0] * c_bce).float()
    if c_focal:
        f = (sums[4] * c_focal).float()
        out = f if out is None else out + f
    if c_jac:
        jac = 1 - (sums[1] + smooth_num) / (sums[2] + sums[3] - sums[1] + smooth_den)
        j = jac.float() * c_jac
        out = j if out is None else out + j
    return out


def _fused(outputs, targets, c_bce=0.0, c_focal=0.0, gamma=0.0, c_jac=0.0, smooth_num=0.0, smooth_den=0.0):
    if _needs_grad(outputs):
        return _FusedLoss.apply(outputs, targets, c_bce, c_focal, gamma, c_jac, smooth_num, smooth_den)
    sums, _ = fused_sums(outputs, targets, focal_gamma=gamma if c_focal else None)
    return _combine(sums, c_bce, c_focal, gamma, c_jac, smooth_num, smooth_den)


class _ElementBCE(torch.autograd.Function):
    """BCEWithSigmoidLoss(reduce=False): the per-element loss tensor; backward scales the element gradient by the
    per-element upstream gradient."""

    @staticmethod
    def forward(ctx, outputs, targets):
        sums, _, elem = fused_sums(outputs, targets, per_element=True)
        ctx.save_for_backward(outputs, targets, sums)
        return elem.view_as(outputs)

    @staticmethod
    def backward(ctx, grad_out):
        outputs, targets, sums = ctx.saved_tensors
        return _loss_grad(outputs, targets, sums, grad_out, True, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0), None


def _needs_grad(outputs):
    return torch.is_grad_enabled() and outputs.requires_grad


class JaccardLoss(_Loss):
    """1 - I / (U - I + 1e-7), I = sum p*t, U = sum p + sum t (lib/losses.py:18-28)."""

    def __init__(self):
        super(JaccardLoss, self).__init__()

    def fused_coefficients(self, numel):
        return dict(c_jac=1.0, smooth_num=0.0, smooth_den=1e-7)

    def forward(self, output, target):
        return _fused(output, target, **self.fused_coefficients(output.numel()))


class SmoothJaccardLoss(_Loss):
    """1 - (I + smooth) / (U - I + smooth), I = sum p*t, U = sum p + sum t (lib/losses.py:31-43)."""

    def __init__(self, smooth=100):
        super(SmoothJaccardLoss, self).__init__()
        self.smooth = smooth

    def fused_coefficients(self, numel):
        return dict(c_jac=1.0, smooth_num=self.smooth, smooth_den=self.smooth)

    def forward(self, output, target):
        return _fused(output, target, **self.fused_coefficients(output.numel()))


class BCEWithSigmoidLoss(_Loss):
    """BCE-with-logits of logsigmoid(outputs) (lib/losses.py:46-53, the reference's double squash): mean
    (size_average), sum, or the per-element tensor (reduce=False)."""

    def __init__(self, size_average=True, reduce=True):
        super().__init__()
        self.size_average = size_average
        self.reduce = reduce

    def forward(self, outputs, targets):
        if not self.reduce:
            if _needs_grad(outputs):
                return _ElementBCE.apply(outputs, targets)
            return fused_sums(outputs, targets, per_element=True)[2].view_as(outputs)
        return _fused(outputs, targets, **self.fused_coefficients(outputs.numel()))

    def fused_coefficients(self, numel):
        return dict(c_bce=1.0 / numel if self.size_average else 1.0)


class BCEWithLogitsLossAndSmoothJaccard(_Loss):
    """(bce_weight * BCE + jaccard_weight * SmoothJaccard) / (bce_weight + jaccard_weight), lib/losses.py:56-75."""

    def __init__(self, bce_weight=1, jaccard_weight=0.5):
        super(BCEWithLogitsLossAndSmoothJaccard, self).__init__()
        self.bce_loss = BCEWithSigmoidLoss()
        self.jac_loss = SmoothJaccardLoss()
        self.bce_weight = bce_weight
        self.jaccard_weight = jaccard_weight

    def fused_coefficients(self, numel):
        tot = float(self.bce_weight + self.jaccard_weight)
        smooth = self.jac_loss.smooth
        return dict(c_bce=self.bce_weight / (tot * numel), c_jac=self.jaccard_weight / tot, smooth_num=smooth, smooth_den=smooth)

    def forward(self, outputs, targets):
        return _fused(outputs, targets, **self.fused_coefficients(outputs.numel()))         # one pass feeds both terms


class FocalLossBinary(_Loss):
    """mean (or sum) of (1 - pt)^gamma * bce_i, pt = exp(-bce_i), bce_i = the double-squashed BCE above
    (lib/losses.py:78-101; like the reference, `reduce` is accepted and ignored)."""

    def __init__(self, gamma=2, size_average=True, reduce=True):
        super(FocalLossBinary, self).__init__()
        self.gamma = gamma
        self.size_average = size_average
        self.reduce = reduce

    def fused_coefficients(self, numel):
        return dict(c_focal=1.0 / numel if self.size_average else 1.0, gamma=float(self.gamma))

    def forward(self, outputs, targets):
        return _fused(outputs, targets, **self.fused_coefficients(outputs.numel()))


def fused_loss_and_grad(criterion, logits, targets, grad_out, loss_scale=1.0):
    """loss = criterion(logits, targets) and d (loss_scale * loss) / d logits written into `grad_out` (a float tensor shaped
    like logits), without autograd: one reduction launch + one elementwise launch, no host synchronisation.  `criterion`
    is one of the fused losses above.  Returns (loss scalar on the device, sums, counts)."""
    if not hasattr(criterion, 'fused_coefficients') or getattr(criterion, 'reduce', True) is False:
        raise NotImplementedError("train_step needs one of the fused scalar losses of snb_b200.lib.losses")
    k = dict(c_bce=0.0, c_focal=0.0, gamma=0.0, c_jac=0.0, smooth_num=0.0, smooth_den=0.0)
    k.update(criterion.fused_coefficients(logits.numel()))
    sums, counts = fused_sums(logits, targets, focal_gamma=k['gamma'] if k['c_focal'] else None)
    loss = _combine(sums, k['c_bce'], k['c_focal'], k['gamma'], k['c_jac'], k['smooth_num'], k['smooth_den'])
    x, t = _prep(logits, targets)
    if grad_out.dtype != torch.float32 or not grad_out.is_contiguous() or grad_out.numel() != x.numel():
        raise ValueError("grad_out must be a contiguous float tensor with one element per logit")
    s = float(loss_scale)
    with torch.cuda.device(x.device):
        N.check(N.lib().snb_loss_grad(N.ptr(x), N.ptr(t), _TARGET_DT[t.dtype], x.numel(), N.ptr(sums), N.c_vp(0), 0,
                                      k['c_bce'] * s, k['c_focal'] * s, k['gamma'], k['c_jac'] * s, k['smooth_num'],
                                      k['smooth_den'], N.ptr(grad_out), N.stream_ptr()))
    return loss, sums, counts
