"""Metrics of the hot path (reference lib/metrics.py:9-40) from the fused device reduction of lib/losses.py."""
import numpy as np
import torch
from torch.nn.modules.loss import _Loss

from .. import _native as N
from .losses import _TARGET_DT, fused_sums


class JaccardScore(_Loss):
    """Soft IoU over sigmoid probabilities: I / (U - I + 1e-7) (lib/metrics.py:9-20)."""

    def __init__(self):
        super(JaccardScore, self).__init__()

    def forward(self, output, target):
        s, _ = fused_sums(output, target)
        intersection, union = s[1], s[2] + s[3]
        return (intersection / (union - intersection + 1e-7)).float()

    def __str__(self):
        return 'JaccardScore'


class PixelAccuracy(_Loss):
    """#[(sigmoid(x) > 0.5) == target] / numel, integer count exact (lib/metrics.py:26-40)."""

    def __init__(self):
        super(PixelAccuracy, self).__init__()

    def forward(self, output, target):
        _, c = fused_sums(output, target)
        n_true = c[0] + c[3]  # tp + tn
        if n_true == 0:       # the reference returns the integer zero tensor here (metrics.py:37-38)
            return n_true
        return n_true.float() / target.numel()

    def __str__(self):
        return 'PixelAccuracy'


def confusion_counts(output, target):
    """int64 [tp, fp, fn, tn] at sigmoid(output) > 0.5 -- the additive form that all-reduces across GPUs."""
    return fused_sums(output, target)[1]


def confusion_counts_from_probs(probs, target, threshold=0.5, out=None):
    """Same counts from probabilities/masks already on the device (`mask > 0.5`, inria_submit.py:305)."""
    N.require_cuda()
    p = probs.detach().float().contiguous()
    t = target.detach()
    if t.dtype not in _TARGET_DT:
        t = t.float()
    t = t.contiguous()
    if p.numel() != t.numel():
        raise ValueError("probs and target must have the same number of elements")
    counts = torch.empty(4, dtype=torch.int64, device=p.device) if out is None else out
    with torch.cuda.device(p.device):
        N.check(N.lib().snb_confusion_counts(N.ptr(p), N.ptr(t), _TARGET_DT[t.dtype], p.numel(), float(threshold),
                                             N.ptr(counts), N.ptr(N.reduce_workspace()), N.stream_ptr()))
    return counts


def hard_iou(counts):
    """tp / (tp + fp + fn) from a (possibly all-reduced) count vector."""
    tp, fp, fn = (counts[i].double() for i in range(3))
    return tp / (tp + fp + fn)


class PRCurveMeter(object):
    """tp/tn/fp/fn per threshold (reference lib/train_utils.py:92-131), counted on the device in one pass."""

    def __init__(self, n_thresholds=127):
        self.n_thresholds = n_thresholds
        self.thresholds = np.arange(0., 1., 1. / n_thresholds, dtype=np.float32)
        self._dev = None
        self._thr = None
        self._acc = None

    def _ensure(self, device):
        if self._dev != device:
            self._dev = device
            self._thr = torch.from_numpy(self.thresholds).to(device)
            self._acc = torch.zeros(4, len(self.thresholds), dtype=torch.int64, device=device)

    def reset(self):
        if self._acc is not None:
            self._acc.zero_()

    def update(self, y_pred, y_true):
        N.require_cuda()
        x = y_pred.detach().float().contiguous()
        t = y_true.detach()
        if t.dtype not in _TARGET_DT:
            t = t.float()
        t = t.contiguous()
        self._ensure(x.device)
        a = self._acc
        with torch.cuda.device(x.device):
            N.check(N.lib().snb_pr_curve_update(N.ptr(x), N.ptr(t), _TARGET_DT[t.dtype], x.numel(), N.ptr(self._thr),
                                                len(self.thresholds), N.ptr(a[0]), N.ptr(a[1]), N.ptr(a[2]),
                                                N.ptr(a[3]), N.ptr(N.reduce_workspace()), N.stream_ptr()))

    def _get(self, i):
        if self._acc is None:
            return np.zeros(len(self.thresholds), dtype=np.uint64)
        return self._acc[i].cpu().numpy().astype(np.uint64)

    tp = property(lambda self: self._get(0))
    tn = property(lambda self: self._get(1))
    fp = property(lambda self: self._get(2))
    fn = property(lambda self: self._get(3))

    def precision(self):
        return np.divide(self.tp, self.tp + self.fp)

    def recall(self):
        return np.divide(self.tp, self.tp + self.fn)
