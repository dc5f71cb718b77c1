"""Reference lib/train_utils.py:14-131: AverageMeter and PRCurveMeter (the meter counts on the device: lib/metrics.py)."""
from .metrics import PRCurveMeter  # noqa: F401


class AverageMeter(object):
    """Computes and stores the average and current value (lib/train_utils.py:14-32)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __str__(self):
        return '%.3f' % self.avg
