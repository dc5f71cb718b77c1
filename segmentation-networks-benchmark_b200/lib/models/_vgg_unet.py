"""Shared body of the TernausNet mirrors: parameter containers with the reference's module tree, forward on
the native engine (snb_b200.engine.VGGUNetPlan)."""
import torch
from torch import nn

from ... import _native as N
from ...engine import PRECISIONS, VGGUNetPlan


def conv3x3(in_, out):
    return nn.Conv2d(in_, out, 3, padding=1)


class ConvRelu(nn.Module):
    """Parameter holder for conv3x3 + ReLU (reference lib/models/unet16.py:12-21)."""

    def __init__(self, in_: int, out: int):
        super().__init__()
        self.conv = conv3x3(in_, out)
        self.activation = nn.ReLU(inplace=True)


class DecoderBlock(nn.Module):
    """conv3x3+ReLU -> ConvTranspose2d(k4,s2,p1)+ReLU (reference lib/models/unet16.py:24-49); same `block.N` keys."""

    def __init__(self, in_channels, middle_channels, out_channels, is_deconv=True):
        super(DecoderBlock, self).__init__()
        self.in_channels = in_channels
        if not is_deconv:
            raise NotImplementedError("the bilinear-upsample decoder variant is not used by any registry model")
        self.block = nn.Sequential(
            ConvRelu(in_channels, middle_channels),
            nn.ConvTranspose2d(middle_channels, out_channels, kernel_size=4, stride=2, padding=1),
            nn.ReLU(inplace=True),
        )


def vgg_features(cfg):
    """torchvision.models.vgg*(...).features layout (Conv2d, ReLU, ..., MaxPool2d) so `encoder.N.*` keys match."""
    layers, c = [], 3
    for v in cfg:
        if v == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(c, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            c = v
    seq = nn.Sequential(*layers)
    for m in seq.modules():  # torchvision's VGG init
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            nn.init.constant_(m.bias, 0)
    return seq


class VGGUNetBase(nn.Module):
    """forward(x: float[N,3,H,W] cuda) -> logits float[N,1,H,W]; plans are cached per input shape.

    `precision` selects the arithmetic of the convolutions: 'bf16' (default; probabilities within 2e-2 of the fp32
    reference) or 'tf32' (fp32 storage, TF32 tensor-core products, fp32 accumulation; within 1e-4)."""

    precision = 'bf16'

    def set_precision(self, precision):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        self.precision = precision
        return self

    def _stages(self):
        raise NotImplementedError

    def _decoders(self):
        return [self.center, self.dec5, self.dec4, self.dec3, self.dec2]

    def _stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def plan(self, n, h, w, sigmoid=False):
        """Cached VGGUNetPlan for this batch/shape; rebuilt when parameters were modified or moved."""
        cache = self.__dict__.setdefault('_plans', {})
        stamp = self._stamp()
        if self.__dict__.get('_plan_stamp') != stamp:
            cache.clear()
            self.__dict__['_plan_stamp'] = stamp
        precision = getattr(self, 'precision', 'bf16')
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        key = (n, h, w, bool(sigmoid), precision)
        if key not in cache:
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError("%s runs on CUDA devices only (no CPU fallback); call .cuda()" % type(self).__name__)
            enc = [[(c.weight, c.bias) for c in st] for st in self._stages()]
            decs = [(d.block[0].conv.weight, d.block[0].conv.bias, d.block[1].weight, d.block[1].bias)
                    for d in self._decoders()]
            with torch.no_grad():
                cache[key] = VGGUNetPlan(enc, decs, (self.dec1.conv.weight, self.dec1.conv.bias),
                                         (self.final.weight, self.final.bias), n, h, w, dev, sigmoid,
                                         PRECISIONS[precision])
        return cache[key]

    def forward(self, x):
        N.require_cuda()
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input of shape [N, 3, H, W]")
        with torch.cuda.device(x.device):
            p = self.plan(x.shape[0], x.shape[2], x.shape[3], sigmoid=False)
            p.load_nchw(x.float())
            out = p.run()
        return out.unsqueeze(1).clone()
