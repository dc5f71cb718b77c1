"""UNet and UNetABN (reference lib/models/unet.py:79-107, unet_abn.py:80-107): (conv3x3 -> norm -> activation) x 2 blocks,
MaxPool2d down, nearest x2 upsampling + torch.cat([skip, upsampled]) up, Dropout2d before the 1x1 output conv.

Same constructors, module trees and state_dict keys as the reference (`inc.conv.conv.N`, `downK.mpconv.1.conv.N`,
`upK.conv.conv.N`, `outc.conv`; N = 0/1/3/4 for Conv / BatchNorm pairs, 0/1/2/3 with InPlaceABN).  The forward runs on the
native sm_100a engine in eval mode (snb_b200.engine.UNetPlan): norms folded, pooling / upsampling / concat / head fused.
"""
import torch
from torch import nn

from ... import _native as N
from ...engine import PRECISIONS, UNetPlan
from ..modules.abn import InPlaceABN


class double_conv(nn.Module):
    """(conv => BN => ReLU) * 2, or (conv => InPlaceABN) * 2"""

    def __init__(self, in_ch, out_ch, abn=False):
        super(double_conv, self).__init__()
        if abn:
            self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, 3, padding=1), InPlaceABN(out_ch),
                                      nn.Conv2d(out_ch, out_ch, 3, padding=1), InPlaceABN(out_ch))
        else:
            self.conv = nn.Sequential(nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True),
                                      nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True))
        self.abn = abn

    def layers(self):
        i2 = 2 if self.abn else 3
        out = []
        for ci in (0, i2):
            c, m = self.conv[ci], self.conv[ci + 1]
            out.append((c.weight, c.bias, (m.weight, m.bias, m.running_mean, m.running_var, m.eps, self.abn)))
        return tuple(out)


class inconv(nn.Module):
    def __init__(self, in_ch, out_ch, abn=False):
        super(inconv, self).__init__()
        self.conv = double_conv(in_ch, out_ch, abn)


class down(nn.Module):
    def __init__(self, in_ch, out_ch, abn=False):
        super(down, self).__init__()
        self.mpconv = nn.Sequential(nn.MaxPool2d(2), double_conv(in_ch, out_ch, abn))


class up(nn.Module):
    def __init__(self, in_ch, out_ch, upsample=True, abn=False):
        super(up, self).__init__()
        if not upsample:
            raise NotImplementedError("the ConvTranspose2d(k2, s2) up-sampling variant is not built (the registry uses upsample=True)")
        self.up = nn.Upsample(scale_factor=2, mode='nearest')
        self.conv = double_conv(in_ch, out_ch, abn)


class outconv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super(outconv, self).__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, 1)


class UNet(nn.Module):
    _abn = False
    precision = 'bf16'

    def __init__(self, n_channels=3, n_classes=1, n_filters=32, upsample=True):
        super().__init__()
        a = self._abn
        self.num_classes = n_classes
        self.inc = inconv(n_channels, n_filters, a)
        self.down1 = down(n_filters, n_filters * 2, a)
        self.down2 = down(n_filters * 2, n_filters * 4, a)
        self.down3 = down(n_filters * 4, n_filters * 8, a)
        self.down4 = down(n_filters * 8, n_filters * 8, a)
        self.up1 = up(n_filters * 16, n_filters * 4, upsample, a)
        self.up2 = up(n_filters * 8, n_filters * 2, upsample, a)
        self.up3 = up(n_filters * 4, n_filters, upsample, a)
        self.up4 = up(n_filters * 2, n_filters, upsample, a)
        self.finaldrop = nn.Dropout2d(p=0.5)
        self.outc = outconv(n_filters, n_classes)

    def set_precision(self, precision):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        self.precision = precision
        return self

    def _stamp(self):
        tensors = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def plan(self, n, h, w, sigmoid=False):
        cache = self.__dict__.setdefault('_plans', {})
        stamp = self._stamp()
        if self.__dict__.get('_plan_stamp') != stamp:
            cache.clear()
            self.__dict__['_plan_stamp'] = stamp
        key = (n, h, w, bool(sigmoid), self.precision)
        if key not in cache:
            dev = self.outc.conv.weight.device
            if dev.type != 'cuda':
                raise RuntimeError("%s runs on CUDA devices only (no CPU fallback); call .cuda()" % type(self).__name__)
            blocks = ([self.inc.conv.layers()] + [getattr(self, 'down%d' % i).mpconv[1].layers() for i in range(1, 5)] +
                      [getattr(self, 'up%d' % i).conv.layers() for i in range(1, 5)])
            slope = self.inc.conv.conv[1].slope if self._abn else 0.0
            with torch.no_grad():
                cache[key] = UNetPlan(blocks, (self.outc.conv.weight, self.outc.conv.bias), n, h, w, dev, sigmoid,
                                      PRECISIONS[self.precision], act_slope=slope)
        return cache[key]

    def forward(self, x):
        N.require_cuda()
        if self.training:
            raise NotImplementedError("%s on the native engine is inference only: call .eval()" % type(self).__name__)
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input of shape [N, 3, H, W]")
        with torch.cuda.device(x.device):
            p = self.plan(x.shape[0], x.shape[2], x.shape[3], sigmoid=False)
            p.load_nchw(x.float())
            out = p.run()
        return out.unsqueeze(1).clone()


class UNetABN(UNet):
    _abn = True
