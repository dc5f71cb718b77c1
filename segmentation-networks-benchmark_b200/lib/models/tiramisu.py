"""FCDenseNet / Tiramisu (reference lib/models/tiramisu.py:93-205) on the native sm_100a engine.

Same constructors (`FCDenseNet`, `FCDenseNet57/67/103`), module tree and state_dict keys (434 entries for
FCDenseNet67); forward in eval mode only (BatchNorm running statistics, Dropout2d inactive).  As in the reference,
`self.softmax` is defined but not applied by forward (lib/models/tiramisu.py:166,183-184).
"""
import torch
import torch.nn as nn

from ... import _native as N
from ...engine import FCDenseNetPlan


def _named_sequential(cls_name, parts):
    """nn.Sequential subclass instance whose children carry the reference's names ('norm', 'relu', 'conv', ...)."""
    seq = type(cls_name, (nn.Sequential,), {})()
    for name, module in parts:
        seq.add_module(name, module)
    return seq


def DenseLayer(in_channels, growth_rate):
    """BN -> ReLU -> conv3x3(in -> growth) -> Dropout2d(0.2)   (parameter holder; tiramisu.py:9-19)."""
    return _named_sequential('DenseLayer', [
        ('norm', nn.BatchNorm2d(in_channels)), ('relu', nn.ReLU(True)),
        ('conv', nn.Conv2d(in_channels, growth_rate, kernel_size=3, stride=1, padding=1, bias=True)),
        ('drop', nn.Dropout2d(0.2))])


def TransitionDown(channels):
    """BN -> ReLU -> conv1x1 -> Dropout2d(0.2) -> MaxPool2d(2)   (tiramisu.py:47-59)."""
    return _named_sequential('TransitionDown', [
        ('norm', nn.BatchNorm2d(num_features=channels)), ('relu', nn.ReLU(inplace=True)),
        ('conv', nn.Conv2d(channels, channels, kernel_size=1, stride=1, padding=0, bias=True)),
        ('drop', nn.Dropout2d(0.2)), ('maxpool', nn.MaxPool2d(2))])


class DenseBlock(nn.Module):
    """`layers[k]` sees the block input plus the k earlier outputs (tiramisu.py:22-44)."""

    def __init__(self, in_channels, growth_rate, n_layers, upsample=False):
        super().__init__()
        self.upsample = upsample
        widths = [in_channels + k * growth_rate for k in range(n_layers)]
        self.layers = nn.ModuleList(DenseLayer(c, growth_rate) for c in widths)


class TransitionUp(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.convTrans = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=3, stride=2, padding=0, bias=True)


def Bottleneck(in_channels, growth_rate, n_layers):
    return _named_sequential('Bottleneck', [('bottleneck', DenseBlock(in_channels, growth_rate, n_layers, upsample=True))])


def _bn(m):
    return (m.weight, m.bias, m.running_mean, m.running_var, m.eps)


class FCDenseNet(nn.Module):
    def __init__(self, in_channels=3, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5,
                 growth_rate=16, out_chans_first_conv=48, n_classes=12):
        super().__init__()
        self.num_classes = n_classes
        self.down_blocks, self.up_blocks, self.growth_rate = down_blocks, up_blocks, growth_rate
        g = growth_rate

        # channel schedule: widths entering each down block, the skips they leave, and what each TransitionUp carries
        down_in, skips, c = [], [], out_chans_first_conv
        for n_layers in down_blocks:
            down_in.append(c)
            c += g * n_layers
            skips.append(c)
        carried = [g * bottleneck_layers] + [g * n for n in up_blocks[:-1]]
        up_in = [carried[i] + skips[-1 - i] for i in range(len(up_blocks))]

        self.firstconv = nn.Conv2d(in_channels, out_chans_first_conv, kernel_size=3, stride=1, padding=1, bias=True)
        self.denseBlocksDown = nn.ModuleList(DenseBlock(ci, g, n) for ci, n in zip(down_in, down_blocks))
        self.transDownBlocks = nn.ModuleList(TransitionDown(sk) for sk in skips)
        self.bottleneck = Bottleneck(skips[-1], g, bottleneck_layers)
        self.transUpBlocks = nn.ModuleList(TransitionUp(cc, cc) for cc in carried)
        last = len(up_blocks) - 1
        self.denseBlocksUp = nn.ModuleList(DenseBlock(ci, g, n, upsample=(i != last))
                                           for i, (ci, n) in enumerate(zip(up_in, up_blocks)))
        self.finalConv = nn.Conv2d(up_in[-1] + g * up_blocks[-1], n_classes, kernel_size=1, stride=1, padding=0, bias=True)
        self.softmax = nn.LogSoftmax(dim=1)     # defined, never applied by forward (as in the reference)

    def _spec(self):
        layer = lambda m: (_bn(m.norm), m.conv.weight, m.conv.bias)
        return dict(
            growth=self.growth_rate,
            first=(self.firstconv.weight, self.firstconv.bias),
            down=[[layer(m) for m in blk.layers] for blk in self.denseBlocksDown],
            trans_down=[(_bn(t.norm), t.conv.weight, t.conv.bias) for t in self.transDownBlocks],
            bottleneck=[layer(m) for m in self.bottleneck.bottleneck.layers],
            trans_up=[(t.convTrans.weight, t.convTrans.bias) for t in self.transUpBlocks],
            up=[[layer(m) for m in blk.layers] for blk in self.denseBlocksUp],
            final=(self.finalConv.weight, self.finalConv.bias))

    def _stamp(self):
        tensors = list(self.parameters()) + list(self.buffers())
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def plan(self, n, h, w, sigmoid=False):
        cache = self.__dict__.setdefault('_plans', {})
        stamp = self._stamp()
        if self.__dict__.get('_plan_stamp') != stamp:
            cache.clear()
            self.__dict__['_plan_stamp'] = stamp
        key = (n, h, w, bool(sigmoid))
        if key not in cache:
            dev = self.finalConv.weight.device
            if dev.type != 'cuda':
                raise RuntimeError("FCDenseNet runs on CUDA devices only (no CPU fallback); call .cuda()")
            with torch.no_grad():
                cache[key] = FCDenseNetPlan(self._spec(), n, h, w, dev, sigmoid)
        return cache[key]

    def forward(self, x):
        N.require_cuda()
        if self.training:
            raise NotImplementedError("FCDenseNet on the native engine is inference only: call .eval()")
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input of shape [N, 3, H, W]")
        with torch.cuda.device(x.device):
            p = self.plan(x.shape[0], x.shape[2], x.shape[3], sigmoid=False)
            p.load_nchw(x.float())
            out = p.run()
        return out.unsqueeze(1).clone()


def FCDenseNet57(n_classes):
    return FCDenseNet(in_channels=3, down_blocks=(4, 4, 4, 4, 4), up_blocks=(4, 4, 4, 4, 4), bottleneck_layers=4,
                      growth_rate=12, out_chans_first_conv=48, n_classes=n_classes)


def FCDenseNet67(n_classes):
    return FCDenseNet(in_channels=3, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5,
                      growth_rate=16, out_chans_first_conv=48, n_classes=n_classes)


def FCDenseNet103(n_classes):
    return FCDenseNet(in_channels=3, down_blocks=(4, 5, 7, 10, 12), up_blocks=(12, 10, 7, 5, 4), bottleneck_layers=15,
                      growth_rate=16, out_chans_first_conv=48, n_classes=n_classes)
